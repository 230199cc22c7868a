/*
 * TEST INFRASTRUCTURE -- not part of the product path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this.
 *
 * Plain-C restatement of the reference's DecodeBox.non_max_suppression up to the host-side un-letterbox
 * (reference utils/utils_bbox.py:87-168) including the greedy NMS of its un-vendored dependency
 * torchvision.ops.nms (called at utils/utils_bbox.py:145; torchvision is unpinned in requirements.txt:2,
 * 0.26.0+cu128 in this image).  torchvision's algorithm, restated from its published sources
 * (torchvision/csrc/ops/cpu/nms_kernel.cpp, torchvision/csrc/ops/cuda/nms_kernel.cu):
 *   order = stable descending sort of the scores; walk it; a box not yet suppressed is kept and
 *   suppresses every later box j with inter/(area_i + area_j - inter) > iou_threshold.
 * iou_mode 0 (CPU op): both areas rounded to fp32 separately, the fp32 IoU is compared with the DOUBLE threshold.
 * iou_mode 1 (CUDA op, as compiled in this image: SASS of nms_kernel_impl<float> shows FMUL/FFMA/FADD,
 *   an IEEE divide and FSETP.GT against F2F.F32.F64(threshold)): area_i + area_j is fma(wj, hj, area_i),
 *   compared with the threshold rounded to float.
 * Parity pinning: the reference holds no golden vectors for this path; this file is pinned against
 * torch.ops.torchvision.nms itself (tests/test_oracle_cpu.py here, tests/test_nms_gpu.py on the GPU box).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int cls;
  float score;
  int idx;
} cand_t;

static int cmp_cand(const void* pa, const void* pb) {
  const cand_t* a = (const cand_t*)pa;
  const cand_t* b = (const cand_t*)pb;
  if (a->cls != b->cls) return a->cls < b->cls ? -1 : 1;          /* unique labels ascending (:130,:136) */
  if (a->score != b->score) return a->score > b->score ? -1 : 1;  /* descending score */
  return a->idx < b->idx ? -1 : (a->idx > b->idx ? 1 : 0);        /* stable: lower index first */
}

static int suppresses(const float* a, const float* b, double thr, int mode) {
  /* a: kept (earlier) box, b: later box; (x1,y1,x2,y2) */
  float left = a[0] > b[0] ? a[0] : b[0], right = a[2] < b[2] ? a[2] : b[2];
  float top = a[1] > b[1] ? a[1] : b[1], bottom = a[3] < b[3] ? a[3] : b[3];
  volatile float w = right - left, h = bottom - top;
  float ww = w > 0.0f ? w : 0.0f, hh = h > 0.0f ? h : 0.0f;
  volatile float inter = ww * hh;
  volatile float wa = a[2] - a[0], ha = a[3] - a[1], wb = b[2] - b[0], hb = b[3] - b[1];
  volatile float sa = wa * ha;
  if (mode == 1) {
    volatile float sum = fmaf(wb, hb, sa);
    volatile float den = sum - inter;
    volatile float iou = inter / den;
    return iou > (float)thr;
  } else {
    volatile float sb = wb * hb;
    volatile float sum = sa + sb;
    volatile float den = sum - inter;
    volatile float iou = inter / den;
    return (double)iou > thr;
  }
}

/*
 * One image.  pred [A][4+nc] is rewritten in place from (cx,cy,w,h) to (x1,y1,x2,y2) (:92-97).
 * out_det [A][6] / out_idx [A] receive the kept rows in the reference's order; returns their count.
 */
int dcfa_oracle_nms_image(float* pred, int A, int nc, float conf_thres, double nms_thres, int iou_mode,
                          float* out_det, int32_t* out_idx, int32_t* out_cand) {
  const int row = 4 + nc;
  cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(A > 0 ? A : 1));
  int n = 0;
  for (int a = 0; a < A; ++a) {
    float* r = pred + (size_t)a * row;
    volatile float hw = r[2] / 2.0f, hh = r[3] / 2.0f;
    volatile float x1 = r[0] - hw, y1 = r[1] - hh, x2 = r[0] + hw, y2 = r[1] + hh;
    r[0] = x1; r[1] = y1; r[2] = x2; r[3] = y2;
    float best = r[4];
    int bi = 0;
    for (int k = 1; k < nc; ++k) {
      if (best != best) break; /* torch.max propagates NaN */
      if (r[4 + k] > best || r[4 + k] != r[4 + k]) { best = r[4 + k]; bi = k; }
    }
    if (best >= conf_thres) { c[n].cls = bi; c[n].score = best; c[n].idx = a; ++n; }
  }
  if (out_cand) *out_cand = n;
  qsort(c, (size_t)n, sizeof(cand_t), cmp_cand);
  unsigned char* removed = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    if (removed[i]) continue;
    const float* bi_ = pred + (size_t)c[i].idx * row;
    float* o = out_det + (size_t)kept * 6;
    o[0] = bi_[0]; o[1] = bi_[1]; o[2] = bi_[2]; o[3] = bi_[3];
    o[4] = c[i].score;
    o[5] = (float)c[i].cls;
    out_idx[kept++] = c[i].idx;
    for (int j = i + 1; j < n && c[j].cls == c[i].cls; ++j) {
      if (removed[j]) continue;
      if (suppresses(bi_, pred + (size_t)c[j].idx * row, nms_thres, iou_mode)) removed[j] = 1;
    }
  }
  free(removed);
  free(c);
  return kept;
}

/* Batched convenience wrapper with the same output layout as dcfa_nms (include/dcfa_b200.h). */
void dcfa_oracle_nms(float* pred, int B, int A, int nc, float conf_thres, double nms_thres, int iou_mode,
                     float* out_det, int32_t* out_idx, int32_t* out_cnt, int32_t* out_cand) {
  for (int b = 0; b < B; ++b) {
    out_cnt[b] = dcfa_oracle_nms_image(pred + (size_t)b * A * (4 + nc), A, nc, conf_thres, nms_thres, iou_mode,
                                       out_det + (size_t)b * A * 6, out_idx + (size_t)b * A,
                                       out_cand ? out_cand + b : 0);
  }
}
