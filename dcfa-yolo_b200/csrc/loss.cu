// dcfa_yolo_loss: the criterion the reference applies to the hot path's outputs in its validation loop
// (utils/utils_fit_mul.py:78-92: eval-mode forward under no_grad, then `yolo_loss(outputs, bboxes)`), forward only:
//   Loss.__call__            nets/yolo_training.py:371-430   (decode, assigner, BCE / CIoU / DFL terms, gains)
//   TaskAlignedAssigner      nets/yolo_training.py:75-225    (alignment metric, in-box test, top-10, conflict resolution, targets)
//   bbox_iou (CIoU)          nets/yolo_training.py:227-262
//   bbox2dist, BboxLoss      nets/yolo_training.py:267-303
// SURVEY 8(f) N4.  The reference materialises (B, G, A) tensors for overlaps, metrics, masks and one-hot targets and
// runs ~120 small torch kernels with several host synchronisations; here the same arithmetic is four launches and
// nothing of size B*G*A is ever written:
//   1. loss_decode_kernel   one thread per (image, anchor): DFL expectation -> predicted box, sigmoid scores, the
//                           target-independent part of the BCE sum
//   2. loss_assign_kernel   one CTA per (image, ground-truth box): metric of every anchor into shared memory, ten
//                           rounds of block-wide arg-max (largest value, lowest anchor index) -> <= 10 candidates
//   3. loss_resolve_kernel  one CTA per image: anchors claimed by several boxes go to the box of highest overlap
//                           (first maximum over ALL boxes, as `overlaps.argmax(1)` does), per-box maxima of metric and
//                           overlap, normalised target score per foreground anchor, its BCE / CIoU / DFL terms
//   4. loss_final_kernel    fixed-order sums, target_scores_sum clamp, gains
// Every sum is taken in a fixed order (no floating-point atomics): the result is deterministic run to run.
// fp32 with the reference's operation order for everything that decides the assignment (`_rn` intrinsics: no FMA
// contraction); partial sums in fp64.  torch.topk's order among EQUAL values is unspecified in the reference; this
// kernel takes the lowest anchor index first (oracle/loss.py documents when that can matter).
#include <math.h>

#include "common.cuh"

namespace dcfa {
namespace {

constexpr int kTopK = 10;          // Loss.__init__ :333 (topk=10, alpha=0.5, beta=6.0)
constexpr int kBins = 16;          // reg_max
constexpr float kEpsIou = 1e-7f;   // bbox_iou eps
constexpr float kEpsTal = 1e-9f;   // assigner eps

struct LossArgs {
  const float* map[3];   // [B, 64 + nc, H_l, W_l] fp32 NCHW (YoloBody.forward's `x`)
  const float* gt;       // [B, G, 5] (class, x1, y1, x2, y2) input pixels, zero padded (Loss.preprocess)
  float4* pbox;          // [B, A] predicted boxes, grid units (x1, y1, x2, y2)
  float* sig;            // [B, nc, A] sigmoid(class logits)
  unsigned* aword;       // [B, A] first candidate slot claiming the anchor | bit 31: claimed more than once
  int* cand_a;           // [B, G, 10] anchor of the k-th candidate, -1 if it is not a positive
  float* cand_ov;        // [B, G, 10] its overlap (CIoU clamped at 0)
  float* cand_al;        // [B, G, 10] its alignment metric
  double* part1;         // [B * nblk1] per-block sums of the target-free BCE part
  double* img_part;      // [B, 5] sum(norm), sum((1-iou) w), sum(dfl w), sum(BCE delta), foreground anchors
  float* out;            // [8] box*7.5, cls*0.5, dfl*1.5, total, target_scores_sum, foreground anchors
  int B, nc, no, A, G, nblk1;
  int hw[3], w[3];
  float stride[3];
};

__device__ __forceinline__ float ciou_rn(float x11, float y11, float x12, float y12, float x21, float y21, float x22, float y22) {
  // bbox_iou(box1, box2, xywh=False, CIoU=True), one rounding per reference operation
  const float w1 = __fsub_rn(x12, x11), h1 = __fadd_rn(__fsub_rn(y12, y11), kEpsIou);
  const float w2 = __fsub_rn(x22, x21), h2 = __fadd_rn(__fsub_rn(y22, y21), kEpsIou);
  const float iw = fmaxf(__fsub_rn(fminf(x12, x22), fmaxf(x11, x21)), 0.0f);
  const float ih = fmaxf(__fsub_rn(fminf(y12, y22), fmaxf(y11, y21)), 0.0f);
  const float inter = __fmul_rn(iw, ih);
  const float uni = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(w1, h1), __fmul_rn(w2, h2)), inter), kEpsIou);
  const float iou = __fdiv_rn(inter, uni);
  const float cw = __fsub_rn(fmaxf(x12, x22), fminf(x11, x21));
  const float ch = __fsub_rn(fmaxf(y12, y22), fminf(y11, y21));
  const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(cw, cw), __fmul_rn(ch, ch)), kEpsIou);
  const float dx = __fsub_rn(__fsub_rn(__fadd_rn(x21, x22), x11), x12);
  const float dy = __fsub_rn(__fsub_rn(__fadd_rn(y21, y22), y11), y12);
  const float rho2 = __fdiv_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), 4.0f);
  const float da = __fsub_rn(atanf(__fdiv_rn(w2, h2)), atanf(__fdiv_rn(w1, h1)));
  const float v = __fmul_rn(0.40528473456935109f, __fmul_rn(da, da));   // 4 / pi^2
  const float alpha = __fdiv_rn(v, __fadd_rn(__fsub_rn(v, iou), 1.0000001f));
  return __fsub_rn(iou, __fadd_rn(__fdiv_rn(rho2, c2), __fmul_rn(v, alpha)));
}

__device__ __forceinline__ void anchor_of(const LossArgs& p, int a, int& l, int& pos, float& ax, float& ay) {
  l = 0; pos = a;
  if (pos >= p.hw[0]) { pos -= p.hw[0]; l = 1; }
  if (l == 1 && pos >= p.hw[1]) { pos -= p.hw[1]; l = 2; }
  const int y = pos / p.w[l];
  ax = (float)(pos - y * p.w[l]) + 0.5f;
  ay = (float)y + 0.5f;
}

// overlap (clamped CIoU), alignment metric and in-box flag of anchor `a` for one ground-truth box, as
// get_box_metrics (:150-173) and select_candidates_in_gts (:12-38) compute them
__device__ __forceinline__ void box_metric(const LossArgs& p, int b, int a, float label_score_sig, const float* g, float& ov, float& al,
                                           bool& inside) {
  int l, pos;
  float ax, ay;
  anchor_of(p, a, l, pos, ax, ay);
  const float s = p.stride[l];
  const float4 q = p.pbox[(int64_t)b * p.A + a];
  const float c = ciou_rn(g[0], g[1], g[2], g[3], __fmul_rn(q.x, s), __fmul_rn(q.y, s), __fmul_rn(q.z, s), __fmul_rn(q.w, s));
  ov = fmaxf(c, 0.0f);
  al = __fmul_rn(__fsqrt_rn(label_score_sig), powf(ov, 6.0f));
  const float px = __fmul_rn(ax, s), py = __fmul_rn(ay, s);
  const float d = fminf(fminf(__fsub_rn(px, g[0]), __fsub_rn(py, g[1])), fminf(__fsub_rn(g[2], px), __fsub_rn(g[3], py)));
  inside = d > kEpsTal;
}

template <int N>
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  // fixed-order tree: lanes by xor shuffles, warps in index order by thread 0
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < N / 32; ++i) t += scratch[i];
  return t;   // valid in thread 0
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_decode_kernel(const LossArgs p) {
  __shared__ double scratch[8];
  const int b = blockIdx.y;
  const int a = blockIdx.x * 256 + threadIdx.x;
  double bce = 0.0;
  if (a < p.A) {
    int l, pos;
    float ax, ay;
    anchor_of(p, a, l, pos, ax, ay);
    const int hw = p.hw[l];
    const float* src = p.map[l] + (int64_t)b * p.no * hw + pos;
    float d[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {   // Loss.bbox_decode :362-369
      float v[kBins];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < kBins; ++k) {
        v[k] = __ldg(src + (int64_t)(side * kBins + k) * hw);
        mx = fmaxf(mx, v[k]);
      }
      float den = 0.0f;
#pragma unroll
      for (int k = 0; k < kBins; ++k) {
        v[k] = expf(__fsub_rn(v[k], mx));
        den = __fadd_rn(den, v[k]);
      }
      float e = 0.0f;
#pragma unroll
      for (int k = 1; k < kBins; ++k) e = __fadd_rn(e, __fmul_rn(__fdiv_rn(v[k], den), (float)k));
      d[side] = e;
    }
    p.pbox[(int64_t)b * p.A + a] = make_float4(__fsub_rn(ax, d[0]), __fsub_rn(ay, d[1]), __fadd_rn(ax, d[2]), __fadd_rn(ay, d[3]));
    p.aword[(int64_t)b * p.A + a] = 0x7fffffffu;
    for (int c = 0; c < p.nc; ++c) {
      const float x = __ldg(src + (int64_t)(4 * kBins + c) * hw);
      p.sig[((int64_t)b * p.nc + c) * p.A + a] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
      // BCEWithLogits with target 0: x - log_sigmoid(x)
      const float ls = __fsub_rn(fminf(x, 0.0f), log1pf(expf(-fabsf(x))));
      bce += (double)__fsub_rn(x, ls);
    }
  }
  const double t = block_sum<256>(bce, scratch);
  if (threadIdx.x == 0) p.part1[(int64_t)b * p.nblk1 + blockIdx.x] = t;
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_assign_kernel(const LossArgs p) {
  extern __shared__ float metric[];   // [A] alignment metric x in-box flag; -1 once taken
  __shared__ unsigned long long wbest[8];
  __shared__ float gbox[4];
  const int g = blockIdx.x, b = blockIdx.y;
  const float* gt = p.gt + ((int64_t)b * p.G + g) * 5;
  const int64_t cbase = ((int64_t)b * p.G + g) * kTopK;
  if (threadIdx.x < 4) gbox[threadIdx.x] = gt[1 + threadIdx.x];
  __syncthreads();
  // mask_gt (:401): padded rows are all zero
  const bool valid = __fadd_rn(__fadd_rn(__fadd_rn(gbox[0], gbox[1]), gbox[2]), gbox[3]) > 0.0f;
  if (!valid) {
    if (threadIdx.x < kTopK) p.cand_a[cbase + threadIdx.x] = -1;
    return;
  }
  const int label = min(max((int)gt[0], 0), p.nc - 1);   // callers validate labels; the clamp only keeps a bad one inside the tensor
  const float* sig = p.sig + ((int64_t)b * p.nc + label) * p.A;
  for (int a = threadIdx.x; a < p.A; a += 256) {
    float ov, al;
    bool inside;
    box_metric(p, b, a, sig[a], gbox, ov, al, inside);
    metric[a] = inside ? al : __fmul_rn(al, 0.0f);
  }
  __syncthreads();
  for (int k = 0; k < kTopK; ++k) {   // select_topk_candidates (:181-204)
    unsigned long long best = 0ull;
    for (int a = threadIdx.x; a < p.A; a += 256) {
      const float m = metric[a];
      if (m >= 0.0f) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(m) << 32) | (unsigned)(0xffffffffu - (unsigned)a);
        best = key > best ? key : best;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < 8; ++i) best = wbest[i] > best ? wbest[i] : best;
      const int a = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
      float ov, al;
      bool inside;
      box_metric(p, b, a, sig[a], gbox, ov, al, inside);
      p.cand_a[cbase + k] = inside ? a : -1;   // mask_pos = mask_topk * mask_in_gts * mask_gt (:146)
      p.cand_ov[cbase + k] = ov;
      p.cand_al[cbase + k] = al;
      metric[a] = -1.0f;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------------------
constexpr int kResolveThreads = 512;

__global__ void __launch_bounds__(kResolveThreads) loss_resolve_kernel(const LossArgs p) {
  extern __shared__ unsigned char smem_raw[];
  const int C = p.G * kTopK;
  int* fin_g = reinterpret_cast<int*>(smem_raw);                 // [C] box the candidate's anchor finally belongs to, -1: not a handler
  float* fin_ov = reinterpret_cast<float*>(fin_g + C);           // [C]
  float* fin_al = fin_ov + C;                                    // [C]
  int* pos_al = reinterpret_cast<int*>(fin_al + C);              // [G] max metric over the box's positives (bits of a float >= 0)
  int* pos_ov = pos_al + p.G;                                    // [G]
  __shared__ double scratch[kResolveThreads / 32];
  const int b = blockIdx.x;
  const float* gt = p.gt + (int64_t)b * p.G * 5;
  unsigned* aword = p.aword + (int64_t)b * p.A;
  const int* cand_a = p.cand_a + (int64_t)b * C;
  for (int g = threadIdx.x; g < p.G; g += kResolveThreads) { pos_al[g] = 0; pos_ov[g] = 0; }
  for (int c = threadIdx.x; c < C; c += kResolveThreads) {
    const int a = cand_a[c];
    if (a >= 0) atomicMin(&aword[a], (unsigned)c);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kResolveThreads) {
    const int a = cand_a[c];
    if (a >= 0 && (__ldcg(&aword[a]) & 0x7fffffffu) != (unsigned)c) atomicOr(&aword[a], 0x80000000u);   // L2 reads: the atomics live there
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kResolveThreads) {
    const int a = cand_a[c];
    int g = -1;
    float ov = 0.0f, al = 0.0f;
    if (a >= 0) {
      const unsigned w = __ldcg(&aword[a]);
      if ((w & 0x7fffffffu) == (unsigned)c) {
        if (!(w >> 31)) {
          g = c / kTopK;
          ov = p.cand_ov[(int64_t)b * C + c];
          al = p.cand_al[(int64_t)b * C + c];
        } else {
          // select_highest_overlaps (:56-69): first maximum of the overlaps over ALL boxes of the image, padded rows included
          float best = -1.0f;
          for (int j = 0; j < p.G; ++j) {
            const float* gj = gt + j * 5;
            float o, m;
            bool inside;
            box_metric(p, b, a, p.sig[((int64_t)b * p.nc + min(max((int)gj[0], 0), p.nc - 1)) * p.A + a], gj + 1, o, m, inside);
            if (o > best) { best = o; g = j; ov = o; al = m; }
          }
        }
        atomicMax(&pos_al[g], __float_as_int(al));
        atomicMax(&pos_ov[g], __float_as_int(ov));
      }
    }
    fin_g[c] = g; fin_ov[c] = ov; fin_al[c] = al;
  }
  __syncthreads();
  double s_norm = 0.0, s_iou = 0.0, s_dfl = 0.0, s_bce = 0.0, s_cnt = 0.0;
  for (int c = threadIdx.x; c < C; c += kResolveThreads) {
    const int g = fin_g[c];
    if (g < 0) continue;
    const int a = cand_a[c];
    const float* gj = gt + g * 5;
    // norm_align_metric (:118-126): the anchor's only non-zero entry of align * pos_overlaps / (pos_align + eps)
    const float norm = __fdiv_rn(__fmul_rn(fin_al[c], __int_as_float(pos_ov[g])), __fadd_rn(__int_as_float(pos_al[g]), kEpsTal));
    int l, pos;
    float ax, ay;
    anchor_of(p, a, l, pos, ax, ay);
    const int hw = p.hw[l];
    const float s = p.stride[l];
    const float* src = p.map[l] + (int64_t)b * p.no * hw + pos;
    // BCE (:411): the one element of this anchor whose target is not zero
    {
      const float x = __ldg(src + (int64_t)(4 * kBins + min(max((int)gj[0], 0), p.nc - 1)) * hw);
      const float ls = __fsub_rn(fminf(x, 0.0f), log1pf(expf(-fabsf(x))));
      const float with_t = __fsub_rn(__fmul_rn(__fsub_rn(1.0f, norm), x), ls);
      s_bce += (double)with_t - (double)__fsub_rn(x, ls);
    }
    // BboxLoss (:278-291): CIoU(pred, target) in grid units, weight = target_scores.sum(-1) = norm
    const float tx1 = __fdiv_rn(gj[1], s), ty1 = __fdiv_rn(gj[2], s), tx2 = __fdiv_rn(gj[3], s), ty2 = __fdiv_rn(gj[4], s);
    const float4 q = p.pbox[(int64_t)b * p.A + a];
    const float iou = ciou_rn(q.x, q.y, q.z, q.w, tx1, ty1, tx2, ty2);
    s_iou += (double)__fmul_rn(__fsub_rn(1.0f, iou), norm);
    // DFL (:267-270, :293-303)
    const float tgt[4] = {__fsub_rn(ax, tx1), __fsub_rn(ay, ty1), __fsub_rn(tx2, ax), __fsub_rn(ty2, ay)};
    float dfl = 0.0f;
#pragma unroll
    for (int side = 0; side < 4; ++side) {
      const float t = fminf(fmaxf(tgt[side], 0.0f), 14.99f);
      const int tl = (int)t;
      const float wl = __fsub_rn((float)(tl + 1), t), wr = __fsub_rn(1.0f, wl);
      float v[kBins];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < kBins; ++k) {
        v[k] = __ldg(src + (int64_t)(side * kBins + k) * hw);
        mx = fmaxf(mx, v[k]);
      }
      float den = 0.0f, xl = 0.0f, xr = 0.0f;
#pragma unroll
      for (int k = 0; k < kBins; ++k) {
        den = __fadd_rn(den, expf(__fsub_rn(v[k], mx)));
        xl = (k == tl) ? v[k] : xl;
        xr = (k == tl + 1) ? v[k] : xr;
      }
      const float lse = __fadd_rn(mx, logf(den));
      dfl = __fadd_rn(dfl, __fadd_rn(__fmul_rn(__fsub_rn(lse, xl), wl), __fmul_rn(__fsub_rn(lse, xr), wr)));
    }
    s_dfl += (double)__fmul_rn(__fdiv_rn(dfl, 4.0f), norm);
    s_norm += (double)norm;
    s_cnt += 1.0;
  }
  double* o = p.img_part + (int64_t)b * 5;
  double t;
  t = block_sum<kResolveThreads>(s_norm, scratch); if (threadIdx.x == 0) o[0] = t;
  t = block_sum<kResolveThreads>(s_iou, scratch);  if (threadIdx.x == 0) o[1] = t;
  t = block_sum<kResolveThreads>(s_dfl, scratch);  if (threadIdx.x == 0) o[2] = t;
  t = block_sum<kResolveThreads>(s_bce, scratch);  if (threadIdx.x == 0) o[3] = t;
  t = block_sum<kResolveThreads>(s_cnt, scratch);  if (threadIdx.x == 0) o[4] = t;
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_final_kernel(const LossArgs p) {
  __shared__ double scratch[8];
  double base = 0.0;
  const int n1 = p.B * p.nblk1;
  for (int i = threadIdx.x; i < n1; i += 256) base += p.part1[i];
  base = block_sum<256>(base, scratch);
  if (threadIdx.x != 0) return;
  double s[5] = {0, 0, 0, 0, 0};
  if (p.G > 0)
    for (int b = 0; b < p.B; ++b)
      for (int j = 0; j < 5; ++j) s[j] += p.img_part[(int64_t)b * 5 + j];
  const float tss = fmaxf((float)s[0], 1.0f);                      // max(target_scores.sum(), 1) (:407)
  const float cls = __fmul_rn(__fdiv_rn((float)(base + s[3]), tss), 0.5f);
  float box = 0.0f, dfl = 0.0f;
  if (s[4] > 0.0) {                                                // `if fg_mask.sum():` (:414)
    box = __fmul_rn(__fdiv_rn((float)s[1], tss), 7.5f);
    dfl = __fmul_rn(__fdiv_rn((float)s[2], tss), 1.5f);
  }
  p.out[0] = box; p.out[1] = cls; p.out[2] = dfl;
  p.out[3] = __fadd_rn(__fadd_rn(box, cls), dfl);
  p.out[4] = tss;
  p.out[5] = (float)s[4];
  p.out[6] = 0.0f; p.out[7] = 0.0f;
}

struct Layout {
  int64_t pbox, sig, aword, cand_a, cand_ov, cand_al, part1, img_part, total;
};
Layout layout(int B, int A, int nc, int G) {
  auto up = [](int64_t v) { return (v + 255) & ~(int64_t)255; };
  Layout L;
  int64_t o = 0;
  const int nblk1 = (A + 255) / 256;
  L.pbox = o; o = up(o + (int64_t)B * A * 16);
  L.sig = o; o = up(o + (int64_t)B * nc * A * 4);
  L.aword = o; o = up(o + (int64_t)B * A * 4);
  L.cand_a = o; o = up(o + (int64_t)B * G * kTopK * 4);
  L.cand_ov = o; o = up(o + (int64_t)B * G * kTopK * 4);
  L.cand_al = o; o = up(o + (int64_t)B * G * kTopK * 4);
  L.part1 = o; o = up(o + (int64_t)B * nblk1 * 8);
  L.img_part = o; o = up(o + (int64_t)B * 5 * 8);
  L.total = o;
  return L;
}

}  // namespace
}  // namespace dcfa

extern "C" int64_t dcfa_loss_workspace_bytes(int B, int A, int nc, int G) {
  if (B <= 0 || A <= 0 || nc <= 0 || G < 0) return -1;
  return dcfa::layout(B, A, nc, G).total;
}

extern "C" int dcfa_yolo_loss(const float* x0, const float* x1, const float* x2, int B, int nc, const int32_t* level_hw,
                              const float* level_stride, const float* gt, int G, float* out, void* workspace,
                              int64_t workspace_bytes, void* stream) {
  using namespace dcfa;
  DCFA_REQUIRE(x0 && x1 && x2 && level_hw && level_stride && out && workspace, "yolo_loss: null pointer");
  DCFA_REQUIRE(B > 0 && nc > 0 && G >= 0, "yolo_loss: bad sizes B=%d nc=%d G=%d", B, nc, G);
  DCFA_REQUIRE(G == 0 || gt, "yolo_loss: %d ground-truth rows per image but no gt tensor", G);
  DCFA_REQUIRE(B <= 65535, "yolo_loss: B=%d beyond the grid limit of 65535 images", B);
  LossArgs a;
  a.map[0] = x0; a.map[1] = x1; a.map[2] = x2;
  a.gt = gt; a.out = out;
  a.B = B; a.nc = nc; a.no = 4 * kBins + nc; a.G = G;
  a.A = 0;
  for (int l = 0; l < 3; ++l) {
    DCFA_REQUIRE(level_hw[2 * l] > 0 && level_hw[2 * l + 1] > 0, "yolo_loss: empty level %d", l);
    a.hw[l] = level_hw[2 * l] * level_hw[2 * l + 1];
    a.w[l] = level_hw[2 * l + 1];
    a.stride[l] = level_stride[l];
    a.A += a.hw[l];
  }
  DCFA_REQUIRE(a.A >= kTopK, "yolo_loss: %d anchors, fewer than topk = %d (torch.topk raises in the reference)", a.A, kTopK);
  a.nblk1 = (a.A + 255) / 256;
  const Layout L = layout(B, a.A, nc, G);
  DCFA_REQUIRE(workspace_bytes >= L.total, "yolo_loss: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)L.total);
  char* ws = static_cast<char*>(workspace);
  a.pbox = reinterpret_cast<float4*>(ws + L.pbox);
  a.sig = reinterpret_cast<float*>(ws + L.sig);
  a.aword = reinterpret_cast<unsigned*>(ws + L.aword);
  a.cand_a = reinterpret_cast<int*>(ws + L.cand_a);
  a.cand_ov = reinterpret_cast<float*>(ws + L.cand_ov);
  a.cand_al = reinterpret_cast<float*>(ws + L.cand_al);
  a.part1 = reinterpret_cast<double*>(ws + L.part1);
  a.img_part = reinterpret_cast<double*>(ws + L.img_part);
  cudaStream_t st = (cudaStream_t)stream;

  loss_decode_kernel<<<dim3((unsigned)a.nblk1, (unsigned)B), 256, 0, st>>>(a);
  DCFA_CHECK_LAUNCH("loss_decode_kernel");
  if (G > 0) {
    const size_t smem2 = (size_t)a.A * sizeof(float);
    DCFA_REQUIRE(smem2 <= 220 * 1024, "yolo_loss: %d anchors do not fit the assigner's shared-memory metric array", a.A);
    static DeviceOnce once2;
    static size_t smem2_set[64] = {};
    const int dev = current_device();
    if (once2.needed() || (dev >= 0 && dev < 64 && smem2_set[dev] < smem2)) {
      cudaError_t e = cudaFuncSetAttribute(loss_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) return fail(DCFA_E_CUDA, "yolo_loss: shared-memory opt-in failed: %s", cudaGetErrorString(e));
      e = cudaFuncSetAttribute(loss_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) return fail(DCFA_E_CUDA, "yolo_loss: shared-memory opt-in failed: %s", cudaGetErrorString(e));
      if (dev >= 0 && dev < 64) smem2_set[dev] = 220 * 1024;
      once2.mark();
    }
    loss_assign_kernel<<<dim3((unsigned)G, (unsigned)B), 256, smem2, st>>>(a);
    DCFA_CHECK_LAUNCH("loss_assign_kernel");
    const size_t smem3 = (size_t)G * kTopK * 12 + (size_t)G * 8;
    DCFA_REQUIRE(smem3 <= 220 * 1024, "yolo_loss: %d boxes per image exceed the resolve kernel's shared memory", G);
    loss_resolve_kernel<<<dim3((unsigned)B), kResolveThreads, smem3, st>>>(a);
    DCFA_CHECK_LAUNCH("loss_resolve_kernel");
  }
  loss_final_kernel<<<1, 256, 0, st>>>(a);
  DCFA_CHECK_LAUNCH("loss_final_kernel");
  return DCFA_OK;
}
