// DCFA_OP_DFL and dcfa_decode_box: the anchor-free box head after the last convolution.
//   DFL        nets/yolo_mul.py:312-322 (softmax over 16 bins, expectation with weights 0..15) together with
//              the view/cat/split of :459-460 that gathers the three NCHW level maps into (B, 64, A) / (B, nc, A).
//   decode_box utils/utils_bbox.py:30-40 (dist2bbox, xywh) and :49-58 (x strides, sigmoid, / input size).
// Single pass, one thread per (image, anchor); every access is coalesced along the anchor axis.  fp32
// throughout, exponentials with expf (not the fast intrinsic): the spec for this stage is 1e-5.
#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {
namespace {

struct DflArgs {
  const float* map[3];   // [B, no, H_l, W_l] fp32 NCHW
  float* dbox;           // [B, 4, A]
  float* cls;            // [B, nc, A]
  int B, nc, no, A;
  int hw[3];             // H_l * W_l
};

__global__ void __launch_bounds__(256) dfl_kernel(const DflArgs p) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)p.B * p.A) return;
  const int b = (int)(i / p.A);
  const int a = (int)(i - (int64_t)b * p.A);
  int l = 0, pos = a;
  if (pos >= p.hw[0]) { pos -= p.hw[0]; l = 1; }
  if (l == 1 && pos >= p.hw[1]) { pos -= p.hw[1]; l = 2; }
  const int hw = p.hw[l];
  const float* src = p.map[l] + (int64_t)b * p.no * hw + pos;
#pragma unroll
  for (int side = 0; side < 4; ++side) {
    float v[16];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      v[k] = __ldg(src + (int64_t)(side * 16 + k) * hw);
      mx = fmaxf(mx, v[k]);
    }
    float den = 0.0f, num = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float e = expf(v[k] - mx);
      den += e;
      num = fmaf((float)k, e, num);
    }
    p.dbox[((int64_t)b * 4 + side) * p.A + a] = num / den;
  }
  for (int c = 0; c < p.nc; ++c) p.cls[((int64_t)b * p.nc + c) * p.A + a] = __ldg(src + (int64_t)(64 + c) * hw);
}

struct DecArgs {
  const float* dbox;
  const float* cls;
  const float* anchors;
  const float* strides;
  float* out;
  int64_t cls_bstride, anc_s0, anc_s1;
  int B, A, nc;
  float img_w, img_h;
};

__global__ void __launch_bounds__(256) decode_kernel(const DecArgs p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)p.B * p.A) return;
  const int b = (int)(i / p.A);
  const int a = (int)(i - (int64_t)b * p.A);
  const float* d = p.dbox + (int64_t)b * 4 * p.A + a;
  const float lt_x = d[0], lt_y = d[p.A], rb_x = d[2 * (int64_t)p.A], rb_y = d[3 * (int64_t)p.A];
  const float ax = __ldg(p.anchors + a * p.anc_s1), ay = __ldg(p.anchors + p.anc_s0 + a * p.anc_s1);
  const float s = __ldg(p.strides + a);
  // same operation order as the reference; explicit _rn intrinsics keep nvcc from contracting into FMAs
  const float x1 = __fsub_rn(ax, lt_x), y1 = __fsub_rn(ay, lt_y);
  const float x2 = __fadd_rn(ax, rb_x), y2 = __fadd_rn(ay, rb_y);
  const float cx = __fdiv_rn(__fadd_rn(x1, x2), 2.0f), cy = __fdiv_rn(__fadd_rn(y1, y2), 2.0f);
  const float w = __fsub_rn(x2, x1), h = __fsub_rn(y2, y1);
  float* o = p.out + i * (4 + p.nc);
  o[0] = __fdiv_rn(__fmul_rn(cx, s), p.img_w);
  o[1] = __fdiv_rn(__fmul_rn(cy, s), p.img_h);
  o[2] = __fdiv_rn(__fmul_rn(w, s), p.img_w);
  o[3] = __fdiv_rn(__fmul_rn(h, s), p.img_h);
  const float* c = p.cls + (int64_t)b * p.cls_bstride + a;
  for (int k = 0; k < p.nc; ++k) o[4 + k] = 1.0f / (1.0f + expf(-c[(int64_t)k * p.A]));
}

}  // namespace

int launch_dfl(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  DflArgs a;
  a.map[0] = resolve_ptr<const float>(op.a0, bufs);
  a.map[1] = resolve_ptr<const float>(op.a1, bufs);
  a.map[2] = resolve_ptr<const float>(op.a2, bufs);
  a.dbox = resolve_ptr<float>(op.y, bufs);
  a.cls = resolve_ptr<float>(op.x2, bufs);
  a.B = op.n_img; a.nc = op.nc; a.no = 64 + op.nc; a.A = op.A;
  int H = op.Hi, W = op.Wi, tot = 0;
  for (int l = 0; l < 3; ++l) {
    a.hw[l] = H * W;
    tot += a.hw[l];
    H = (H + 1) / 2;
    W = (W + 1) / 2;
  }
  DCFA_REQUIRE(a.map[0] && a.map[1] && a.map[2] && a.dbox && a.cls, "dfl: missing tensor");
  DCFA_REQUIRE(tot == a.A, "dfl: anchors %d != level sizes %d", a.A, tot);
  DCFA_REQUIRE(a.nc >= 1, "dfl: nc must be >= 1");
  const int64_t total = (int64_t)a.B * a.A;
  launch_pdl(dfl_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, a);
  DCFA_CHECK_LAUNCH("dfl_kernel");
  return DCFA_OK;
}

}  // namespace dcfa

extern "C" int dcfa_decode_box(const float* dbox, const float* cls, int64_t cls_bstride, const float* anchors,
                               int64_t anc_s0, int64_t anc_s1, const float* strides, int B, int A, int nc,
                               float img_w, float img_h, float* out, void* stream) {
  using namespace dcfa;
  DCFA_REQUIRE(dbox && cls && anchors && strides && out, "decode_box: null pointer");
  DCFA_REQUIRE(B > 0 && A > 0 && nc > 0, "decode_box: bad sizes B=%d A=%d nc=%d", B, A, nc);
  DecArgs a;
  a.dbox = dbox; a.cls = cls; a.anchors = anchors; a.strides = strides; a.out = out;
  a.cls_bstride = cls_bstride; a.anc_s0 = anc_s0; a.anc_s1 = anc_s1;
  a.B = B; a.A = A; a.nc = nc; a.img_w = img_w; a.img_h = img_h;
  const int64_t total = (int64_t)B * A;
  decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  DCFA_CHECK_LAUNCH("decode_kernel");
  return DCFA_OK;
}
