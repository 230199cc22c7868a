// DCFA_OP_DWCONV: depthwise 3x3 stride 1 pad 1 on bf16 NHWC with folded BatchNorm.
//   ShuffleNetV2 branch2 depthwise (bias=True) + BN           nets/yolo_mul.py:144-146
//   RepGhostModule cheap_operation + fusion_bn identity branch nets/repghost.py:98-115 (the identity-BN
//   is folded into the centre tap exactly as switch_to_deploy does, :117-123), optional SiLU (:105-106,115)
//   and the bottleneck's identity shortcut (:279).
// Memory-bound: each thread owns 8 channels (one 128-bit word) of 4 horizontally adjacent output pixels,
// so the 72 fp32 weights it needs are loaded once and the 3x6 input window is read with 18 vector loads.
#include "common.cuh"

namespace dcfa {
namespace {

constexpr int XT = 4;

struct DwArgs {
  View<const __nv_bfloat16> x;
  View<const __nv_bfloat16> res;
  View<__nv_bfloat16> y;
  const float* w;     // [G][9][C]
  const float* bias;  // [G][C]
  int n_img, group_imgs, H, W, C, act;
  int xg;             // ceil(W / XT)
  int64_t total;
};

__global__ void __launch_bounds__(256) dwconv3x3_kernel(const DwArgs p) {
  const int c8n = p.C >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int c8 = (int)(t % c8n); t /= c8n;
    const int xg = (int)(t % p.xg); t /= p.xg;
    const int y = (int)(t % p.H);
    const int n = (int)(t / p.H);
    const int g = n / p.group_imgs;
    const int c = c8 * 8;
    const int x0 = xg * XT;

    const float* wp = p.w + ((int64_t)g * 9) * p.C + c;
    const float* bp = p.bias + (int64_t)g * p.C + c;
    float acc[XT][8];
    {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bp) + 1);
#pragma unroll
      for (int j = 0; j < XT; ++j) {
        acc[j][0] = b0.x; acc[j][1] = b0.y; acc[j][2] = b0.z; acc[j][3] = b0.w;
        acc[j][4] = b1.x; acc[j][5] = b1.y; acc[j][6] = b1.z; acc[j][7] = b1.w;
      }
    }
    const __nv_bfloat16* xin = p.x.p + p.x.img_off(n) + c;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = y + r - 1;
      if (iy < 0 || iy >= p.H) continue;
      float wr[3][8];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)(r * 3 + q) * p.C));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)(r * 3 + q) * p.C) + 1);
        wr[q][0] = w0.x; wr[q][1] = w0.y; wr[q][2] = w0.z; wr[q][3] = w0.w;
        wr[q][4] = w1.x; wr[q][5] = w1.y; wr[q][6] = w1.z; wr[q][7] = w1.w;
      }
#pragma unroll
      for (int q = 0; q < XT + 2; ++q) {
        const int ix = x0 + q - 1;
        if (ix < 0 || ix >= p.W) continue;
        float v[8];
        unpack8(ldg128(xin + (int64_t)(iy * p.W + ix) * p.x.ld), v);
#pragma unroll
        for (int j = 0; j < XT; ++j) {
          const int tap = q - j;  // input column q feeds output j through kernel column q - j
          if (tap >= 0 && tap < 3) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(v[e], wr[tap][e], acc[j][e]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < XT; ++j) {
      const int ox = x0 + j;
      if (ox >= p.W) continue;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = apply_act(acc[j][e], p.act);
      const int64_t pix = (int64_t)y * p.W + ox;
      if (p.res.p) {
        float rr[8];
        unpack8(ldg128(p.res.p + p.res.img_off(n) + pix * p.res.ld + c), rr);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] += rr[e];
      }
      stg128(p.y.p + p.y.img_off(n) + pix * p.y.ld + c, pack8(o));
    }
  }
}

}  // namespace

int launch_dwconv(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  DwArgs a;
  a.x = resolve<const __nv_bfloat16>(op.x, bufs);
  a.res = resolve<const __nv_bfloat16>(op.x2, bufs);
  a.y = resolve<__nv_bfloat16>(op.y, bufs);
  a.w = resolve_ptr<const float>(op.w, bufs);
  a.bias = resolve_ptr<const float>(op.bias, bufs);
  a.n_img = op.n_img;
  a.group_imgs = op.group_imgs > 0 ? op.group_imgs : op.n_img;
  a.H = op.Hi; a.W = op.Wi; a.C = op.Cin; a.act = op.act;
  DCFA_REQUIRE(a.x.p && a.y.p && a.w && a.bias, "dwconv: missing tensor");
  DCFA_REQUIRE(a.C > 0 && a.C % 8 == 0, "dwconv: C %d must be a multiple of 8", a.C);
  DCFA_REQUIRE(a.n_img > 0 && a.n_img % a.group_imgs == 0, "dwconv: bad grouping");
  DCFA_REQUIRE(((uintptr_t)a.x.p % 16) == 0 && a.x.ld % 8 == 0 && a.x.img_stride % 8 == 0 && a.x.gstride % 8 == 0 &&
                   ((uintptr_t)a.y.p % 16) == 0 && a.y.ld % 8 == 0 && a.y.img_stride % 8 == 0 && a.y.gstride % 8 == 0,
               "dwconv: views must be 16-byte aligned");
  if (a.res.p)
    DCFA_REQUIRE(((uintptr_t)a.res.p % 16) == 0 && a.res.ld % 8 == 0 && a.res.img_stride % 8 == 0 && a.res.gstride % 8 == 0,
                 "dwconv: residual view must be 16-byte aligned");
  DCFA_REQUIRE(((uintptr_t)a.w % 16) == 0 && ((uintptr_t)a.bias % 16) == 0, "dwconv: params must be 16-byte aligned");
  a.xg = ceil_div(a.W, XT);
  a.total = (int64_t)a.n_img * a.H * a.xg * (a.C >> 3);
  int64_t blocks = (a.total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  dwconv3x3_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  DCFA_CHECK_LAUNCH("dwconv3x3_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
