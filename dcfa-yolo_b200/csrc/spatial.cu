// DCFA_OP_MAXPOOL5: MaxPool2d(5, stride 1, pad 2) of SPPF_CBAM (nets/yolo_mul.py:17, :26-30); out-of-image
//   positions are excluded (the reference pads with -inf; CBAM outputs can be negative).
// DCFA_OP_UPSAMPLE: F.interpolate(mode='bilinear', align_corners=True) of the FPN top-down path
//   (nets/yolo_mul.py:426, :433), optionally summing two inputs first (feat3_rgb + feat3_nir, :421) and
//   writing straight into a channel slot of the BiFPN concat buffer (:428, :435).
// Both are bandwidth-bound gathers on bf16 NHWC with 128-bit channel vectors.
#include "common.cuh"

namespace dcfa {
namespace {

struct PoolArgs5 {
  View<const __nv_bfloat16> x;
  View<__nv_bfloat16> y;
  int n_img, H, W, C;
  int64_t total;
};

__global__ void __launch_bounds__(256) maxpool5_kernel(const PoolArgs5 p) {
  const int c8n = p.C >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int c8 = (int)(t % c8n); t /= c8n;
    const int x = (int)(t % p.W); t /= p.W;
    const int y = (int)(t % p.H);
    const int n = (int)(t / p.H);
    const __nv_bfloat16* xin = p.x.p + p.x.img_off(n) + c8 * 8;
    const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
    __nv_bfloat162 m[4] = {ninf, ninf, ninf, ninf};
    const int y_lo = max(y - 2, 0), y_hi = min(y + 2, p.H - 1);
    const int x_lo = max(x - 2, 0), x_hi = min(x + 2, p.W - 1);
    for (int iy = y_lo; iy <= y_hi; ++iy)
      for (int ix = x_lo; ix <= x_hi; ++ix) {
        const uint4 v = ldg128(xin + (int64_t)(iy * p.W + ix) * p.x.ld);
        m[0] = __hmax2(m[0], *reinterpret_cast<const __nv_bfloat162*>(&v.x));
        m[1] = __hmax2(m[1], *reinterpret_cast<const __nv_bfloat162*>(&v.y));
        m[2] = __hmax2(m[2], *reinterpret_cast<const __nv_bfloat162*>(&v.z));
        m[3] = __hmax2(m[3], *reinterpret_cast<const __nv_bfloat162*>(&v.w));
      }
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&m[0]);
    o.y = *reinterpret_cast<uint32_t*>(&m[1]);
    o.z = *reinterpret_cast<uint32_t*>(&m[2]);
    o.w = *reinterpret_cast<uint32_t*>(&m[3]);
    stg128(p.y.p + p.y.img_off(n) + (int64_t)(y * p.W + x) * p.y.ld + c8 * 8, o);
  }
}

struct UpArgs {
  View<const __nv_bfloat16> a;
  View<const __nv_bfloat16> b;  // optional second addend
  View<__nv_bfloat16> y;
  int n_img, Hi, Wi, Ho, Wo, C;
  float sy, sx;  // (in-1)/(out-1)
  int64_t total;
};

__device__ __forceinline__ void load_sum8(const UpArgs& p, int64_t offa, int64_t offb, float* v) {
  unpack8(ldg128(p.a.p + offa), v);
  if (p.b.p) {
    float w[8];
    unpack8(ldg128(p.b.p + offb), w);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += w[e];
  }
}

__global__ void __launch_bounds__(256) upsample_kernel(const UpArgs p) {
  const int c8n = p.C >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int c8 = (int)(t % c8n); t /= c8n;
    const int ox = (int)(t % p.Wo); t /= p.Wo;
    const int oy = (int)(t % p.Ho);
    const int n = (int)(t / p.Ho);
    // ATen upsample_bilinear2d, align_corners=True: src = scale * dst, scale = (in-1)/(out-1)
    const float fy = p.sy * (float)oy, fx = p.sx * (float)ox;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < p.Hi - 1 ? 1 : 0), x1 = x0 + (x0 < p.Wi - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    const int64_t ba = p.a.img_off(n) + c8 * 8;
    const int64_t bb = p.b.p ? p.b.img_off(n) + c8 * 8 : 0;
    float v00[8], v01[8], v10[8], v11[8];
    load_sum8(p, ba + (int64_t)(y0 * p.Wi + x0) * p.a.ld, bb + (int64_t)(y0 * p.Wi + x0) * p.b.ld, v00);
    load_sum8(p, ba + (int64_t)(y0 * p.Wi + x1) * p.a.ld, bb + (int64_t)(y0 * p.Wi + x1) * p.b.ld, v01);
    load_sum8(p, ba + (int64_t)(y1 * p.Wi + x0) * p.a.ld, bb + (int64_t)(y1 * p.Wi + x0) * p.b.ld, v10);
    load_sum8(p, ba + (int64_t)(y1 * p.Wi + x1) * p.a.ld, bb + (int64_t)(y1 * p.Wi + x1) * p.b.ld, v11);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = hy * (hx * v00[e] + lx * v01[e]) + ly * (hx * v10[e] + lx * v11[e]);
    stg128(p.y.p + p.y.img_off(n) + (int64_t)(oy * p.Wo + ox) * p.y.ld + c8 * 8, pack8(o));
  }
}

inline bool aligned(const void* ptr, int ld, int64_t img_stride, int64_t gstride) {
  return ((uintptr_t)ptr % 16) == 0 && ld % 8 == 0 && img_stride % 8 == 0 && gstride % 8 == 0;
}

}  // namespace

int launch_maxpool5(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  PoolArgs5 a;
  a.x = resolve<const __nv_bfloat16>(op.x, bufs);
  a.y = resolve<__nv_bfloat16>(op.y, bufs);
  a.n_img = op.n_img; a.H = op.Hi; a.W = op.Wi; a.C = op.Cin;
  DCFA_REQUIRE(a.x.p && a.y.p, "maxpool5: missing tensor");
  DCFA_REQUIRE(a.C > 0 && a.C % 8 == 0, "maxpool5: C %d unsupported", a.C);
  DCFA_REQUIRE(aligned(a.x.p, a.x.ld, a.x.img_stride, a.x.gstride) && aligned(a.y.p, a.y.ld, a.y.img_stride, a.y.gstride),
               "maxpool5: views must be 16-byte aligned");
  a.total = (int64_t)a.n_img * a.H * a.W * (a.C >> 3);
  int64_t blocks = (a.total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  maxpool5_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  DCFA_CHECK_LAUNCH("maxpool5_kernel");
  return DCFA_OK;
}

int launch_upsample(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  UpArgs a;
  a.a = resolve<const __nv_bfloat16>(op.x, bufs);
  a.b = resolve<const __nv_bfloat16>(op.x2, bufs);
  a.y = resolve<__nv_bfloat16>(op.y, bufs);
  a.n_img = op.n_img; a.Hi = op.Hi; a.Wi = op.Wi; a.Ho = op.Ho; a.Wo = op.Wo; a.C = op.Cin;
  DCFA_REQUIRE(a.a.p && a.y.p, "upsample: missing tensor");
  DCFA_REQUIRE(a.C > 0 && a.C % 8 == 0, "upsample: C %d unsupported", a.C);
  DCFA_REQUIRE(a.Hi > 0 && a.Wi > 0 && a.Ho > 0 && a.Wo > 0, "upsample: bad sizes");
  DCFA_REQUIRE(aligned(a.a.p, a.a.ld, a.a.img_stride, a.a.gstride) && aligned(a.y.p, a.y.ld, a.y.img_stride, a.y.gstride) &&
                   (!a.b.p || aligned(a.b.p, a.b.ld, a.b.img_stride, a.b.gstride)),
               "upsample: views must be 16-byte aligned");
  a.sy = a.Ho > 1 ? (float)(a.Hi - 1) / (float)(a.Ho - 1) : 0.0f;
  a.sx = a.Wo > 1 ? (float)(a.Wi - 1) / (float)(a.Wo - 1) : 0.0f;
  a.total = (int64_t)a.n_img * a.Ho * a.Wo * (a.C >> 3);
  int64_t blocks = (a.total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  upsample_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  DCFA_CHECK_LAUNCH("upsample_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
