// DCFA_OP_MAXPOOL5: MaxPool2d(5, stride 1, pad 2) of SPPF_CBAM (nets/yolo_mul.py:17, :26-30); out-of-image
//   positions are excluded (the reference pads with -inf; CBAM outputs can be negative).
// DCFA_OP_UPSAMPLE: F.interpolate(mode='bilinear', align_corners=True) of the FPN top-down path
//   (nets/yolo_mul.py:426, :433), optionally summing two inputs first (feat3_rgb + feat3_nir, :421) and
//   writing straight into a channel slot of the BiFPN concat buffer (:428, :435).
// Both are bandwidth-bound gathers on bf16 NHWC with 128-bit channel vectors.
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {
namespace {

struct PoolArgs5 {
  View<const __nv_bfloat16> x;
  View<__nv_bfloat16> y;
  int n_img, H, W, C;
  int strip;  // output rows per thread
};

struct Max8 {
  __nv_bfloat162 m[4];
  __device__ __forceinline__ void fill(float v) { m[0] = m[1] = m[2] = m[3] = __float2bfloat162_rn(v); }
  __device__ __forceinline__ void take(const uint4& v) {
    m[0] = __hmax2(m[0], *reinterpret_cast<const __nv_bfloat162*>(&v.x));
    m[1] = __hmax2(m[1], *reinterpret_cast<const __nv_bfloat162*>(&v.y));
    m[2] = __hmax2(m[2], *reinterpret_cast<const __nv_bfloat162*>(&v.z));
    m[3] = __hmax2(m[3], *reinterpret_cast<const __nv_bfloat162*>(&v.w));
  }
  __device__ __forceinline__ void take(const Max8& o) {
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = __hmax2(m[i], o.m[i]);
  }
};

// Separable 5x5 max: a thread owns one (column, 8-channel chunk) and walks down a strip of rows keeping the
// horizontal 5-max of the last five input rows in registers -- 5 loads per input row instead of 25 per output.
// grid = (ceil(W * C/8 / 256), strips, images)
__global__ void __launch_bounds__(256) maxpool5_kernel(const PoolArgs5 p) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int c8n = p.C >> 3;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= p.W * c8n) return;
  const int x = idx / c8n, c8 = idx - x * c8n;
  const int n = blockIdx.z;
  const int ys = blockIdx.y * p.strip, ye = min(p.H, ys + p.strip);
  const __nv_bfloat16* xin = p.x.p + p.x.img_off(n) + c8 * 8;
  __nv_bfloat16* yout = p.y.p + p.y.img_off(n) + c8 * 8;
  // the five column taps, clamped into the image (a duplicate tap does not change a max)
  int64_t xo[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) xo[k] = (int64_t)min(max(x + k - 2, 0), p.W - 1) * p.x.ld;
  Max8 w[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) w[i].fill(-INFINITY);
  // rows are software-pipelined: the loads of row iy + 1 are in flight while row iy is reduced
  const int r_lo = max(ys - 2, 0), r_hi = min(ye + 2, p.H);   // input rows [r_lo, r_hi)
  uint4 nxt[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) nxt[k] = ldg128(xin + (int64_t)r_lo * p.W * p.x.ld + xo[k]);
  for (int iy = ys - 2; iy < ye + 2; ++iy) {
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = w[i + 1];
    w[4].fill(-INFINITY);
    if (iy >= r_lo && iy < r_hi) {
      uint4 cur[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) cur[k] = nxt[k];
      if (iy + 1 < r_hi) {
        const __nv_bfloat16* row = xin + (int64_t)(iy + 1) * p.W * p.x.ld;
#pragma unroll
        for (int k = 0; k < 5; ++k) nxt[k] = ldg128(row + xo[k]);
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) w[4].take(cur[k]);
    }
    const int oy = iy - 2;
    if (oy >= ys) {
      Max8 o = w[0];
#pragma unroll
      for (int i = 1; i < 5; ++i) o.take(w[i]);
      uint4 v;
      v.x = *reinterpret_cast<uint32_t*>(&o.m[0]);
      v.y = *reinterpret_cast<uint32_t*>(&o.m[1]);
      v.z = *reinterpret_cast<uint32_t*>(&o.m[2]);
      v.w = *reinterpret_cast<uint32_t*>(&o.m[3]);
      stg128(yout + (int64_t)(oy * p.W + x) * p.y.ld, v);
    }
  }
}

struct UpArgs {
  View<const __nv_bfloat16> a;
  View<const __nv_bfloat16> b;  // optional second addend
  View<__nv_bfloat16> y;
  int n_img, Hi, Wi, Ho, Wo, C;
  float sy, sx;  // (in-1)/(out-1)
};

// 8 bf16 channels of one source pixel (plus the second tensor's, if any) as four packed fp32 pairs
template <bool TWO>
__device__ __forceinline__ void load_sum8(const __nv_bfloat16* a, const __nv_bfloat16* b, F2* v) {
  unpack8_f2(ldg128(a), v);
  if (TWO) {
    F2 w[4];
    unpack8_f2(ldg128(b), w);
    const F2 one = f2_make(1.0f, 1.0f);
#pragma unroll
    for (int e = 0; e < 4; ++e) f2_fma(v[e], w[e], one);   // v + w (x * 1 + v is exact)
  }
}

// grid = (ceil(Wo * C/8 / 256), ceil(Ho / kUpRows), images): a thread produces kUpRows vertically adjacent
// outputs of one (column, 8-channel chunk), so the column terms (one 32-bit division) are paid once and
// 4 * kUpRows loads are in flight.
constexpr int kUpRows = 4;

template <bool TWO>
__global__ void __launch_bounds__(256) upsample_kernel(const UpArgs p) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int c8n = p.C >> 3;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= p.Wo * c8n) return;
  const int ox = idx / c8n, c8 = idx - ox * c8n;
  const int n = blockIdx.z;
  // ATen upsample_bilinear2d, align_corners=True: src = scale * dst, scale = (in-1)/(out-1)
  const float fx = p.sx * (float)ox;
  const int x0 = (int)fx;
  const int x1 = x0 + (x0 < p.Wi - 1 ? 1 : 0);
  const float lx = fx - (float)x0, hx = 1.0f - lx;
  const __nv_bfloat16* pa = p.a.p + p.a.img_off(n) + c8 * 8;
  const __nv_bfloat16* pb = TWO ? p.b.p + p.b.img_off(n) + c8 * 8 : nullptr;
  __nv_bfloat16* py = p.y.p + p.y.img_off(n) + (int64_t)ox * p.y.ld + c8 * 8;
#pragma unroll
  for (int r = 0; r < kUpRows; ++r) {
    const int oy = blockIdx.y * kUpRows + r;
    if (oy < p.Ho) {
      const float fy = p.sy * (float)oy;
      const int y0 = (int)fy;
      const int y1 = y0 + (y0 < p.Hi - 1 ? 1 : 0);
      const float ly = fy - (float)y0, hy = 1.0f - ly;
      const int i00 = y0 * p.Wi + x0, i01 = y0 * p.Wi + x1, i10 = y1 * p.Wi + x0, i11 = y1 * p.Wi + x1;
      // packed fp32x2 arithmetic (fma.rn.f32x2): the kernel is issue-bound, this halves its FMA count
      F2 v00[4], v01[4], v10[4], v11[4];
      load_sum8<TWO>(pa + (int64_t)i00 * p.a.ld, pb + (int64_t)i00 * p.b.ld, v00);
      load_sum8<TWO>(pa + (int64_t)i01 * p.a.ld, pb + (int64_t)i01 * p.b.ld, v01);
      load_sum8<TWO>(pa + (int64_t)i10 * p.a.ld, pb + (int64_t)i10 * p.b.ld, v10);
      load_sum8<TWO>(pa + (int64_t)i11 * p.a.ld, pb + (int64_t)i11 * p.b.ld, v11);
      const F2 hx2 = f2_make(hx, hx), lx2 = f2_make(lx, lx), hy2 = f2_make(hy, hy), ly2 = f2_make(ly, ly);
      const F2 zero = f2_make(0.0f, 0.0f);
      float o[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        F2 t0 = zero, t1 = zero, r = zero;
        f2_fma(t0, hx2, v00[e]);      // hx * v00
        f2_fma(t0, lx2, v01[e]);      //  + lx * v01
        f2_fma(t1, hx2, v10[e]);
        f2_fma(t1, lx2, v11[e]);
        f2_fma(r, ly2, t1);           // ly * bottom
        f2_fma(r, hy2, t0);           //  + hy * top
        f2_get(r, o[2 * e], o[2 * e + 1]);
      }
      stg128(py + (int64_t)oy * p.Wo * p.y.ld, pack8(o));
    }
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
  DCFA_REQUIRE(a.n_img <= 65535, "maxpool5: n_img %d too large", a.n_img);
  // strips tall enough to amortise the 4 halo rows, short enough to fill the SMs
  const int bx = ceil_div(a.W * (a.C >> 3), 256);
  int strips = ceil_div(sm_count() * 4, bx * a.n_img);
  strips = std::max(1, std::min(strips, ceil_div(a.H, 4)));
  a.strip = ceil_div(a.H, strips);
  strips = ceil_div(a.H, a.strip);
  launch_pdl(maxpool5_kernel, dim3((unsigned)bx, (unsigned)strips, (unsigned)a.n_img), dim3(256), 0, st, a);
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
  DCFA_REQUIRE(a.n_img <= 65535 && a.Ho <= 65535, "upsample: grid too large");
  const dim3 grid((unsigned)ceil_div(a.Wo * (a.C >> 3), 256), (unsigned)ceil_div(a.Ho, kUpRows), (unsigned)a.n_img);
  if (a.b.p) launch_pdl(upsample_kernel<true>, grid, dim3(256), 0, st, a);
  else launch_pdl(upsample_kernel<false>, grid, dim3(256), 0, st, a);
  DCFA_CHECK_LAUNCH("upsample_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
