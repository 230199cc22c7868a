// DCFA_OP_STEM: Conv_maxpool (nets/yolo_mul.py:104-115) fused into one pass:
//   fp32 NCHW image -> conv3x3 s1 p1 (3 -> C0, BN folded into the weights) -> ReLU -> maxpool 3x3 s2 p1
//   -> bf16 NHWC.  The full-resolution conv map (the largest tensor of the whole network) only ever
//   exists as a shared-memory tile.  Both modalities run in one launch (weight group = image / group_imgs).
//
// CTA tile: 4 x 32 pooled pixels = 9 x 65 conv pixels = 11 x 67 x 3 input patch.  Each thread computes a
// horizontal pair of conv pixels for 8 output channels at a time (fp32 FMA, weights broadcast from smem).
// Post-ReLU values are >= 0 and every pool window holds a valid pixel, so out-of-image conv positions are
// written as 0 (equivalent to the reference's -inf padding).
#include "common.cuh"

namespace dcfa {
namespace {

constexpr int TPH = 4, TPW = 32;              // pooled tile
constexpr int CH = 2 * TPH + 1;               // 9 conv rows
constexpr int CW = 2 * TPW + 1;               // 65 conv cols
constexpr int PH = CH + 2, PW = CW + 2;       // 11 x 67 input patch
constexpr int PWP = 68;                       // padded patch row
constexpr int PAIRS = (CW + 1) / 2;           // 33 conv-pixel pairs per row
constexpr int kStemThreads = 320;

struct StemArgs {
  const float* x0;
  const float* x1;
  const float* w;     // [G][27][C0]
  const float* bias;  // [G][C0]
  View<__nv_bfloat16> y;
  int n_img, group_imgs, Hi, Wi, Ho, Wo, C0;
  int tiles_x, tiles_y;
};

__global__ void __launch_bounds__(kStemThreads) stem_kernel(const StemArgs p) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* s_in = reinterpret_cast<float*>(smem);                  // [3][PH][PWP]
  float* s_w = s_in + 3 * PH * PWP;                              // [27][C0]
  float* s_b = s_w + 27 * p.C0;                                  // [C0]
  __nv_bfloat16* s_conv = reinterpret_cast<__nv_bfloat16*>(s_b + p.C0);  // [CH*CW][C0]

  const int tid = threadIdx.x;
  int t = blockIdx.x;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y; t /= p.tiles_y;
  const int n = t;
  const int g = n / p.group_imgs;
  const int nl = n - g * p.group_imgs;
  const float* img = (g == 0 ? p.x0 : p.x1) + (int64_t)nl * 3 * p.Hi * p.Wi;

  const int py0 = ty * TPH, px0 = tx * TPW;
  const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;  // conv-map origin of the tile
  const int iy0 = cy0 - 1, ix0 = cx0 - 1;          // input origin of the patch

  for (int i = tid; i < 3 * PH * PW; i += kStemThreads) {
    const int c = i / (PH * PW);
    const int r = (i - c * PH * PW) / PW;
    const int q = i - c * PH * PW - r * PW;
    const int iy = iy0 + r, ix = ix0 + q;
    float v = 0.0f;
    if (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) v = __ldg(img + ((int64_t)c * p.Hi + iy) * p.Wi + ix);
    s_in[(c * PH + r) * PWP + q] = v;
  }
  for (int i = tid; i < 27 * p.C0; i += kStemThreads) s_w[i] = __ldg(p.w + (int64_t)g * 27 * p.C0 + i);
  for (int i = tid; i < p.C0; i += kStemThreads) s_b[i] = __ldg(p.bias + (int64_t)g * p.C0 + i);
  __syncthreads();

  if (tid < CH * PAIRS) {
    const int cy = tid / PAIRS;
    const int cxa = 2 * (tid - cy * PAIRS);  // conv col of pixel a inside the tile; pixel b = cxa + 1
    const bool b_in_tile = cxa + 1 < CW;
    float in[3][3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) in[c][r][q] = s_in[(c * PH + cy + r) * PWP + cxa + q];
    const int gy = cy0 + cy;
    const bool row_ok = gy >= 0 && gy < p.Hi;
    const bool a_ok = row_ok && (cx0 + cxa) >= 0 && (cx0 + cxa) < p.Wi;
    const bool b_ok = row_ok && b_in_tile && (cx0 + cxa + 1) >= 0 && (cx0 + cxa + 1) < p.Wi;
    __nv_bfloat16* dst_a = s_conv + (int64_t)(cy * CW + cxa) * p.C0;
    for (int co = 0; co < p.C0; co += 8) {
      float a[8], b[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = b[e] = s_b[co + e];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* wp = s_w + ((r * 3 + q) * 3 + c) * p.C0 + co;
            const float4 w0 = *reinterpret_cast<const float4*>(wp);
            const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
            const float xa = in[c][r][q], xb = in[c][r][q + 1];
            a[0] = fmaf(w0.x, xa, a[0]); a[1] = fmaf(w0.y, xa, a[1]); a[2] = fmaf(w0.z, xa, a[2]); a[3] = fmaf(w0.w, xa, a[3]);
            a[4] = fmaf(w1.x, xa, a[4]); a[5] = fmaf(w1.y, xa, a[5]); a[6] = fmaf(w1.z, xa, a[6]); a[7] = fmaf(w1.w, xa, a[7]);
            b[0] = fmaf(w0.x, xb, b[0]); b[1] = fmaf(w0.y, xb, b[1]); b[2] = fmaf(w0.z, xb, b[2]); b[3] = fmaf(w0.w, xb, b[3]);
            b[4] = fmaf(w1.x, xb, b[4]); b[5] = fmaf(w1.y, xb, b[5]); b[6] = fmaf(w1.z, xb, b[6]); b[7] = fmaf(w1.w, xb, b[7]);
          }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a[e] = a_ok ? fmaxf(a[e], 0.0f) : 0.0f;
        b[e] = b_ok ? fmaxf(b[e], 0.0f) : 0.0f;
      }
      *reinterpret_cast<uint4*>(dst_a + co) = pack8(a);
      if (b_in_tile) *reinterpret_cast<uint4*>(dst_a + p.C0 + co) = pack8(b);
    }
  }
  __syncthreads();

  const int c8n = p.C0 >> 3;
  for (int i = tid; i < TPH * TPW * c8n; i += kStemThreads) {
    const int c8 = i % c8n;
    const int pp = i / c8n;
    const int pxl = pp % TPW, pyl = pp / TPW;
    const int py = py0 + pyl, px = px0 + pxl;
    if (py >= p.Ho || px >= p.Wo) continue;
    __nv_bfloat162 m[4];
    {
      const uint4 v = *reinterpret_cast<const uint4*>(s_conv + (int64_t)((2 * pyl) * CW + 2 * pxl) * p.C0 + c8 * 8);
      m[0] = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
      m[1] = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
      m[2] = *reinterpret_cast<const __nv_bfloat162*>(&v.z);
      m[3] = *reinterpret_cast<const __nv_bfloat162*>(&v.w);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (r == 0 && q == 0) continue;
        const uint4 v =
            *reinterpret_cast<const uint4*>(s_conv + (int64_t)((2 * pyl + r) * CW + 2 * pxl + q) * p.C0 + c8 * 8);
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
    stg128(p.y.p + p.y.img_off(n) + (int64_t)(py * p.Wo + px) * p.y.ld + c8 * 8, o);
  }
}

}  // namespace

int launch_stem(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  StemArgs a;
  a.x0 = resolve_ptr<const float>(op.x, bufs);
  a.x1 = resolve_ptr<const float>(op.x2, bufs);
  a.w = resolve_ptr<const float>(op.w, bufs);
  a.bias = resolve_ptr<const float>(op.bias, bufs);
  a.y = resolve<__nv_bfloat16>(op.y, bufs);
  a.n_img = op.n_img;
  a.group_imgs = op.group_imgs > 0 ? op.group_imgs : op.n_img;
  a.Hi = op.Hi; a.Wi = op.Wi; a.Ho = op.Ho; a.Wo = op.Wo; a.C0 = op.Cout;
  DCFA_REQUIRE(a.x0 && a.w && a.bias && a.y.p, "stem: missing tensor");
  DCFA_REQUIRE(a.n_img == a.group_imgs || (a.n_img == 2 * a.group_imgs && a.x1), "stem: needs 1 or 2 groups");
  DCFA_REQUIRE(a.Hi > 0 && a.Wi > 0 && a.Ho == (a.Hi - 1) / 2 + 1 && a.Wo == (a.Wi - 1) / 2 + 1,
               "stem: pooled size %dx%d inconsistent with %dx%d", a.Ho, a.Wo, a.Hi, a.Wi);
  DCFA_REQUIRE(a.C0 % 8 == 0 && a.C0 >= 8 && a.C0 <= 128, "stem: C0 %d unsupported", a.C0);
  DCFA_REQUIRE(((uintptr_t)a.y.p % 16) == 0 && a.y.ld % 8 == 0 && a.y.img_stride % 8 == 0 && a.y.gstride % 8 == 0,
               "stem: output view must be 16-byte aligned");
  DCFA_REQUIRE(((uintptr_t)a.w % 16) == 0, "stem: weights must be 16-byte aligned");
  a.tiles_x = ceil_div(a.Wo, TPW);
  a.tiles_y = ceil_div(a.Ho, TPH);
  const int64_t blocks = (int64_t)a.n_img * a.tiles_x * a.tiles_y;
  DCFA_REQUIRE(blocks < (1ll << 31), "stem: grid too large");
  const size_t smem = (size_t)(3 * PH * PWP + 28 * a.C0) * sizeof(float) + (size_t)CH * CW * a.C0 * 2;
  static int max_set = 0;
  if ((int)smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(DCFA_E_CUDA, "stem: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    max_set = (int)smem;
  }
  stem_kernel<<<(unsigned)blocks, kStemThreads, smem, st>>>(a);
  DCFA_CHECK_LAUNCH("stem_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
