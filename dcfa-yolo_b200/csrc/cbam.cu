// CBAM attention (nets/yolo_mul.py:56-102) as four memory-bound kernels on bf16 NHWC:
//   CBAM_POOL   per-(image, channel) sum and max over HW, written as per-chunk partials   (:59-60,:70-71)
//   CBAM_MLP    fixed-order reduction of the partials, fc1 -> ReLU -> fc2 on the average and the max
//               vector, add, sigmoid -> channel gate                                      (:63-73)
//   CBAM_STATS  t = x * gate; per-pixel mean and max over channels                         (:100, :86-88)
//   CBAM_APPLY  7x7 conv over the 2-plane map (zero padded), sigmoid, y = x * gate * s     (:89-90, :101)
// Used at the six fusion sites (:346-353, :403-415) and inside SPPF_CBAM (:18-31, hidden width 1).
// All channel accesses are 128-bit (8 x bf16); reductions are warp shuffles / fixed-order loops, so the
// result is deterministic.
#include "common.cuh"

namespace dcfa {
namespace {

// ------------------------------------------------------------------------------------------ CBAM_POOL
struct PoolArgs {
  View<const __nv_bfloat16> x;
  float* psum;  // [n_img][parts][C]
  float* pmax;
  int n_img, HW, C, parts, pix_per_part;
};

__global__ void __launch_bounds__(256) cbam_pool_kernel(const PoolArgs p) {
  extern __shared__ float s_red[];  // [planes][2][C]
  const int c8n = p.C >> 3;          // <= 256 (C <= 2048)
  const int planes = 256 / c8n;      // pixel planes: threads with the same channel chunk
  const int n = blockIdx.x / p.parts;
  const int part = blockIdx.x - n * p.parts;
  const int p0 = part * p.pix_per_part;
  const int p1 = min(p.HW, p0 + p.pix_per_part);
  const __nv_bfloat16* xin = p.x.p + p.x.img_off(n);
  const int c8 = (int)threadIdx.x % c8n;
  const int plane = (int)threadIdx.x / c8n;
  if (plane < planes) {
    float s[8], m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.0f; m[e] = -INFINITY; }
    for (int px = p0 + plane; px < p1; px += planes) {
      float v[8];
      unpack8(ldg128(xin + (int64_t)px * p.x.ld + c8 * 8), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] += v[e]; m[e] = fmaxf(m[e], v[e]); }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s_red[(plane * 2 + 0) * p.C + c8 * 8 + e] = s[e];
      s_red[(plane * 2 + 1) * p.C + c8 * 8 + e] = m[e];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += 256) {
    float ss = 0.0f, mm = -INFINITY;
    for (int q = 0; q < planes; ++q) {
      ss += s_red[(q * 2 + 0) * p.C + c];
      mm = fmaxf(mm, s_red[(q * 2 + 1) * p.C + c]);
    }
    p.psum[((int64_t)n * p.parts + part) * p.C + c] = ss;
    p.pmax[((int64_t)n * p.parts + part) * p.C + c] = mm;
  }
}

// ------------------------------------------------------------------------------------------ CBAM_MLP
struct MlpArgs {
  const float* psum;
  const float* pmax;
  const float* fc1;  // [G][hidden][C]
  const float* fc2;  // [G][C][hidden]
  float* gate;       // [n_img][C]
  int n_img, group_imgs, C, hidden, parts;
  float inv_hw;
};

__global__ void __launch_bounds__(256) cbam_mlp_kernel(const MlpArgs p) {
  extern __shared__ float s_mlp[];  // avg[C], max[C], hid[hidden]
  float* s_avg = s_mlp;
  float* s_max = s_mlp + p.C;
  float* s_hid = s_mlp + 2 * p.C;
  const int n = blockIdx.x;
  const int g = n / p.group_imgs;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    float ss = 0.0f, mm = -INFINITY;
    for (int q = 0; q < p.parts; ++q) {
      ss += p.psum[((int64_t)n * p.parts + q) * p.C + c];
      mm = fmaxf(mm, p.pmax[((int64_t)n * p.parts + q) * p.C + c]);
    }
    s_avg[c] = ss * p.inv_hw;
    s_max[c] = mm;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int h = warp; h < p.hidden; h += nwarps) {
    const float* w1 = p.fc1 + ((int64_t)g * p.hidden + h) * p.C;
    float da = 0.0f, dm = 0.0f;
    for (int c = lane; c < p.C; c += 32) {
      const float w = __ldg(w1 + c);
      da = fmaf(w, s_avg[c], da);
      dm = fmaf(w, s_max[c], dm);
    }
    da = warp_sum(da);
    dm = warp_sum(dm);
    // fc2 is linear: fc2(relu(a)) + fc2(relu(m)) == fc2(relu(a) + relu(m))
    if (lane == 0) s_hid[h] = fmaxf(da, 0.0f) + fmaxf(dm, 0.0f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const float* w2 = p.fc2 + ((int64_t)g * p.C + c) * p.hidden;
    float o = 0.0f;
    for (int h = 0; h < p.hidden; ++h) o = fmaf(__ldg(w2 + h), s_hid[h], o);
    p.gate[(int64_t)n * p.C + c] = 1.0f / (1.0f + expf(-o));
  }
}

// ------------------------------------------------------------------------------------------ CBAM_STATS
struct StatsArgs {
  View<const __nv_bfloat16> x;
  const float* gate;  // [n_img][C]
  float* stats;       // [n_img][HW][2]
  int n_img, HW, C;
  int L;              // lanes cooperating on one pixel (power of two, <= 32)
  int64_t total_pix;
};

__global__ void __launch_bounds__(256) cbam_stats_kernel(const StatsArgs p) {
  const int c8n = p.C >> 3;
  const int lane = threadIdx.x & 31;
  const int sub = lane % p.L;        // lane inside the pixel group
  const int grp = lane / p.L;        // pixel group inside the warp
  const int gpw = 32 / p.L;          // pixels per warp iteration
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float inv_c = 1.0f / (float)p.C;
  for (int64_t base = warp_id * gpw; base < p.total_pix; base += nwarps * gpw) {
    const int64_t pi = base + grp;
    const bool ok = pi < p.total_pix;
    float s = 0.0f, m = -INFINITY;
    if (ok) {
      const int n = (int)(pi / p.HW);
      const int px = (int)(pi - (int64_t)n * p.HW);
      const __nv_bfloat16* xin = p.x.p + p.x.img_off(n) + (int64_t)px * p.x.ld;
      const float* gt = p.gate + (int64_t)n * p.C;
      for (int c8 = sub; c8 < c8n; c8 += p.L) {
        float v[8];
        unpack8(ldg128(xin + c8 * 8), v);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gt + c8 * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gt + c8 * 8) + 1);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float t = v[e] * gg[e];
          s += t;
          m = fmaxf(m, t);
        }
      }
    }
    for (int o = p.L >> 1; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if (ok && sub == 0) *reinterpret_cast<float2*>(p.stats + pi * 2) = make_float2(s * inv_c, m);
  }
}

// ------------------------------------------------------------------------------------------ CBAM_APPLY
constexpr int RB = 4;  // image rows per CTA

struct ApplyArgs {
  View<const __nv_bfloat16> x;
  View<__nv_bfloat16> y;
  const float* gate;   // [n_img][C]
  const float* stats;  // [n_img][H][W][2]
  const float* w7;     // [G][2][7][7]
  int n_img, group_imgs, H, W, C, bands;
};

__global__ void __launch_bounds__(256) cbam_apply_kernel(const ApplyArgs p) {
  extern __shared__ float s_ap[];
  const int SW = p.W + 6;
  float* s_st = s_ap;                        // [RB+6][SW][2]
  float* s_s = s_ap + (RB + 6) * SW * 2;     // [RB][W]
  float* s_w = s_s + RB * p.W;               // [98]
  const int n = blockIdx.x / p.bands;
  const int band = blockIdx.x - n * p.bands;
  const int y0 = band * RB;
  const int g = n / p.group_imgs;
  const float* st = p.stats + (int64_t)n * p.H * p.W * 2;
  for (int i = threadIdx.x; i < (RB + 6) * SW; i += blockDim.x) {
    const int r = i / SW, q = i - r * SW;
    const int iy = y0 + r - 3, ix = q - 3;
    float2 v = make_float2(0.0f, 0.0f);
    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v = *reinterpret_cast<const float2*>(st + ((int64_t)iy * p.W + ix) * 2);
    s_st[i * 2] = v.x;
    s_st[i * 2 + 1] = v.y;
  }
  for (int i = threadIdx.x; i < 98; i += blockDim.x) s_w[i] = __ldg(p.w7 + (int64_t)g * 98 + i);
  __syncthreads();
  for (int i = threadIdx.x; i < RB * p.W; i += blockDim.x) {
    const int r = i / p.W, q = i - r * p.W;
    float acc = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float* sp = s_st + ((r + ky) * SW + q + kx) * 2;
        acc = fmaf(s_w[ky * 7 + kx], sp[0], acc);
        acc = fmaf(s_w[49 + ky * 7 + kx], sp[1], acc);
      }
    s_s[i] = 1.0f / (1.0f + __expf(-acc));
  }
  __syncthreads();
  const int c8n = p.C >> 3;
  const int rows = min(RB, p.H - y0);
  const __nv_bfloat16* xin = p.x.p + p.x.img_off(n);
  __nv_bfloat16* yout = p.y.p + p.y.img_off(n);
  const float* gt = p.gate + (int64_t)n * p.C;
  for (int i = threadIdx.x; i < rows * p.W * c8n; i += blockDim.x) {
    const int c8 = i % c8n;
    const int lp = i / c8n;  // pixel inside the band
    const int64_t pix = (int64_t)y0 * p.W + lp;
    float v[8];
    unpack8(ldg128(xin + pix * p.x.ld + c8 * 8), v);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gt + c8 * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gt + c8 * 8) + 1);
    const float sp = s_s[lp];
    v[0] *= g0.x * sp; v[1] *= g0.y * sp; v[2] *= g0.z * sp; v[3] *= g0.w * sp;
    v[4] *= g1.x * sp; v[5] *= g1.y * sp; v[6] *= g1.z * sp; v[7] *= g1.w * sp;
    stg128(yout + pix * p.y.ld + c8 * 8, pack8(v));
  }
}

inline bool view_aligned(const void* ptr, int ld, int64_t img_stride, int64_t gstride) {
  return ((uintptr_t)ptr % 16) == 0 && ld % 8 == 0 && img_stride % 8 == 0 && gstride % 8 == 0;
}

}  // namespace

int launch_cbam_pool(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  PoolArgs a;
  a.x = resolve<const __nv_bfloat16>(op.x, bufs);
  a.psum = resolve_ptr<float>(op.a0, bufs);
  a.pmax = resolve_ptr<float>(op.a1, bufs);
  a.n_img = op.n_img; a.HW = op.Hi * op.Wi; a.C = op.Cin; a.parts = op.parts;
  DCFA_REQUIRE(a.x.p && a.psum && a.pmax, "cbam_pool: missing tensor");
  DCFA_REQUIRE(a.C > 0 && a.C % 8 == 0 && a.C <= 2048, "cbam_pool: C %d unsupported", a.C);
  DCFA_REQUIRE(a.parts >= 1 && a.parts <= a.HW, "cbam_pool: parts %d out of range", a.parts);
  DCFA_REQUIRE(view_aligned(a.x.p, a.x.ld, a.x.img_stride, a.x.gstride), "cbam_pool: view must be 16-byte aligned");
  a.pix_per_part = ceil_div(a.HW, a.parts);
  const int planes = 256 / (a.C >> 3);
  const size_t smem = (size_t)planes * 2 * a.C * sizeof(float);
  DCFA_REQUIRE(smem <= 48 * 1024, "cbam_pool: shared memory %zu too large", smem);
  cbam_pool_kernel<<<(unsigned)(a.n_img * a.parts), 256, smem, st>>>(a);
  DCFA_CHECK_LAUNCH("cbam_pool_kernel");
  return DCFA_OK;
}

int launch_cbam_mlp(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  MlpArgs a;
  a.psum = resolve_ptr<const float>(op.a0, bufs);
  a.pmax = resolve_ptr<const float>(op.a1, bufs);
  a.fc1 = resolve_ptr<const float>(op.w, bufs);
  a.fc2 = resolve_ptr<const float>(op.scale, bufs);
  a.gate = resolve_ptr<float>(op.a2, bufs);
  a.n_img = op.n_img;
  a.group_imgs = op.group_imgs > 0 ? op.group_imgs : op.n_img;
  a.C = op.Cin; a.hidden = op.hidden; a.parts = op.parts;
  a.inv_hw = 1.0f / (float)(op.Hi * op.Wi);
  DCFA_REQUIRE(a.psum && a.pmax && a.fc1 && a.fc2 && a.gate, "cbam_mlp: missing tensor");
  DCFA_REQUIRE(a.hidden >= 1 && a.C >= 1, "cbam_mlp: bad sizes");
  const size_t smem = (size_t)(2 * a.C + a.hidden) * sizeof(float);
  cbam_mlp_kernel<<<(unsigned)a.n_img, 256, smem, st>>>(a);
  DCFA_CHECK_LAUNCH("cbam_mlp_kernel");
  return DCFA_OK;
}

int launch_cbam_stats(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  StatsArgs a;
  a.x = resolve<const __nv_bfloat16>(op.x, bufs);
  a.gate = resolve_ptr<const float>(op.a2, bufs);
  a.stats = resolve_ptr<float>(op.a0, bufs);
  a.n_img = op.n_img; a.HW = op.Hi * op.Wi; a.C = op.Cin;
  DCFA_REQUIRE(a.x.p && a.gate && a.stats, "cbam_stats: missing tensor");
  DCFA_REQUIRE(a.C > 0 && a.C % 8 == 0, "cbam_stats: C %d unsupported", a.C);
  DCFA_REQUIRE(view_aligned(a.x.p, a.x.ld, a.x.img_stride, a.x.gstride) && ((uintptr_t)a.gate % 16) == 0 &&
                   ((uintptr_t)a.stats % 8) == 0,
               "cbam_stats: misaligned tensor");
  int L = 1;
  while (L < 32 && L < (a.C >> 3)) L <<= 1;
  a.L = L;
  a.total_pix = (int64_t)a.n_img * a.HW;
  const int gpw = 32 / L;
  int64_t blocks = (a.total_pix + (int64_t)8 * gpw - 1) / ((int64_t)8 * gpw);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cbam_stats_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  DCFA_CHECK_LAUNCH("cbam_stats_kernel");
  return DCFA_OK;
}

int launch_cbam_apply(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  ApplyArgs a;
  a.x = resolve<const __nv_bfloat16>(op.x, bufs);
  a.y = resolve<__nv_bfloat16>(op.y, bufs);
  a.gate = resolve_ptr<const float>(op.a2, bufs);
  a.stats = resolve_ptr<const float>(op.a0, bufs);
  a.w7 = resolve_ptr<const float>(op.w, bufs);
  a.n_img = op.n_img;
  a.group_imgs = op.group_imgs > 0 ? op.group_imgs : op.n_img;
  a.H = op.Hi; a.W = op.Wi; a.C = op.Cin;
  DCFA_REQUIRE(a.x.p && a.y.p && a.gate && a.stats && a.w7, "cbam_apply: missing tensor");
  DCFA_REQUIRE(a.C > 0 && a.C % 8 == 0, "cbam_apply: C %d unsupported", a.C);
  DCFA_REQUIRE(view_aligned(a.x.p, a.x.ld, a.x.img_stride, a.x.gstride) &&
                   view_aligned(a.y.p, a.y.ld, a.y.img_stride, a.y.gstride) && ((uintptr_t)a.gate % 16) == 0 &&
                   ((uintptr_t)a.stats % 8) == 0,
               "cbam_apply: misaligned tensor");
  a.bands = ceil_div(a.H, RB);
  const size_t smem = (size_t)((RB + 6) * (a.W + 6) * 2 + RB * a.W + 98) * sizeof(float);
  DCFA_REQUIRE(smem <= 48 * 1024, "cbam_apply: image width %d too large", a.W);
  cbam_apply_kernel<<<(unsigned)(a.n_img * a.bands), 256, smem, st>>>(a);
  DCFA_CHECK_LAUNCH("cbam_apply_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
