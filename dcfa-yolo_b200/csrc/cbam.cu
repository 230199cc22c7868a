// CBAM attention (nets/yolo_mul.py:56-102) as four memory-bound kernels on bf16 NHWC:
//   CBAM_POOL   per-(image, channel) sum and max over HW, written as per-chunk partials   (:59-60,:70-71)
//   CBAM_MLP    fixed-order reduction of the partials, fc1 -> ReLU -> fc2 on the average and the max
//               vector, add, sigmoid -> channel gate                                      (:63-73)
//   CBAM_STATS  t = x * gate; per-pixel mean and max over channels                         (:100, :86-88)
//   CBAM_APPLY  7x7 conv over the 2-plane map (zero padded), sigmoid, y = x * gate * s     (:89-90, :101)
// Used at the six fusion sites (:346-353, :403-415) and inside SPPF_CBAM (:18-31, hidden width 1).
// All channel accesses are 128-bit (8 x bf16); reductions are warp shuffles / fixed-order loops, so the
// result is deterministic.
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {
namespace {

// ------------------------------------------------------------------------------------------ CBAM_POOL
struct PoolArgs {
  View<const __nv_bfloat16> x;
  float* psum;  // [n_img][parts][C]
  float* pmax;
  int n_img, HW, C, parts, pix_per_part;
};

constexpr int kU = 4;  // independent 128-bit loads in flight per thread in the streaming loops below

// Thread mapping shared by POOL / STATS / APPLY: the CTA's 256 threads form `planes` pixel planes of c8n
// threads; a thread keeps ONE 8-channel chunk for its whole life (its gate values stay in registers) and walks
// pixels plane, plane + planes, ...   (c8n <= 256; threads beyond planes * c8n idle when 256 % c8n != 0)
__global__ void __launch_bounds__(256) cbam_pool_kernel(const PoolArgs p) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  extern __shared__ float s_red[];  // [planes][2][C]
  const int c8n = p.C >> 3;          // <= 256 (C <= 2048)
  const int planes = 256 / c8n;
  const int n = blockIdx.x / p.parts;
  const int part = blockIdx.x - n * p.parts;
  const int p0 = part * p.pix_per_part;
  const int p1 = min(p.HW, p0 + p.pix_per_part);
  const int c8 = (int)threadIdx.x % c8n;
  const int plane = (int)threadIdx.x / c8n;
  if (plane < planes) {
    const __nv_bfloat16* xin = p.x.p + p.x.img_off(n) + c8 * 8;
    float s[8], m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] = 0.0f; m[e] = -INFINITY; }
    for (int px = p0 + plane; px < p1; px += planes * kU) {
      uint4 r[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (px + u * planes < p1) r[u] = ldg128(xin + (int64_t)(px + u * planes) * p.x.ld);
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (px + u * planes < p1) {
          float v[8];
          unpack8(r[u], v);
#pragma unroll
          for (int e = 0; e < 8; ++e) { s[e] += v[e]; m[e] = fmaxf(m[e], v[e]); }
        }
    }
    float4* ds = reinterpret_cast<float4*>(s_red + (plane * 2 + 0) * p.C + c8 * 8);
    float4* dm = reinterpret_cast<float4*>(s_red + (plane * 2 + 1) * p.C + c8 * 8);
    ds[0] = make_float4(s[0], s[1], s[2], s[3]); ds[1] = make_float4(s[4], s[5], s[6], s[7]);
    dm[0] = make_float4(m[0], m[1], m[2], m[3]); dm[1] = make_float4(m[4], m[5], m[6], m[7]);
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
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  extern __shared__ float s_mlp[];  // avg[C], max[C], hid[hidden]
  float* s_avg = s_mlp;
  float* s_max = s_mlp + p.C;
  float* s_hid = s_mlp + 2 * p.C;
  const int n = blockIdx.x;
  const int g = n / p.group_imgs;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    float ss = 0.0f, mm = -INFINITY;
    const float* ps = p.psum + (int64_t)n * p.parts * p.C + c;
    const float* pm = p.pmax + (int64_t)n * p.parts * p.C + c;
    for (int q0 = 0; q0 < p.parts; q0 += 8) {
      float a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (q0 + u < p.parts) { a[u] = __ldg(ps + (int64_t)(q0 + u) * p.C); b[u] = __ldg(pm + (int64_t)(q0 + u) * p.C); }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (q0 + u < p.parts) { ss += a[u]; mm = fmaxf(mm, b[u]); }
    }
    s_avg[c] = ss * p.inv_hw;
    s_max[c] = mm;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int h = warp; h < p.hidden; h += nwarps) {
    const float* w1 = p.fc1 + ((int64_t)g * p.hidden + h) * p.C;
    float da = 0.0f, dm = 0.0f;
    for (int c0 = lane; c0 < p.C; c0 += 32 * 8) {
      float w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u * 32 < p.C) w[u] = __ldg(w1 + c0 + u * 32);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u * 32 < p.C) {
          da = fmaf(w[u], s_avg[c0 + u * 32], da);
          dm = fmaf(w[u], s_max[c0 + u * 32], dm);
        }
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
    if ((p.hidden & 3) == 0 && ((uintptr_t)p.fc2 & 15) == 0) {
      // the row is read as independent 128-bit loads (one L2 round trip, not `hidden` of them); same add order
      const float4* w4 = reinterpret_cast<const float4*>(w2);
      for (int h0 = 0; h0 < p.hidden; h0 += 16) {
        float4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (h0 + u * 4 < p.hidden) r[u] = __ldg(w4 + (h0 >> 2) + u);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (h0 + u * 4 < p.hidden) {
            const float* sh = s_hid + h0 + u * 4;
            o = fmaf(r[u].x, sh[0], o); o = fmaf(r[u].y, sh[1], o); o = fmaf(r[u].z, sh[2], o); o = fmaf(r[u].w, sh[3], o);
          }
      }
    } else {
      for (int h = 0; h < p.hidden; ++h) o = fmaf(__ldg(w2 + h), s_hid[h], o);
    }
    p.gate[(int64_t)n * p.C + c] = 1.0f / (1.0f + expf(-o));
  }
}

// ------------------------------------------------------------------------------------------ CBAM_STATS
struct StatsArgs {
  View<const __nv_bfloat16> x;
  const float* gate;  // [n_img][C]
  float* stats;       // [n_img][HW][2]
  int n_img, HW, C;
  int chunk;          // pixels per CTA (grid = chunks per image x images)
};

// Per pixel: mean and max over channels of x * gate.  Each thread reduces its own 8 channels, the per-chunk
// partials of one round of planes * kU pixels meet in shared memory ([pixel][c8n + 1] float2, double buffered:
// one barrier per round) and one thread per pixel adds them up in channel order.
__global__ void __launch_bounds__(256) cbam_stats_kernel(const StatsArgs p) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  extern __shared__ float2 s_part[];  // [2][planes * kU][c8n + 1]
  const int c8n = p.C >> 3;
  const int planes = 256 / c8n;
  const int round_px = planes * kU;
  const int pitch = c8n + 1;
  const int n = blockIdx.y;
  const int q0 = blockIdx.x * p.chunk;
  const int q1 = min(p.HW, q0 + p.chunk);
  const int c8 = (int)threadIdx.x % c8n;
  const int plane = (int)threadIdx.x / c8n;
  const bool active = plane < planes;
  const __nv_bfloat16* xin = p.x.p + p.x.img_off(n) + c8 * 8;
  float g[8];
  {
    const float4* gp = reinterpret_cast<const float4*>(p.gate + (int64_t)n * p.C + c8 * 8);
    const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
  }
  const float inv_c = 1.0f / (float)p.C;
  int buf = 0;
  for (int base = q0; base < q1; base += round_px, buf ^= 1) {
    float2* part = s_part + buf * round_px * pitch;
    if (active) {
      uint4 r[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int px = base + u * planes + plane;
        if (px < q1) r[u] = ldg128(xin + (int64_t)px * p.x.ld);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int px = base + u * planes + plane;
        if (px < q1) {
          float v[8];
          unpack8(r[u], v);
          float sm = 0.0f, mx = -INFINITY;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float t = v[e] * g[e];
            sm += t;
            mx = fmaxf(mx, t);
          }
          part[(u * planes + plane) * pitch + c8] = make_float2(sm, mx);
        }
      }
    }
    __syncthreads();
    const int px = base + (int)threadIdx.x;
    if ((int)threadIdx.x < round_px && px < q1) {
      const float2* row = part + threadIdx.x * pitch;
      float sm = 0.0f, mx = -INFINITY;
      for (int k = 0; k < c8n; ++k) {
        const float2 t = row[k];
        sm += t.x;
        mx = fmaxf(mx, t.y);
      }
      *reinterpret_cast<float2*>(p.stats + ((int64_t)n * p.HW + px) * 2) = make_float2(sm * inv_c, mx);
    }
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
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
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
  const int planes = 256 / c8n;
  const int c8 = (int)threadIdx.x % c8n;
  const int plane = (int)threadIdx.x / c8n;
  if (plane >= planes) return;
  const int npix = min(RB, p.H - y0) * p.W;
  const __nv_bfloat16* xin = p.x.p + p.x.img_off(n) + (int64_t)y0 * p.W * p.x.ld + c8 * 8;
  __nv_bfloat16* yout = p.y.p + p.y.img_off(n) + (int64_t)y0 * p.W * p.y.ld + c8 * 8;
  float gt8[8];
  {
    const float4* gp = reinterpret_cast<const float4*>(p.gate + (int64_t)n * p.C + c8 * 8);
    const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
    gt8[0] = g0.x; gt8[1] = g0.y; gt8[2] = g0.z; gt8[3] = g0.w; gt8[4] = g1.x; gt8[5] = g1.y; gt8[6] = g1.z; gt8[7] = g1.w;
  }
  for (int lp0 = plane; lp0 < npix; lp0 += planes * kU) {
    uint4 r[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (lp0 + u * planes < npix) r[u] = ldg128(xin + (int64_t)(lp0 + u * planes) * p.x.ld);
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int lp = lp0 + u * planes;
      if (lp < npix) {
        float v[8];
        unpack8(r[u], v);
        const float sp = s_s[lp];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= gt8[e] * sp;
        stg128(yout + (int64_t)lp * p.y.ld, pack8(v));
      }
    }
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
  launch_pdl(cbam_pool_kernel, dim3((unsigned)(a.n_img * a.parts)), dim3(256), smem, st, a);
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
  launch_pdl(cbam_mlp_kernel, dim3((unsigned)a.n_img), dim3(256), smem, st, a);
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
  DCFA_REQUIRE(a.C <= 2048, "cbam_stats: C %d unsupported", a.C);
  const int c8n = a.C >> 3;
  const int planes = 256 / c8n;
  const int round_px = planes * kU;
  // about 8 CTAs per SM over the whole grid, each CTA a whole number of rounds
  int chunks = ceil_div(sm_count() * 8, a.n_img);
  chunks = std::max(1, std::min(chunks, ceil_div(a.HW, round_px)));
  a.chunk = ceil_div(ceil_div(a.HW, chunks), round_px) * round_px;
  chunks = ceil_div(a.HW, a.chunk);
  const size_t smem = (size_t)2 * round_px * (c8n + 1) * sizeof(float2);
  launch_pdl(cbam_stats_kernel, dim3((unsigned)chunks, (unsigned)a.n_img), dim3(256), smem, st, a);
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
  launch_pdl(cbam_apply_kernel, dim3((unsigned)(a.n_img * a.bands)), dim3(256), smem, st, a);
  DCFA_CHECK_LAUNCH("cbam_apply_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
