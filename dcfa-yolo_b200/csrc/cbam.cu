// CBAM attention (nets/yolo_mul.py:56-102) as four memory-bound kernels on bf16 NHWC:
//   CBAM_POOL   per-(image, channel) sum and max over HW, written as per-chunk partials   (:59-60,:70-71)
//   CBAM_MLP    fixed-order reduction of the partials, fc1 -> ReLU -> fc2 on the average and the max
//               vector, add, sigmoid -> channel gate                                      (:63-73)
//   CBAM_STATS  t = x * gate; per-pixel mean and max over channels                         (:100, :86-88)
//   CBAM_APPLY  7x7 conv over the 2-plane map (zero padded), sigmoid, y = x * gate * s     (:89-90, :101)
// Used at the six fusion sites (:346-353, :403-415) and inside SPPF_CBAM (:18-31, hidden width 1).
// All channel accesses are 128-bit (8 x bf16); reductions are warp shuffles / fixed-order loops, so the
// result is deterministic.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <cstdio>

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
    for (int t = threadIdx.x; t < round_px && base + t < q1; t += 256) {
      const float2* row = part + t * pitch;
      float sm = 0.0f, mx = -INFINITY;
      for (int k = 0; k < c8n; ++k) {
        const float2 v = row[k];
        sm += v.x;
        mx = fmaxf(mx, v.y);
      }
      *reinterpret_cast<float2*>(p.stats + ((int64_t)n * p.HW + base + t) * 2) = make_float2(sm * inv_c, mx);
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

// ------------------------------------------------------------------------------------------ fused CBAM
// One thread-block CLUSTER per image runs the whole CBAM (pool -> MLP -> stats -> 7x7 -> apply) in one launch; the
// four kernels above stay as the path for shapes it does not take.  CTA r of the cluster owns a band of image rows.
//   A  channel sums / maxima of the band, combined across the cluster through DSMEM (fixed order: deterministic);
//   B  the MLP, recomputed by every CTA (weights are L2-resident, C*hidden*2 floats);
//   C  per-pixel mean/max over channels of x*gate for the band, into a zero-bordered stats tile in shared memory;
//      the 3 halo rows on each side are read from the neighbouring CTAs' tiles through DSMEM;
//   D  7x7 conv + sigmoid for the band's pixels;   E  y = x * gate * s  (x read a third time: L2).
// The launch count of a forward drops by 21, and the per-launch fill/drain of three kernels per CBAM disappears.
constexpr int kFusedThreads = 1024;

struct FusedArgs {
  View<const __nv_bfloat16> x;
  View<__nv_bfloat16> y;
  const float* fc1;   // [G][hidden][C]
  const float* fc2;   // [G][C][hidden]
  const float* w7;    // [G][2][7][7]
  int n_img, group_imgs, H, W, C, hidden;
  int rows_per;       // image rows per CTA
};

__global__ void __launch_bounds__(kFusedThreads) cbam_fused_kernel(const FusedArgs p) {
  extern __shared__ __align__(16) float s_f[];
  namespace cg = cooperative_groups;
  ptx::pdl_launch_dependents();
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  const int n = blockIdx.x / CS;
  const int g = n / p.group_imgs;
  const int tid = threadIdx.x;
  const int C = p.C, W = p.W;
  const int c8n = C >> 3;
  const int planes = kFusedThreads / c8n;
  const int c8 = tid % c8n, plane = tid / c8n;
  const bool active = plane < planes;
  const int y0 = min(p.H, crank * p.rows_per), y1 = min(p.H, y0 + p.rows_per);
  const int npix = (y1 - y0) * W;
  const int SW = W + 6, SR = p.rows_per + 6;
  const int round_px = planes * kU;

  // shared memory (floats)
  float* part_sum = s_f;                       // [C]   this CTA's band
  float* part_max = part_sum + C;              // [C]
  float* s_avg = part_max + C;                 // [C]
  float* s_max = s_avg + C;                    // [C]
  float* s_gate = s_max + C;                   // [C]
  float* s_hid = s_gate + C;                   // [hidden] (padded to a multiple of 4)
  float* s_w = s_hid + ((p.hidden + 3) & ~3);  // [100]
  float* s_st = s_w + 100;                     // [SR][SW][2] stats tile, zero border
  float* s_sig = s_st + SR * SW * 2;           // [rows_per * W]
  float* scratch = s_sig + ((p.rows_per * W + 3) & ~3);   // max(planes*2*C floats, 2*round_px*(c8n+1) float2)

  ptx::pdl_wait();
  const __nv_bfloat16* xin = p.x.p + p.x.img_off(n) + (int64_t)y0 * W * p.x.ld + c8 * 8;

  // ---- A: channel sums / maxima of the band
  {
    float sm[8], mx[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sm[e] = 0.0f; mx[e] = -INFINITY; }
    if (active) {
      for (int px = plane; px < npix; px += planes * kU) {
        uint4 r[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u)
          if (px + u * planes < npix) r[u] = ldg128(xin + (int64_t)(px + u * planes) * p.x.ld);
#pragma unroll
        for (int u = 0; u < kU; ++u)
          if (px + u * planes < npix) {
            float v[8];
            unpack8(r[u], v);
#pragma unroll
            for (int e = 0; e < 8; ++e) { sm[e] += v[e]; mx[e] = fmaxf(mx[e], v[e]); }
          }
      }
      float4* ds = reinterpret_cast<float4*>(scratch + (plane * 2 + 0) * C + c8 * 8);
      float4* dm = reinterpret_cast<float4*>(scratch + (plane * 2 + 1) * C + c8 * 8);
      ds[0] = make_float4(sm[0], sm[1], sm[2], sm[3]); ds[1] = make_float4(sm[4], sm[5], sm[6], sm[7]);
      dm[0] = make_float4(mx[0], mx[1], mx[2], mx[3]); dm[1] = make_float4(mx[4], mx[5], mx[6], mx[7]);
    }
    __syncthreads();
    for (int c = tid; c < C; c += kFusedThreads) {
      float ss = 0.0f, mm = -INFINITY;
      for (int q = 0; q < planes; ++q) {
        ss += scratch[(q * 2 + 0) * C + c];
        mm = fmaxf(mm, scratch[(q * 2 + 1) * C + c]);
      }
      part_sum[c] = ss;
      part_max[c] = mm;
    }
    for (int i = tid; i < 98; i += kFusedThreads) s_w[i] = __ldg(p.w7 + (int64_t)g * 98 + i);
    for (int i = tid; i < SR * SW * 2; i += kFusedThreads) s_st[i] = 0.0f;
  }
  cluster.sync();
  {
    const float inv_hw = 1.0f / (float)(p.H * W);
    for (int c = tid; c < C; c += kFusedThreads) {
      float ss = 0.0f, mm = -INFINITY;
      for (int r = 0; r < CS; ++r) {             // fixed order over the cluster: deterministic
        ss += cluster.map_shared_rank(part_sum, r)[c];
        mm = fmaxf(mm, cluster.map_shared_rank(part_max, r)[c]);
      }
      s_avg[c] = ss * inv_hw;
      s_max[c] = mm;
    }
  }
  __syncthreads();

  // ---- B: fc1 -> ReLU -> fc2 on both vectors, add, sigmoid (fc2 is linear: one pass over relu(a)+relu(m))
  {
    const int warp = tid >> 5, lane = tid & 31, nwarps = kFusedThreads >> 5;
    for (int h = warp; h < p.hidden; h += nwarps) {
      const float* w1 = p.fc1 + ((int64_t)g * p.hidden + h) * C;
      float da = 0.0f, dm = 0.0f;
      for (int c0 = lane; c0 < C; c0 += 32 * 8) {
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (c0 + u * 32 < C) w[u] = __ldg(w1 + c0 + u * 32);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (c0 + u * 32 < C) {
            da = fmaf(w[u], s_avg[c0 + u * 32], da);
            dm = fmaf(w[u], s_max[c0 + u * 32], dm);
          }
      }
      da = warp_sum(da);
      dm = warp_sum(dm);
      if (lane == 0) s_hid[h] = fmaxf(da, 0.0f) + fmaxf(dm, 0.0f);
    }
    __syncthreads();
    for (int c = tid; c < C; c += kFusedThreads) {
      const float* w2 = p.fc2 + ((int64_t)g * C + c) * p.hidden;
      float o = 0.0f;
      if ((p.hidden & 3) == 0 && ((uintptr_t)p.fc2 & 15) == 0) {
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
      s_gate[c] = 1.0f / (1.0f + expf(-o));
    }
  }
  __syncthreads();

  float g8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) g8[e] = s_gate[c8 * 8 + e];

  // ---- C: per-pixel mean / max over channels of x * gate for the band -> interior of the stats tile
  {
    float2* part = reinterpret_cast<float2*>(scratch);
    const int pitch = c8n + 1;
    const float inv_c = 1.0f / (float)C;
    int buf = 0;
    for (int base = 0; base < npix; base += round_px, buf ^= 1) {
      float2* pb = part + buf * round_px * pitch;
      if (active) {
        uint4 r[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int px = base + u * planes + plane;
          if (px < npix) r[u] = ldg128(xin + (int64_t)px * p.x.ld);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int px = base + u * planes + plane;
          if (px < npix) {
            float v[8];
            unpack8(r[u], v);
            float sm = 0.0f, mx = -INFINITY;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float t = v[e] * g8[e];
              sm += t;
              mx = fmaxf(mx, t);
            }
            pb[(u * planes + plane) * pitch + c8] = make_float2(sm, mx);
          }
        }
      }
      __syncthreads();
      for (int t = tid; t < round_px && base + t < npix; t += kFusedThreads) {
        const float2* row = pb + t * pitch;
        float sm = 0.0f, mx = -INFINITY;
        for (int k = 0; k < c8n; ++k) {
          const float2 v = row[k];
          sm += v.x;
          mx = fmaxf(mx, v.y);
        }
        const int px = base + t;
        const int ry = px / W, rx = px - ry * W;
        *reinterpret_cast<float2*>(s_st + ((ry + 3) * SW + rx + 3) * 2) = make_float2(sm * inv_c, mx);
      }
    }
  }
  cluster.sync();
  // halo rows: the last 3 rows of the band above, the first 3 rows of the band below (zero outside the image)
  for (int i = tid; i < 6 * W; i += kFusedThreads) {
    const int hr = i / W, rx = i - hr * W;               // hr 0..2 above, 3..5 below
    const int gy = hr < 3 ? y0 - 3 + hr : y1 + hr - 3;   // image row
    if (gy >= 0 && gy < p.H && y1 > y0) {
      const int owner = gy / p.rows_per;
      const int ly = gy - owner * p.rows_per;
      const float2 v = *reinterpret_cast<const float2*>(cluster.map_shared_rank(s_st, owner) + ((ly + 3) * SW + rx + 3) * 2);
      const int dr = hr < 3 ? hr : (y1 - y0) + hr;       // tile row
      *reinterpret_cast<float2*>(s_st + (dr * SW + rx + 3) * 2) = v;
    }
  }
  cluster.sync();   // no remote access to this CTA's shared memory after this point

  // ---- D: 7x7 conv over (mean, max), sigmoid
  for (int i = tid; i < npix; i += kFusedThreads) {
    const int r = i / W, q = i - r * W;
    float acc = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float2 sp = *reinterpret_cast<const float2*>(s_st + ((r + ky) * SW + q + kx) * 2);
        acc = fmaf(s_w[ky * 7 + kx], sp.x, acc);
        acc = fmaf(s_w[49 + ky * 7 + kx], sp.y, acc);
      }
    s_sig[i] = 1.0f / (1.0f + __expf(-acc));
  }
  __syncthreads();

  // ---- E: y = x * gate * s
  if (!active) return;
  __nv_bfloat16* yout = p.y.p + p.y.img_off(n) + (int64_t)y0 * W * p.y.ld + c8 * 8;
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
        const float sp = s_sig[lp];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= g8[e] * sp;
        stg128(yout + (int64_t)lp * p.y.ld, pack8(v));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ fused SPPF_CBAM
// SPPF_CBAM (nets/yolo_mul.py:18-31) after its first conv:  x1 = cbam1(t), x2 = cbam2(pool(x1)), x3 = cbam3(pool(x2)),
// x4 = cbam4(pool(x3)), each written to its slot of the concat buffer.  As separate launches that is 4 fused CBAMs and
// 3 max-pools on a per-image tensor of a few hundred KB -- seven launches of 14-21 us that are all fill and drain.
// Here one thread-block cluster per image keeps the image RESIDENT in shared memory for the whole sequence: CTA r owns
// a band of rows ([rows][W][C] bf16), a stage runs the CBAM phases of cbam_fused_kernel on the band in shared memory
// (channel sums and the stats halo cross the cluster through DSMEM), stores y = x * gate * s to the stage's concat
// slot AND back into the band, and the band is then max-pooled 5x5 in place: a horizontal 1x5 pass, the two edge rows
// of each neighbouring band fetched through DSMEM, a vertical 5x1 pass.  bf16 maxima are exact, so the values equal
// the separate-kernel path wherever the CBAM sums are associated in the same order.
constexpr int kSppfStages = 4;

struct SppfArgs {
  View<const __nv_bfloat16> x;
  View<__nv_bfloat16> y[kSppfStages];
  const float* fc1[kSppfStages];   // [G][hidden][C]
  const float* fc2[kSppfStages];   // [G][C][hidden]
  const float* w7[kSppfStages];    // [G][2][7][7]
  int n_img, group_imgs, H, W, C, hidden, rows_per, stages;
};

__device__ __forceinline__ uint4 bf16x8_max(const uint4& a, const uint4& b) {
  uint4 r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r.x) : "r"(a.x), "r"(b.x));
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r.y) : "r"(a.y), "r"(b.y));
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r.z) : "r"(a.z), "r"(b.z));
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r.w) : "r"(a.w), "r"(b.w));
  return r;
}

// NT threads per CTA: 1024 (one CTA per SM) or 512 (two CTAs per SM: one works while the other waits at a cluster barrier)
template <int NT>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : 1) sppf_cbam_kernel(const SppfArgs p) {
  constexpr int kSppfThreads = NT;
  extern __shared__ __align__(16) uint8_t s_raw[];
  namespace cg = cooperative_groups;
  ptx::pdl_launch_dependents();
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  const int n = blockIdx.x / CS;
  const int g = n / p.group_imgs;
  const int tid = threadIdx.x, lane = tid & 31;
  const int C = p.C, W = p.W, rows = p.rows_per;       // H == CS * rows (checked by the launcher)
  const int c8n = C >> 3;                              // 8, 16 or 32: a pixel's chunks live in one warp
  const int planes = kSppfThreads / c8n;
  const int c8 = tid % c8n, plane = tid / c8n;
  const int y0 = crank * rows, y1 = y0 + rows;
  const int npix = rows * W;
  const int SW = W + 6, SR = rows + 6;
  const int planesA = min(planes, W);                  // phase-A partials live in the halo buffer: W planes of 2 * C floats fit
  const uint4 neg_inf = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
  const size_t row_bytes = (size_t)W * C * 2;

  uint8_t* X = s_raw;                                  // [rows][W][C] bf16: the band
  uint8_t* HALO = X + (size_t)rows * row_bytes;        // [4][W][C] bf16: rows y0-2, y0-1, y1, y1+1 (horizontally pooled)
  float* part_sum = reinterpret_cast<float*>(HALO + 4 * row_bytes);
  float* part_max = part_sum + C;
  float* s_avg = part_max + C;
  float* s_max = s_avg + C;
  float* s_gate = s_max + C;
  float* s_hid = s_gate + C;
  float* s_w4 = s_hid + ((p.hidden + 3) & ~3);         // [stages][100]: the 7x7 taps of every stage
  float* s_st = s_w4 + 100 * kSppfStages;              // [SR][SW][2] stats tile, zero border
  float* s_sig = s_st + SR * SW * 2;                   // [rows * W]
  float* scratch = reinterpret_cast<float*>(HALO);     // phase A: [planesA][2][C]

#ifdef DCFA_SPPF_TIMING
  long long ts[96];
  int nts = 0;
#define SPPF_TS() do { if (tid == 0 && blockIdx.x == 0 && nts < 96) ts[nts++] = clock64(); } while (0)
#else
#define SPPF_TS() do {} while (0)
#endif
  // parameters do not depend on the previous kernel: the 7x7 taps of all stages into shared memory, the MLP weights of
  // all stages pulled into L2 (each stage reads them once, right on its critical path)
  for (int i = tid; i < 98 * p.stages; i += kSppfThreads) {
    const int st = i / 98, j = i - st * 98;
    s_w4[st * 100 + j] = __ldg(p.w7[st] + (int64_t)g * 98 + j);
  }
  {
    const int lines = (C * p.hidden * 4 + 127) / 128;
    for (int i = tid; i < 2 * p.stages * lines; i += kSppfThreads) {
      const int st = i / (2 * lines), r = i - st * 2 * lines;
      const float* w = (r < lines ? p.fc1[st] : p.fc2[st]) + (int64_t)g * C * p.hidden;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(w) + (size_t)(r % lines) * 128));
    }
  }
  ptx::pdl_wait();
  SPPF_TS();
  {
    const __nv_bfloat16* xin = p.x.p + p.x.img_off(n) + (int64_t)y0 * W * p.x.ld + c8 * 8;
    for (int px = plane; px < npix; px += planes * kU) {   // kU independent 128-bit loads in flight
      uint4 r[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (px + u * planes < npix) r[u] = ldg128(xin + (int64_t)(px + u * planes) * p.x.ld);
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (px + u * planes < npix) *reinterpret_cast<uint4*>(X + (size_t)(px + u * planes) * C * 2 + c8 * 16) = r[u];
    }
  }
  __syncthreads();
  SPPF_TS();   // 1: band loaded

  for (int s = 0; s < p.stages; ++s) {
    // ---- A: channel sums / maxima of the band
    if (plane < planesA) {
      float sm[8], mx[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { sm[e] = 0.0f; mx[e] = -INFINITY; }
      for (int px = plane; px < npix; px += planesA) {
        float v[8];
        unpack8(*reinterpret_cast<const uint4*>(X + (size_t)px * C * 2 + c8 * 16), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) { sm[e] += v[e]; mx[e] = fmaxf(mx[e], v[e]); }
      }
      float4* ds = reinterpret_cast<float4*>(scratch + (plane * 2 + 0) * C + c8 * 8);
      float4* dm = reinterpret_cast<float4*>(scratch + (plane * 2 + 1) * C + c8 * 8);
      ds[0] = make_float4(sm[0], sm[1], sm[2], sm[3]); ds[1] = make_float4(sm[4], sm[5], sm[6], sm[7]);
      dm[0] = make_float4(mx[0], mx[1], mx[2], mx[3]); dm[1] = make_float4(mx[4], mx[5], mx[6], mx[7]);
    }
    SPPF_TS();   // A1 accumulate
    __syncthreads();
    SPPF_TS();   // A2 barrier
    for (int c = tid; c < C; c += kSppfThreads) {
      float ss = 0.0f, mm = -INFINITY;
#pragma unroll 4
      for (int q = 0; q < planesA; ++q) {
        ss += scratch[(q * 2 + 0) * C + c];
        mm = fmaxf(mm, scratch[(q * 2 + 1) * C + c]);
      }
      part_sum[c] = ss;
      part_max[c] = mm;
    }
    for (int i = tid; i < SR * SW * 2; i += kSppfThreads) s_st[i] = 0.0f;
    SPPF_TS();   // A done
    cluster.sync();
    SPPF_TS();   // cluster sync
    {
      const float inv_hw = 1.0f / (float)(p.H * W);
      for (int c = tid; c < C; c += kSppfThreads) {
        float ss = 0.0f, mm = -INFINITY;
        for (int r = 0; r < CS; ++r) {             // fixed order over the cluster: deterministic
          ss += cluster.map_shared_rank(part_sum, r)[c];
          mm = fmaxf(mm, cluster.map_shared_rank(part_max, r)[c]);
        }
        s_avg[c] = ss * inv_hw;
        s_max[c] = mm;
      }
    }
    __syncthreads();

    SPPF_TS();   // combine
    // ---- B: fc1 -> ReLU -> fc2 on both vectors, add, sigmoid
    {
      const int warp = tid >> 5, nwarps = kSppfThreads >> 5;
      for (int h = warp; h < p.hidden; h += nwarps) {
        const float* w1 = p.fc1[s] + ((int64_t)g * p.hidden + h) * C;
        float da = 0.0f, dm = 0.0f;
        float w[8];                                   // C <= 256: all of a lane's weights in flight at once
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (lane + u * 32 < C) w[u] = __ldg(w1 + lane + u * 32);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (lane + u * 32 < C) {
            da = fmaf(w[u], s_avg[lane + u * 32], da);
            dm = fmaf(w[u], s_max[lane + u * 32], dm);
          }
        da = warp_sum(da);
        dm = warp_sum(dm);
        if (lane == 0) s_hid[h] = fmaxf(da, 0.0f) + fmaxf(dm, 0.0f);
      }
      __syncthreads();
      for (int c = tid; c < C; c += kSppfThreads) {
        const float* w2 = p.fc2[s] + ((int64_t)g * C + c) * p.hidden;
        float o = 0.0f;
        for (int h0 = 0; h0 < p.hidden; h0 += 8) {
          float w[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (h0 + u < p.hidden) w[u] = __ldg(w2 + h0 + u);
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (h0 + u < p.hidden) o = fmaf(w[u], s_hid[h0 + u], o);
        }
        s_gate[c] = 1.0f / (1.0f + expf(-o));
      }
    }
    __syncthreads();
    SPPF_TS();   // B (MLP)
    const float* s_w = s_w4 + s * 100;
    float g8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) g8[e] = s_gate[c8 * 8 + e];

    // ---- C: per-pixel mean / max over channels of x * gate -> interior of the stats tile (shuffles inside a pixel's lanes)
    {
      const float inv_c = 1.0f / (float)C;
      // A thread first forms the partials of up to 8 pixels (its chunk's 8 channels each), then the lanes of a pixel group
      // reduce them TRANSPOSED: in each of three rounds a lane hands half of its pixels to its partner and keeps the other
      // half (4 + 2 + 1 shuffles per value instead of 3 x 8), the remaining rounds are plain butterflies on one pixel.
      constexpr int PI = 8;
      for (int base = 0; base < npix; base += planes * PI) {
        float sm[PI], mx[PI];
#pragma unroll
        for (int u = 0; u < PI; ++u) {
          const int px = base + u * planes + plane;
          sm[u] = 0.0f; mx[u] = -INFINITY;
          if (px < npix) {
            float v[8];
            unpack8(*reinterpret_cast<const uint4*>(X + (size_t)px * C * 2 + c8 * 16), v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float t = v[e] * g8[e];
              sm[u] += t;
              mx[u] = fmaxf(mx[u], t);
            }
          }
        }
        int o = c8n >> 1, sel = 0;
#pragma unroll
        for (int half = PI / 2; half >= 1; half >>= 1) {
          const bool upper = (c8 & o) != 0;
#pragma unroll
          for (int u = 0; u < half; ++u) {
            const float s_send = upper ? sm[u] : sm[u + half], s_keep = upper ? sm[u + half] : sm[u];
            const float m_send = upper ? mx[u] : mx[u + half], m_keep = upper ? mx[u + half] : mx[u];
            sm[u] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, o);
            mx[u] = fmaxf(m_keep, __shfl_xor_sync(0xffffffffu, m_send, o));
          }
          if (upper) sel += half;
          o >>= 1;
        }
        for (; o > 0; o >>= 1) {
          sm[0] += __shfl_xor_sync(0xffffffffu, sm[0], o);
          mx[0] = fmaxf(mx[0], __shfl_xor_sync(0xffffffffu, mx[0], o));
        }
        const int px = base + sel * planes + plane;
        if ((c8 & ((c8n >> 3) - 1)) == 0 && px < npix) {
          const int ry = px / W, rx = px - ry * W;
          *reinterpret_cast<float2*>(s_st + ((ry + 3) * SW + rx + 3) * 2) = make_float2(sm[0] * inv_c, mx[0]);
        }
      }
    }
    SPPF_TS();   // C (stats)
    cluster.sync();
    // stats halo rows: the last 3 rows of the band above, the first 3 rows of the band below (zero outside the image)
    for (int i = tid; i < 6 * W; i += kSppfThreads) {
      const int hr = i / W, rx = i - hr * W;
      const int gy = hr < 3 ? y0 - 3 + hr : y1 + hr - 3;
      if (gy >= 0 && gy < p.H) {
        const int owner = gy / rows;
        const int ly = gy - owner * rows;
        const float2 v = *reinterpret_cast<const float2*>(cluster.map_shared_rank(s_st, owner) + ((ly + 3) * SW + rx + 3) * 2);
        const int dr = hr < 3 ? hr : rows + hr;
        *reinterpret_cast<float2*>(s_st + (dr * SW + rx + 3) * 2) = v;
      }
    }
    cluster.sync();
    SPPF_TS();   // stats halo (2 cluster syncs)

    // ---- D: 7x7 conv over (mean, max), sigmoid
    for (int i = tid; i < npix; i += kSppfThreads) {
      const int r = i / W, q = i - r * W;
      float acc = 0.0f;
#pragma unroll
      for (int ky = 0; ky < 7; ++ky)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float2 sp = *reinterpret_cast<const float2*>(s_st + ((r + ky) * SW + q + kx) * 2);
          acc = fmaf(s_w[ky * 7 + kx], sp.x, acc);
          acc = fmaf(s_w[49 + ky * 7 + kx], sp.y, acc);
        }
      s_sig[i] = 1.0f / (1.0f + __expf(-acc));
    }
    __syncthreads();
    SPPF_TS();   // D (7x7)

    // ---- E: y = x * gate * s -> the stage's concat slot and, in place, the band
    {
      __nv_bfloat16* yout = p.y[s].p + p.y[s].img_off(n) + (int64_t)y0 * W * p.y[s].ld + c8 * 8;
      for (int px = plane; px < npix; px += planes) {
        uint4* xp = reinterpret_cast<uint4*>(X + (size_t)px * C * 2 + c8 * 16);
        float v[8];
        unpack8(*xp, v);
        const float sp = s_sig[px];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= g8[e] * sp;
        const uint4 o = pack8(v);
        stg128(yout + (int64_t)px * p.y[s].ld, o);
        *xp = o;
      }
    }
    SPPF_TS();   // E (apply)
    if (s + 1 == p.stages) break;
    __syncthreads();

    // ---- 5x5 max-pool of the band, in place: horizontal pass (one thread = one row x one chunk, window in registers)
    for (int i = tid; i < rows * c8n; i += kSppfThreads) {
      const int r = i / c8n, cc = i - r * c8n;
      uint4* rowp = reinterpret_cast<uint4*>(X + (size_t)r * row_bytes + cc * 16);
      const int pitch = C * 2 / 16;   // uint4 per pixel
      uint4 a = neg_inf, b = neg_inf, c = rowp[0], d = W > 1 ? rowp[pitch] : neg_inf;
      for (int x = 0; x < W; ++x) {
        const uint4 e = x + 2 < W ? rowp[(x + 2) * pitch] : neg_inf;
        rowp[x * pitch] = bf16x8_max(bf16x8_max(bf16x8_max(a, b), bf16x8_max(c, d)), e);
        a = b; b = c; c = d; d = e;
      }
    }
    SPPF_TS();   // pool H
    cluster.sync();
    // the two edge rows of each neighbouring band (horizontally pooled), -inf outside the image
    for (int i = tid; i < 4 * W * c8n; i += kSppfThreads) {
      const int hr = i / (W * c8n), rem = i - hr * (W * c8n);
      const int gy = hr < 2 ? y0 - 2 + hr : y1 + hr - 2;
      uint4 v = neg_inf;
      if (gy >= 0 && gy < p.H) {
        const int owner = gy / rows, ly = gy - owner * rows;
        v = *reinterpret_cast<const uint4*>(cluster.map_shared_rank(X, owner) + (size_t)ly * row_bytes + (size_t)rem * 16);
      }
      *reinterpret_cast<uint4*>(HALO + (size_t)hr * row_bytes + (size_t)rem * 16) = v;
    }
    cluster.sync();   // every neighbour has its copy: the band may be overwritten
    SPPF_TS();   // pool halo (2 cluster syncs)
    // vertical pass (one thread = one column x one chunk)
    for (int i = tid; i < W * c8n; i += kSppfThreads) {
      const size_t off = (size_t)i * 16;     // (x, chunk) inside a row
      auto row_at = [&](int r) -> uint4 {    // r in [-2, rows + 2)
        if (r < 0) return *reinterpret_cast<const uint4*>(HALO + (size_t)(r + 2) * row_bytes + off);
        if (r >= rows) return *reinterpret_cast<const uint4*>(HALO + (size_t)(r - rows + 2) * row_bytes + off);
        return *reinterpret_cast<const uint4*>(X + (size_t)r * row_bytes + off);
      };
      uint4 a = row_at(-2), b = row_at(-1), c = row_at(0), d = row_at(1);
      for (int r = 0; r < rows; ++r) {
        const uint4 e = row_at(r + 2);
        *reinterpret_cast<uint4*>(X + (size_t)r * row_bytes + off) = bf16x8_max(bf16x8_max(bf16x8_max(a, b), bf16x8_max(c, d)), e);
        a = b; b = c; c = d; d = e;
      }
    }
    __syncthreads();
    SPPF_TS();   // pool V
  }
#ifdef DCFA_SPPF_TIMING
  if (tid == 0 && blockIdx.x == 0) {
    for (int i = 1; i < nts; ++i) printf("%d:%lld ", i, ts[i] - ts[i - 1]);
    printf("total %lld\n", ts[nts - 1] - ts[0]);
  }
#endif
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

// The four ops of one CBAM (consecutive POOL, MLP, STATS, APPLY records of the plan) as ONE cluster launch.
// Returns 1 if the fused kernel was launched, 0 if the shape is left to the four separate kernels, < 0 on error.
int launch_cbam_fused(const dcfa_op& pool, const dcfa_op& mlp, const dcfa_op& stats, const dcfa_op& apply, void* const* bufs,
                      cudaStream_t st) {
  bool force = false;
  {
    const char* e = getenv("DCFA_CBAM_FUSED");   // debug / tests: 0 keeps the four-kernel path, 2 fuses even tiny batches
    if (e && atoi(e) == 0) return 0;
    force = e && atoi(e) == 2;
  }
  FusedArgs a;
  a.x = resolve<const __nv_bfloat16>(apply.x, bufs);
  a.y = resolve<__nv_bfloat16>(apply.y, bufs);
  a.fc1 = resolve_ptr<const float>(mlp.w, bufs);
  a.fc2 = resolve_ptr<const float>(mlp.scale, bufs);
  a.w7 = resolve_ptr<const float>(apply.w, bufs);
  a.n_img = apply.n_img;
  a.group_imgs = apply.group_imgs > 0 ? apply.group_imgs : apply.n_img;
  a.H = apply.Hi; a.W = apply.Wi; a.C = apply.Cin; a.hidden = mlp.hidden;
  // the four records must describe the same tensor
  const View<const __nv_bfloat16> xp = resolve<const __nv_bfloat16>(pool.x, bufs), xs = resolve<const __nv_bfloat16>(stats.x, bufs);
  if (!(a.x.p && a.y.p && a.fc1 && a.fc2 && a.w7) || xp.p != a.x.p || xs.p != a.x.p || pool.n_img != a.n_img ||
      stats.n_img != a.n_img || mlp.n_img != a.n_img || pool.Cin != a.C || stats.Cin != a.C || mlp.Cin != a.C ||
      pool.Hi != a.H || pool.Wi != a.W || mlp.Hi != a.H || mlp.Wi != a.W)
    return 0;
  if (a.C % 8 != 0 || (a.C >> 3) > kFusedThreads || a.hidden < 1 || a.n_img > 65535) return 0;
  if (!(view_aligned(a.x.p, a.x.ld, a.x.img_stride, a.x.gstride) && view_aligned(a.y.p, a.y.ld, a.y.img_stride, a.y.gstride)))
    return 0;
  // cluster size: as many CTAs per image as keep the whole grid within one wave, bands of at least 3 rows
  int CS = 8;
  while (CS > 1 && ((int64_t)a.n_img * CS > sm_count() || ceil_div(a.H, CS) < 3)) CS >>= 1;
  {
    const char* e = getenv("DCFA_CBAM_CS");   // experiments: force the cluster size (more CTAs per image, several waves)
    if (e && atoi(e) >= 1) {
      CS = atoi(e);
      while (CS > 1 && ceil_div(a.H, CS) < 3) CS >>= 1;
    }
  }
  if (!force && (int64_t)a.n_img * CS < sm_count() / 4) return 0;   // too few CTAs to fill the GPU: keep the wide kernels
  a.rows_per = ceil_div(a.H, CS);
  const int c8n = a.C >> 3, planes = kFusedThreads / c8n, round_px = planes * kU;
  const size_t scratch = std::max((size_t)planes * 2 * a.C, (size_t)2 * round_px * (c8n + 1) * 2);
  const size_t floats = (size_t)5 * a.C + ((a.hidden + 3) & ~3) + 100 + (size_t)(a.rows_per + 6) * (a.W + 6) * 2 +
                        ((a.rows_per * a.W + 3) & ~3) + scratch;
  const size_t smem = floats * sizeof(float);
  if (smem > 200 * 1024) return 0;
  static DeviceOnce attr_set;
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(cbam_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail(DCFA_E_CUDA, "cbam_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.mark();
  }
  cudaError_t le = launch_k(cbam_fused_kernel, dim3((unsigned)(a.n_img * CS)), dim3(kFusedThreads), smem, st, CS, true, a);
  if (le != cudaSuccess) return fail(DCFA_E_CUDA, "cbam_fused: launch (cluster %d, smem %zu): %s", CS, smem, cudaGetErrorString(le));
  count_launch();
  return 1;
}

// The 19 records of SPPF_CBAM's attention / pooling sequence (CBAM x4 with a MAXPOOL5 between consecutive ones, each pool
// reading the previous CBAM's output): 1 = the resident cluster kernel was launched, 0 = left to the separate units.
int launch_sppf_fused(const dcfa_op* ops, void* const* bufs, cudaStream_t st) {
  bool force = false;
  {
    const char* e = getenv("DCFA_SPPF_FUSED");   // debug / tests: 0 keeps the seven separate launches, 2 fuses even tiny batches
    if (e && atoi(e) == 0) return 0;
    force = e && atoi(e) == 2;
  }
  SppfArgs a;
  a.stages = kSppfStages;
  const dcfa_op& ap0 = ops[3];
  a.n_img = ap0.n_img;
  a.group_imgs = ap0.group_imgs > 0 ? ap0.group_imgs : ap0.n_img;
  a.H = ap0.Hi; a.W = ap0.Wi; a.C = ap0.Cin; a.hidden = ops[1].hidden;
  a.x = resolve<const __nv_bfloat16>(ap0.x, bufs);
  for (int s = 0; s < kSppfStages; ++s) {
    const dcfa_op* q = ops + 5 * s;   // POOL, MLP, STATS, APPLY of stage s
    const dcfa_op &pool = q[0], &mlp = q[1], &stats = q[2], &apply = q[3];
    const int gi = apply.group_imgs > 0 ? apply.group_imgs : apply.n_img;
    if (pool.n_img != a.n_img || mlp.n_img != a.n_img || stats.n_img != a.n_img || apply.n_img != a.n_img || gi != a.group_imgs ||
        pool.Cin != a.C || mlp.Cin != a.C || stats.Cin != a.C || apply.Cin != a.C || apply.Hi != a.H || apply.Wi != a.W ||
        mlp.hidden != a.hidden)
      return 0;
    const View<const __nv_bfloat16> xa = resolve<const __nv_bfloat16>(apply.x, bufs), xp = resolve<const __nv_bfloat16>(pool.x, bufs),
                                    xs = resolve<const __nv_bfloat16>(stats.x, bufs);
    if (!xa.p || xp.p != xa.p || xs.p != xa.p) return 0;
    a.y[s] = resolve<__nv_bfloat16>(apply.y, bufs);
    a.fc1[s] = resolve_ptr<const float>(mlp.w, bufs);
    a.fc2[s] = resolve_ptr<const float>(mlp.scale, bufs);
    a.w7[s] = resolve_ptr<const float>(apply.w, bufs);
    if (!(a.y[s].p && a.fc1[s] && a.fc2[s] && a.w7[s])) return 0;
    if (!view_aligned(a.y[s].p, a.y[s].ld, a.y[s].img_stride, a.y[s].gstride)) return 0;
    if (s > 0) {   // the pool between stage s-1 and s: reads stage s-1's output, feeds stage s
      const dcfa_op& mp = ops[5 * s - 1];
      const View<const __nv_bfloat16> mx = resolve<const __nv_bfloat16>(mp.x, bufs);
      const View<__nv_bfloat16> my = resolve<__nv_bfloat16>(mp.y, bufs);
      if (mp.n_img != a.n_img || mp.Cin != a.C || mp.Hi != a.H || mp.Wi != a.W || mx.p != a.y[s - 1].p || mx.ld != a.y[s - 1].ld ||
          mx.img_stride != a.y[s - 1].img_stride || my.p != xa.p || my.ld != xa.ld)
        return 0;
    }
  }
  if (!view_aligned(a.x.p, a.x.ld, a.x.img_stride, a.x.gstride)) return 0;
  const int c8n = a.C >> 3;
  if (a.C % 8 != 0 || !(c8n == 8 || c8n == 16 || c8n == 32) || a.hidden < 1 || a.n_img > 65535) return 0;
  // cluster size: the smallest power of two (<= 8) whose band + pooling halo fit in shared memory, bands of >= 3 rows
  const size_t row_bytes = (size_t)a.W * a.C * 2;
  auto smem_for = [&](int cs) {
    const int rows = a.H / cs;
    const size_t floats = (size_t)5 * a.C + ((a.hidden + 3) & ~3) + 100 * kSppfStages + (size_t)(rows + 6) * (a.W + 6) * 2 +
                          ((rows * a.W + 3) & ~3);
    return (size_t)(rows + 4) * row_bytes + floats * sizeof(float);
  };
  int CS = 1;
  size_t smem = 0;
  for (;; CS <<= 1) {
    if (CS > 8 || a.H % CS != 0 || a.H / CS < 3) return 0;
    smem = smem_for(CS);
    if (smem <= 200 * 1024 && (force || (int64_t)a.n_img * CS >= sm_count() / 4)) break;
  }
  int nt = 1024;
  {
    const int cs2 = CS * 2;
    // experiment (DCFA_SPPF_NT=512): measured 65 us either way at s/B=32 -- the phases are latency chains, not occupancy
    const char* e = getenv("DCFA_SPPF_NT");
    const bool allow = e && atoi(e) == 512;
    if (allow && cs2 <= 8 && a.H % cs2 == 0 && a.H / cs2 >= 3 && smem_for(cs2) <= 100 * 1024 && (int64_t)a.n_img * cs2 <= 2 * sm_count()) {
      CS = cs2;
      smem = smem_for(CS);
      nt = 512;
    }
  }
  a.rows_per = a.H / CS;
  if ((int64_t)a.n_img * CS > 2 * sm_count()) return 0;
  static DeviceOnce attr_set;
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(sppf_cbam_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sppf_cbam_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return fail(DCFA_E_CUDA, "sppf_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.mark();
  }
  cudaError_t le = nt == 512 ? launch_k(sppf_cbam_kernel<512>, dim3((unsigned)(a.n_img * CS)), dim3(512), smem, st, CS, true, a)
                             : launch_k(sppf_cbam_kernel<1024>, dim3((unsigned)(a.n_img * CS)), dim3(1024), smem, st, CS, true, a);
  if (le != cudaSuccess) return fail(DCFA_E_CUDA, "sppf_fused: launch (cluster %d, smem %zu): %s", CS, smem, cudaGetErrorString(le));
  count_launch();
  return 1;
}

}  // namespace dcfa
