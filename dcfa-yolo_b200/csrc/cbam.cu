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
constexpr int kFusedThreads = 512;

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

}  // namespace dcfa
