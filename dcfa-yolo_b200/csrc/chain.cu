// Fused ShuffleNetV2 branch 2 (nets/yolo_mul.py:138-162): 1x1 conv + BN + ReLU -> depthwise 3x3 + BN -> 1x1 conv + BN + ReLU
// as ONE kernel.  The plan still holds the three records (CONV, DWCONV, CONV); dcfa_run_ops recognises the chain
// and calls launch_chain, falling back to the three kernels for shapes this one does not take.
//
// At 160x160 / 80x80 the three kernels are DRAM traffic (each reads and writes the whole half tensor) plus three
// launches.  Here one CTA owns an 8 x 16 pixel tile of one image:
//   1. the 10 x 18 halo of the input arrives by TMA as the K-major A operand (hardware zero fill outside the image);
//   2. GEMM 1 (tcgen05, M = 2 x 128 halo rows, N = K = C) -> TMEM; the epilogue applies BN + ReLU, forces pixels
//      outside the image to zero (the depthwise conv pads its INPUT with zeros) and writes bf16 rows to shared memory;
//   3. the depthwise 3x3 (packed fp32x2 FMAs, register-tiled down a column) writes its bf16 result straight into
//      the swizzled K-major A tile of GEMM 2;
//   4. GEMM 2 -> TMEM -> BN + ReLU -> one 256-bit store per thread and 16 channels into the unit's output slot.
// Both weight matrices, the depthwise taps and the BN vectors stay resident in shared memory; the halo of the next
// tile is requested as soon as GEMM 1 has consumed the current one.  Intermediates are rounded to bf16 exactly where
// the three-kernel path stores them, so both paths produce the same values.
//
// The same kernel in GHOST mode runs a RepGhostModule in deploy algebra (nets/repghost.py:98-123): 1x1 conv + BN (+SiLU)
// -> depthwise 3x3 (identity-BN branch folded into the centre tap) (+SiLU) (+ the bottleneck's residual, :279) --
// steps 1-3, with the depthwise result stored straight to global memory instead of feeding a second GEMM.  The plan
// holds the two records (CONV with DCFA_CONV_FLAG_GHOST_HEAD, DWCONV); launch_ghost falls back to the two kernels.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {
namespace {

constexpr int TH = 8, TW = 16;                 // output tile
constexpr int HH = TH + 2, HW = TW + 2;        // halo
constexpr int NHALO = HH * HW;                 // 180 halo pixels = rows of GEMM 1 (padded to 256)
constexpr int kChainThreads = 256;

struct ChainArgs {
  const __nv_bfloat16* w1;   // [G][katoms][C x bk] packed conv tiles (pack_conv_weight, n_tiles == 1)
  const __nv_bfloat16* w2;
  int64_t w1_gstride, w2_gstride;
  const float *s1, *b1, *s2, *b2;   // BN scale / bias of the two 1x1 convs, [G][sb_gstride]
  int64_t sb1_gstride, sb2_gstride;
  const float* wd;           // [G][9][C] depthwise taps (BN folded)
  const float* bd;           // [G][C]
  View<__nv_bfloat16> y;
  int n_img, group_imgs, H, W, C;
  int act1, actd, act2;
  int bk, katoms;            // K atom width (64 / 32 channels) and C / bk
  int tiles_x, tiles_y;
  uint32_t row_bytes;        // bk * 2
  uint32_t layout_type, sbo; // UMMA descriptor swizzle code and 8-row group stride
  uint32_t tmem_cols;
  uint32_t off_a2, off_w1, off_w2, off_t1, off_par, off_bar;   // shared-memory offsets from the 1024-aligned base
  uint32_t t1_pitch;         // bytes per row of the GEMM-1 result tile
  int ghost;                 // 1: no second GEMM -- the depthwise result (+ residual) is the output
  View<const __nv_bfloat16> res;   // ghost mode: optional residual, added after the depthwise activation
};

// 16 consecutive channels of one row: act(acc * scale + bias) -> two 128-bit words of bf16.  128-bit shared-memory loads
// of the BN vectors, packed fp32x2 FMAs (same IEEE results, half the issue slots), the (uniform) activation branch taken
// once per chunk.  ReLU is folded into the conversion (cvt.rn.relu.bf16x2: the same values as max(x, 0) then round);
// for SiLU the caller stages scale and bias HALVED (exact), so the FMA yields h = x / 2 and x * sigmoid(x) = h + h * tanh(h).
__device__ __forceinline__ void bn_act16(const uint32_t (&acc)[16], const float* sc, const float* bi, int act, uint4& lo, uint4& hi) {
  F2 v2[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 s4 = *reinterpret_cast<const float4*>(sc + 4 * q);
    const float4 b4 = *reinterpret_cast<const float4*>(bi + 4 * q);
    v2[2 * q] = f2_make(b4.x, b4.y);
    v2[2 * q + 1] = f2_make(b4.z, b4.w);
    f2_fma(v2[2 * q], f2_make(__uint_as_float(acc[4 * q + 0]), __uint_as_float(acc[4 * q + 1])), f2_make(s4.x, s4.y));
    f2_fma(v2[2 * q + 1], f2_make(__uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3])), f2_make(s4.z, s4.w));
  }
  uint32_t w[8];
  if (act == DCFA_ACT_RELU) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a, b;
      f2_get(v2[e], a, b);
      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(w[e]) : "f"(b), "f"(a));
    }
  } else if (act == DCFA_ACT_SILU) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float h0, h1, t0, t1;
      f2_get(v2[e], h0, h1);
      asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
      asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
      F2 h = v2[e];
      f2_fma(h, h, f2_make(t0, t1));   // h * t + h
      f2_get(h, h0, h1);
      w[e] = pack_bf16x2(h0, h1);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a, b;
      f2_get(v2[e], a, b);
      w[e] = pack_bf16x2(a, b);
    }
  }
  lo = make_uint4(w[0], w[1], w[2], w[3]);
  hi = make_uint4(w[4], w[5], w[6], w[7]);
}

__device__ __forceinline__ void chain_tma_load(uint32_t dst, const CUtensorMap* map, int c, int x, int y, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c), "r"(x), "r"(y), "r"(n), "r"(bar)
      : "memory");
}

// Depthwise 3x3 over the GEMM-1 tile (same register tiling as dw_tile_compute in dwconv.cu): one thread = one
// 8-channel chunk of one tile column and R consecutive output rows, for the 2R chunks starting at c8_base; the bf16
// results go straight into the swizzled K-major A tile of GEMM 2.
// Tile of the output tensor in ghost mode: pointer to pixel (0, 0) of the tile (channel 0), rows / columns inside the image
struct GhostTile {
  __nv_bfloat16* y;
  const __nv_bfloat16* res;
  int rows, cols;
  int64_t y_row, res_row;   // elements per image row
};

template <int R, int CT>
__device__ __forceinline__ void chain_dw(const ChainArgs& p, const uint8_t* t1, uint32_t s_a2, const float* wd_s, const float* bd_s,
                                         int c8_base, int tid, const GhostTile& gt) {
  constexpr int C8N = 2 * R;                 // chunks handled per pass: 4 (R = 2) or 8 (R = 4)
  constexpr int RG = TH / R;                 // row groups
  // Thread mapping.  Chain: columns fastest, so a warp works on ONE chunk -- its tap loads are broadcasts (1 shared-memory
  // wavefront instead of 4-8: the taps were the largest consumer of shared-memory bandwidth) and its T1 reads / A-tile
  // stores walk neighbouring pixels, conflict-free in both layouts.  Ghost: chunks fastest, so that the lanes of a quarter
  // warp store the contiguous channels of one pixel to global memory (taps padded to 48 bytes per chunk: no bank conflict).
  const int c8 = c8_base + (p.ghost ? tid % C8N : tid / (TW * RG));
  const int col = p.ghost ? (tid / C8N) % TW : tid % TW;
  const int oy0 = (p.ghost ? tid / (C8N * TW) : (tid / TW) % RG) * R;
  constexpr int C = CT;
  constexpr int WTAP = (C / 8) * 12;         // floats per tap: 8 weights + 4 floats of padding per chunk
  F2 acc[R][4];
  {
    const float4 b0 = *reinterpret_cast<const float4*>(bd_s + c8 * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(bd_s + c8 * 8 + 4);
#pragma unroll
    for (int o = 0; o < R; ++o) {
      acc[o][0] = f2_make(b0.x, b0.y); acc[o][1] = f2_make(b0.z, b0.w);
      acc[o][2] = f2_make(b1.x, b1.y); acc[o][3] = f2_make(b1.z, b1.w);
    }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    F2 w[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float4 w0 = *reinterpret_cast<const float4*>(wd_s + (r * 3 + q) * WTAP + c8 * 12);
      const float4 w1 = *reinterpret_cast<const float4*>(wd_s + (r * 3 + q) * WTAP + c8 * 12 + 4);
      w[r][0] = f2_make(w0.x, w0.y); w[r][1] = f2_make(w0.z, w0.w);
      w[r][2] = f2_make(w1.x, w1.y); w[r][3] = f2_make(w1.z, w1.w);
    }
#pragma unroll
    for (int ri = 0; ri < R + 2; ++ri) {
      F2 v[4];
      const int hp = (oy0 + ri) * HW + col + q;   // halo pixel
      // C = 32: unpadded 64-byte rows with the chunk index XOR-swizzled by the row pair (conflict-free for the row-wise
      // writes of epilogue 1 AND for these reads, where a quarter warp covers 4 chunks of 2 neighbouring pixels)
      const uint32_t t1_off = C == 32 ? (uint32_t)hp * 64u + (uint32_t)((c8 ^ (hp >> 1)) & 3) * 16u : (uint32_t)hp * p.t1_pitch + (uint32_t)c8 * 16u;
      unpack8_f2(*reinterpret_cast<const uint4*>(t1 + t1_off), v);
#pragma unroll
      for (int o = 0; o < R; ++o) {
        const int r = ri - o;
        if (r >= 0 && r < 3) {
#pragma unroll
          for (int e = 0; e < 4; ++e) f2_fma(acc[o][e], v[e], w[r][e]);
        }
      }
    }
  }
  const int ka = (c8 * 8) / p.bk;
  const uint32_t ch = (uint32_t)(((c8 * 8) % p.bk) >> 3);
#pragma unroll
  for (int o = 0; o < R; ++o) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) f2_get(acc[o][e], f[2 * e], f[2 * e + 1]);
    if (p.actd != DCFA_ACT_NONE) {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = apply_act(f[e], p.actd);
    }
    if (p.ghost) {   // the module's output: (+ residual) -> 16 bytes of one pixel, neighbouring threads cover the neighbouring chunks
      if (oy0 + o < gt.rows && col < gt.cols) {
        if (gt.res) {
          float rr[8];
          unpack8(ldg128(gt.res + (int64_t)(oy0 + o) * gt.res_row + (int64_t)col * p.res.ld + c8 * 8), rr);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += rr[e];
        }
        stg128(gt.y + (int64_t)(oy0 + o) * gt.y_row + (int64_t)col * p.y.ld + c8 * 8, pack8(f));
      }
      continue;
    }
    const uint4 v = pack8(f);
    const uint32_t pp = (uint32_t)((oy0 + o) * TW + col);
    const uint32_t swz = p.bk == 64 ? (pp & 7u) : ((pp >> 1) & 3u);
    const uint32_t addr = s_a2 + (uint32_t)ka * 128u * p.row_bytes + pp * p.row_bytes + ((ch ^ swz) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
}

// CT = channel count (compile time: loop bounds, shifts and the register tile follow it); CTAs per SM by CT
template <int CT>
__global__ void __launch_bounds__(kChainThreads, CT == 32 ? 3 : (CT == 64 ? 2 : 1)) chain_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                                             const ChainArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gb = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t s_xa = base, s_a2 = base + p.off_a2, s_w1 = base + p.off_w1, s_w2 = base + p.off_w2;
  uint8_t* t1 = gb + p.off_t1;
  constexpr int C = CT;
  float* wd_s = reinterpret_cast<float*>(gb + p.off_par);   // [9][C]
  float* bd_s = wd_s + 9 * (C / 8) * 12;   // taps padded to 48 bytes per chunk
  float* s1_s = bd_s + C;
  float* b1_s = s1_s + C;
  float* s2_s = b1_s + C;
  float* b2_s = s2_s + C;
  const uint32_t bar_x = base + p.off_bar, bar_mma = bar_x + 8u, tmem_slot = bar_mma + 8u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gb + p.off_bar + 16u);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_init(bar_x, 1);
      ptx::mbar_init(bar_mma, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (tid == 32) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
  ptx::pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  ptx::pdl_wait();

  const uint32_t idesc = ptx::make_idesc_bf16_f32(128, C);
  const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(p.sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)p.layout_type << 61);
  const int tiles_img = p.tiles_x * p.tiles_y;
  const int total = p.n_img * tiles_img;
  const int ksteps = p.bk >> 4;
  const uint32_t xa_atom = 256u * p.row_bytes, a2_atom = 128u * p.row_bytes, w_atom = (uint32_t)C * p.row_bytes;

  auto issue_x = [&](int tile) {   // one thread: the halo box(es) of `tile`
    const int n = tile / tiles_img;
    const int r = tile - n * tiles_img;
    const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
    ptx::mbar_arrive_expect_tx(bar_x, (uint32_t)p.katoms * NHALO * p.row_bytes);
    for (int ka = 0; ka < p.katoms; ++ka)
      chain_tma_load(s_xa + (uint32_t)ka * xa_atom, &map_x, ka * p.bk, tx * TW - 1, ty * TH - 1, n, bar_x);
  };

  int tile = blockIdx.x;
  if (tid == 32 && tile < total) issue_x(tile);
  int cur_g = -1;
  uint32_t xph = 0, mph = 0;
  for (; tile < total; tile += gridDim.x) {
    const int n = tile / tiles_img;
    const int trem = tile - n * tiles_img;
    const int ty = trem / p.tiles_x, tx = trem - ty * p.tiles_x;
    int g = 0;
    for (int nn = n; nn >= p.group_imgs; nn -= p.group_imgs) ++g;
    if (g != cur_g) {   // this modality's weights, taps and BN vectors (the previous tile ended with a barrier)
      const uint4* src1 = reinterpret_cast<const uint4*>(p.w1 + (int64_t)g * p.w1_gstride);
      const uint4* src2 = reinterpret_cast<const uint4*>(p.w2 + (int64_t)g * p.w2_gstride);
      uint4* d1 = reinterpret_cast<uint4*>(gb + p.off_w1);
      uint4* d2 = reinterpret_cast<uint4*>(gb + p.off_w2);
      const int n16 = C * C * 2 / 16;
      for (int i = tid; i < n16; i += kChainThreads) d1[i] = __ldg(src1 + i);
      if (!p.ghost)
        for (int i = tid; i < n16; i += kChainThreads) d2[i] = __ldg(src2 + i);
      for (int i = tid; i < 9 * C; i += kChainThreads)   // [tap][chunk][8 + 4 pad]
        wd_s[(i / C) * (C / 8) * 12 + ((i % C) >> 3) * 12 + (i & 7)] = __ldg(p.wd + (int64_t)g * 9 * C + i);
      for (int i = tid; i < C; i += kChainThreads) {
        bd_s[i] = __ldg(p.bd + (int64_t)g * C + i);
        const float pre1 = p.act1 == DCFA_ACT_SILU ? 0.5f : 1.0f, pre2 = p.act2 == DCFA_ACT_SILU ? 0.5f : 1.0f;   // see bn_act16
        s1_s[i] = pre1 * __ldg(p.s1 + (int64_t)g * p.sb1_gstride + i);
        b1_s[i] = pre1 * __ldg(p.b1 + (int64_t)g * p.sb1_gstride + i);
        if (!p.ghost) {
          s2_s[i] = pre2 * __ldg(p.s2 + (int64_t)g * p.sb2_gstride + i);
          b2_s[i] = pre2 * __ldg(p.b2 + (int64_t)g * p.sb2_gstride + i);
        }
      }
      cur_g = g;
      __syncthreads();
    }

    // ---- GEMM 1: [256 halo rows, C] x W1^T -> TMEM columns [0, C) (rows 0..127) and [C, 2C) (rows 128..255)
    if (tid == 0) {
      ptx::mbar_wait(bar_x, xph);
      ptx::fence_proxy_async_smem();   // the weight tiles were written with generic stores
      ptx::tc_fence_after();
      for (int mb = 0; mb < 2; ++mb)
        for (int ka = 0; ka < p.katoms; ++ka)
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t a_addr = s_xa + (uint32_t)ka * xa_atom + (uint32_t)mb * 128u * p.row_bytes + (uint32_t)k * 32u;
            const uint32_t b_addr = s_w1 + (uint32_t)ka * w_atom + (uint32_t)k * 32u;
            ptx::umma_bf16(tmem_base + (uint32_t)(mb * C), desc_hi | (uint64_t)((a_addr & 0x3FFFFu) >> 4),
                           desc_hi | (uint64_t)((b_addr & 0x3FFFFu) >> 4), idesc, (ka | k) ? 1u : 0u);
          }
      ptx::umma_commit(bar_mma);
    }
    xph ^= 1u;
    ptx::mbar_wait(bar_mma, mph);
    mph ^= 1u;
    ptx::tc_fence_after();
    if (tid == 32) {   // the halo tile has been consumed: request the next one (overlaps everything below)
      const int nxt = tile + (int)gridDim.x;
      if (nxt < total) issue_x(nxt);
    }

    // ---- epilogue 1: BN + act, zero outside the image, bf16 rows -> T1
    {
      const int mb = warp >> 2, q4 = warp & 3;
      const int r = mb * 128 + q4 * 32 + lane;     // halo row
      const int hy = r / HW, hx = r - hy * HW;
      const int iy = ty * TH - 1 + hy, ix = tx * TW - 1 + hx;
      const bool row_ok = r < NHALO;
      const bool inside = row_ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
      const uint32_t taddr = tmem_base + (uint32_t)(mb * C) + ((uint32_t)(q4 * 32) << 16);
      uint8_t* dst = t1 + (size_t)r * p.t1_pitch;
      for (int j = 0; j < (C >> 4); ++j) {
        uint32_t acc[16];
        ptx::tmem_ld_x16(taddr + (uint32_t)(j * 16), acc);
        ptx::tmem_ld_wait();
        uint4 lo, hi;
        bn_act16(acc, s1_s + j * 16, b1_s + j * 16, p.act1, lo, hi);
        if (row_ok) {   // pixels outside the image are the depthwise conv's zero padding
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
          const uint32_t sw = C == 32 ? (uint32_t)(r >> 1) & 3u : 0u;
          *reinterpret_cast<uint4*>(dst + (((uint32_t)(2 * j) ^ sw) << 4)) = inside ? lo : z;
          *reinterpret_cast<uint4*>(dst + (((uint32_t)(2 * j + 1) ^ sw) << 4)) = inside ? hi : z;
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();   // T1 complete; the GEMM-1 accumulators have been read

    // ---- depthwise 3x3 -> A tile of GEMM 2 (ghost mode: -> global memory, the module's output)
    GhostTile gt;
    gt.y = nullptr; gt.res = nullptr; gt.rows = gt.cols = 0; gt.y_row = gt.res_row = 0;
    if (p.ghost) {
      gt.y_row = (int64_t)p.W * p.y.ld;
      gt.y = p.y.p + p.y.img_off(n) + (int64_t)(ty * TH) * gt.y_row + (int64_t)(tx * TW) * p.y.ld;
      gt.rows = p.H - ty * TH;
      gt.cols = p.W - tx * TW;
      if (p.res.p) {
        gt.res_row = (int64_t)p.W * p.res.ld;
        gt.res = p.res.p + p.res.img_off(n) + (int64_t)(ty * TH) * gt.res_row + (int64_t)(tx * TW) * p.res.ld;
      }
    }
    if (C == 32) {
      chain_dw<2, CT>(p, t1, s_a2, wd_s, bd_s, 0, tid, gt);
    } else {
#pragma unroll
      for (int cb = 0; cb < (C >> 3); cb += 8) chain_dw<4, CT>(p, t1, s_a2, wd_s, bd_s, cb, tid, gt);
    }
    __syncthreads();
    if (p.ghost) continue;   // T1 may be rewritten by the next tile's epilogue 1

    // ---- GEMM 2: [128 pixels, C] x W2^T -> TMEM columns [0, C)
    if (tid == 0) {
      ptx::fence_proxy_async_smem();   // the A tile was written with generic stores by the other threads
      ptx::tc_fence_after();
      for (int ka = 0; ka < p.katoms; ++ka)
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t a_addr = s_a2 + (uint32_t)ka * a2_atom + (uint32_t)k * 32u;
          const uint32_t b_addr = s_w2 + (uint32_t)ka * w_atom + (uint32_t)k * 32u;
          ptx::umma_bf16(tmem_base, desc_hi | (uint64_t)((a_addr & 0x3FFFFu) >> 4), desc_hi | (uint64_t)((b_addr & 0x3FFFFu) >> 4),
                         idesc, (ka | k) ? 1u : 0u);
        }
      ptx::umma_commit(bar_mma);
    }
    ptx::mbar_wait(bar_mma, mph);
    mph ^= 1u;
    ptx::tc_fence_after();

    // ---- epilogue 2: BN + act -> one 256-bit store per 16 channels; the two warp halves split the channel chunks
    {
      const int q4 = warp & 3, half = warp >> 2;
      const int r = q4 * 32 + lane;
      const int oy = ty * TH + r / TW, ox = tx * TW + (r & (TW - 1));
      const bool valid = oy < p.H && ox < p.W;
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
      __nv_bfloat16* yrow = p.y.p + p.y.img_off(n) + (int64_t)(oy * p.W + ox) * p.y.ld;
      const int nch = C >> 4;
      for (int j = half * (nch >> 1); j < (half + 1) * (nch >> 1); ++j) {
        uint32_t acc[16];
        ptx::tmem_ld_x16(taddr + (uint32_t)(j * 16), acc);
        ptx::tmem_ld_wait();
        uint4 lo, hi;
        bn_act16(acc, s2_s + j * 16, b2_s + j * 16, p.act2, lo, hi);
        if (valid) {
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yrow + j * 16), "r"(lo.x), "r"(lo.y),
                       "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                       : "memory");
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();   // accumulators read; T1 and the A tile may be rewritten
  }

  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}


// Two second-generation designs of this unit were built on top of it in round 2, passed every parity test, were measured
// on B200 (s, B=32) and removed again (profiles/README.md, round 2, items 2 and 8):
//   * the depthwise stage on the TENSOR pipe (T1 as channel planes, nine diagonal M128 x N16 x K16 MMAs per 16-channel block
//     over shifted no-swizzle views): an MMA costs ~47 cycles for every N <= 64 (tools/umma_nsweep_test.cu), so the tensor
//     pipe became the bottleneck -- 0.187 / 0.165 ms against 0.161 / 0.078 ms here at C = 32 / 64;
//   * a warp-specialised PIPELINE (producer, control, 4 epilogue-1 warps, 8 depthwise warps, 4 epilogue-2 warps, every
//     hand-over double buffered, 12 x 16 tiles): bit-identical, 0.170 / 0.096 ms.  ncu: every role waits on its input
//     barrier 11-15 % of the samples; one tile takes ~11 000 cycles end to end in both designs, this kernel keeps three
//     tiles in flight per SM (three CTAs), the pipeline two.  The unit is bound by its instruction count (7 400 warp
//     instructions per 128-pixel tile, 2.0 IPC per SM), not by phase serialisation.

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn chain_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

bool same_view(const dcfa_view& a, const dcfa_view& b) {
  return a.buf == b.buf && a.off == b.off && a.ld == b.ld && a.img_stride == b.img_stride && a.gi == b.gi && a.gstride == b.gstride;
}

}  // namespace

namespace {
int launch_chain_impl(const dcfa_op& pw1, const dcfa_op& dw, const dcfa_op* pw2p, void* const* bufs, cudaStream_t st);
}

// pw1 (CONV 1x1), dw (DWCONV), pw2 (CONV 1x1): returns 1 if the fused kernel was launched, 0 if the chain is left to
// the three kernels, < 0 on error.  Only called for chains the plan marked as private (DCFA_CONV_FLAG_CHAIN_HEAD).
int launch_chain(const dcfa_op& pw1, const dcfa_op& dw, const dcfa_op& pw2, void* const* bufs, cudaStream_t st) {
  return launch_chain_impl(pw1, dw, &pw2, bufs, st);
}

// RepGhostModule: pw (CONV 1x1 with DCFA_CONV_FLAG_GHOST_HEAD), dw (DWCONV, optional residual): 1 = fused kernel launched,
// 0 = left to the two kernels, < 0 on error.
int launch_ghost(const dcfa_op& pw, const dcfa_op& dw, void* const* bufs, cudaStream_t st) {
  return launch_chain_impl(pw, dw, nullptr, bufs, st);
}

namespace {
int launch_chain_impl(const dcfa_op& pw1, const dcfa_op& dw, const dcfa_op* pw2p, void* const* bufs, cudaStream_t st) {
  const bool ghost = pw2p == nullptr;
  const dcfa_op& pw2 = ghost ? pw1 : *pw2p;   // ghost mode: the checks on the second conv degenerate to the first
  {
    const char* e = getenv(ghost ? "DCFA_GHOST" : "DCFA_CHAIN");   // debug: =0 keeps the separate kernels
    if (e && atoi(e) == 0) return 0;
  }
  bool force = false;
  {
    const char* e = getenv(ghost ? "DCFA_GHOST" : "DCFA_CHAIN");
    force = e && atoi(e) == 2;   // tests: also take the shapes the heuristic below leaves to the separate kernels
  }
  const int C = pw1.Cin;
  if (!(C == 32 || C == 64 || C == 128)) return 0;
  // measured (s, B=32): 0.172 vs 0.188 ms at C=32/160^2, 0.074 vs 0.090 at C=64/80^2, but 0.065 vs 0.062 at C=128/40^2,
  // where one CTA per SM (217 KB of shared memory) no longer hides the phase latencies
  // ghost mode (s, B=32): 0.041 vs 0.042 / 0.045 vs 0.050 ms at C=64/80^2, but 0.041 vs 0.032 at C=128/40^2
  if (pw1.ksize != 1 || pw2.ksize != 1 || pw1.stride != 1 || pw2.stride != 1 || pw1.Cout != C || pw2.Cin != C || pw2.Cout != C ||
      dw.Cin != C || pw1.n_tiles != 1 || pw2.n_tiles != 1 || pw1.BN != C || pw2.BN != C)
    return 0;
  const int bk = pw1.flags & 0xff;
  if (bk != (pw2.flags & 0xff) || (bk != 64 && bk != 32) || C % bk != 0 || pw1.k_blocks != C / bk || pw2.k_blocks != C / bk) return 0;
  if (pw1.out_mode != DCFA_OUT_BF16_NHWC || pw2.out_mode != DCFA_OUT_BF16_NHWC || pw1.x2.buf >= 0 || pw2.x2.buf >= 0 ||
      (!ghost && dw.x2.buf >= 0) || pw1.parts != 0 || pw2.parts != 0 || pw1.f0 != 1.0f || pw2.f0 != 1.0f)
    return 0;
  if (!same_view(pw1.y, dw.x) || (!ghost && !same_view(dw.y, pw2.x))) return 0;
  if (pw1.n_img != dw.n_img || pw1.n_img != pw2.n_img || pw1.Hi != dw.Hi || pw1.Wi != dw.Wi || pw2.Hi != dw.Hi || pw2.Wi != dw.Wi ||
      pw1.group_imgs != dw.group_imgs || pw1.group_imgs != pw2.group_imgs)
    return 0;

  ChainArgs a;
  View<const __nv_bfloat16> x = resolve<const __nv_bfloat16>(pw1.x, bufs);
  a.y = resolve<__nv_bfloat16>(ghost ? dw.y : pw2.y, bufs);
  a.ghost = ghost ? 1 : 0;
  a.res = resolve<const __nv_bfloat16>(dw.x2, bufs);
  if (!ghost) a.res.p = nullptr;
  a.w1 = resolve_ptr<const __nv_bfloat16>(pw1.w, bufs);
  a.w2 = resolve_ptr<const __nv_bfloat16>(pw2.w, bufs);
  a.w1_gstride = pw1.w_gstride; a.w2_gstride = pw2.w_gstride;
  a.s1 = resolve_ptr<const float>(pw1.scale, bufs); a.b1 = resolve_ptr<const float>(pw1.bias, bufs);
  a.s2 = resolve_ptr<const float>(pw2.scale, bufs); a.b2 = resolve_ptr<const float>(pw2.bias, bufs);
  a.sb1_gstride = pw1.sb_gstride; a.sb2_gstride = pw2.sb_gstride;
  a.wd = resolve_ptr<const float>(dw.w, bufs);
  a.bd = resolve_ptr<const float>(dw.bias, bufs);
  a.n_img = pw1.n_img;
  a.group_imgs = pw1.group_imgs > 0 ? pw1.group_imgs : pw1.n_img;
  a.H = pw1.Hi; a.W = pw1.Wi; a.C = C;
  a.act1 = pw1.act; a.actd = dw.act; a.act2 = pw2.act;
  if (!(x.p && a.y.p && a.w1 && a.w2 && a.s1 && a.b1 && a.s2 && a.b2 && a.wd && a.bd)) return 0;
  if (x.gi > 0 && x.gstride != (int64_t)x.gi * x.img_stride) return 0;   // grouped input views: not through one tensor map
  if (((uintptr_t)x.p % 16) != 0 || x.ld % 8 != 0 || x.img_stride % 8 != 0) return 0;
  if (((uintptr_t)a.y.p % 32) != 0 || a.y.ld % 16 != 0 || a.y.img_stride % 16 != 0 || a.y.gstride % 16 != 0) return 0;
  if (a.res.p && (((uintptr_t)a.res.p % 16) != 0 || a.res.ld % 8 != 0 || a.res.img_stride % 8 != 0 || a.res.gstride % 8 != 0)) return 0;
  if (((uintptr_t)a.w1 % 16) != 0 || ((uintptr_t)a.w2 % 16) != 0 || a.w1_gstride % 8 != 0 || a.w2_gstride % 8 != 0) return 0;

  a.bk = bk;
  a.katoms = C / bk;
  a.row_bytes = (uint32_t)bk * 2u;
  a.layout_type = bk == 64 ? 2u : 4u;
  a.sbo = 8u * a.row_bytes;

  EncodeTiledFn enc = chain_encode_fn();
  if (!enc) return fail(DCFA_E_CUDA, "chain: cuTensorMapEncodeTiled entry point unavailable");
  const int64_t run = (C == x.ld) ? (int64_t)C * 2 * a.W : (int64_t)C * 2;
  const CUtensorMapL2promotion promo = run >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : (run >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B);
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.n_img};
  const cuuint64_t gstr[3] = {(cuuint64_t)x.ld * 2, (cuuint64_t)a.W * x.ld * 2, (cuuint64_t)x.img_stride * 2};
  const cuuint32_t es[4] = {1u, 1u, 1u, 1u};

  if (C == 128 && !force) return 0;
  a.tiles_x = ceil_div(a.W, TW);
  a.tiles_y = ceil_div(a.H, TH);
  const int64_t total = (int64_t)a.n_img * a.tiles_x * a.tiles_y;
  if (total >= (1ll << 31)) return 0;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * C)) cols <<= 1;
  a.tmem_cols = cols;

  // shared memory: XA | A2 | W1 | W2 | T1 | parameters | barriers
  uint32_t off = (uint32_t)a.katoms * 256u * a.row_bytes;
  a.off_a2 = off; off += ghost ? 0u : (uint32_t)a.katoms * 128u * a.row_bytes;
  a.off_w1 = off; off += (uint32_t)C * C * 2u;
  a.off_w2 = off; off += ghost ? 0u : (uint32_t)C * C * 2u;
  a.t1_pitch = C == 32 ? 64u : (uint32_t)C * 2u + 16u;   // C = 32: swizzled instead of padded (see chain_dw)
  a.off_t1 = off; off += ((uint32_t)NHALO * a.t1_pitch + 127u) & ~127u;
  a.off_par = off; off += (uint32_t)(9 * (C / 8) * 12 + 5 * C) * 4u;
  a.off_bar = (off + 15u) & ~15u;
  const size_t smem = 1024 + a.off_bar + 64;
  if (smem > 227 * 1024) return 0;

  alignas(64) CUtensorMap map;
  const cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)HW, (cuuint32_t)HH, 1u};
  CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x.p), gdim, gstr, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, promo,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(DCFA_E_CUDA, "chain: cuTensorMapEncodeTiled failed with %d", (int)cr);

  static DeviceOnce attr_set;
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(chain_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chain_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(chain_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(DCFA_E_CUDA, "chain: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.mark();
  }
  const int ctas = C == 32 ? 3 : (C == 64 ? 2 : 1);   // matches the kernel's launch bounds; shared memory allows it
  int64_t grid = (int64_t)sm_count() * ctas;
  if (grid > total) grid = total;
  const dim3 g3((unsigned)grid), b3(kChainThreads);
  cudaError_t le = C == 32 ? launch_pdl(chain_kernel<32>, g3, b3, smem, st, map, a)
                           : (C == 64 ? launch_pdl(chain_kernel<64>, g3, b3, smem, st, map, a)
                                      : launch_pdl(chain_kernel<128>, g3, b3, smem, st, map, a));
  if (le != cudaSuccess) return fail(DCFA_E_CUDA, "chain: launch: %s", cudaGetErrorString(le));
  le = cudaGetLastError();
  if (le != cudaSuccess) return fail(DCFA_E_CUDA, "chain_kernel launch failed: %s", cudaGetErrorString(le));
  count_launch();
  return 1;
}
}  // namespace

}  // namespace dcfa
