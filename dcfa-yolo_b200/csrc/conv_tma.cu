// DCFA_OP_CONV, TMA path: implicit-GEMM convolution whose A operand is fetched by the Tensor Memory
// Accelerator instead of per-thread cp.async (ncu, round 1: the cp.async gather saturated at ~8 B/cycle/SM).
//
// The NHWC bf16 input view is described once per launch by a rank-4 tensor map (C, W, H, N).  An M tile is a
// th x tw rectangle of output pixels of ONE image; the A tile of filter tap (dy, dx) and channel block cb is then
// a single box load at coordinates (cb*BK, x0*s - pad + dx, y0*s - pad + dy, n):
//   * out-of-image taps and partial tiles are zero-filled by the TMA (conv zero padding for free);
//   * stride-2 convolutions use the map's elementStrides = 2 traversal;
//   * the box lands in shared memory already in the K-major swizzled layout tcgen05.mma reads
//     (SWIZZLE_128B / 64B / 32B for BK = 64 / 32 / 16 channels per stage).
// One thread issues, per k-block, one cp.async.bulk.tensor (A) and one cp.async.bulk (pre-swizzled W tile);
// both complete on the stage's mbarrier through transaction bytes.  One thread issues tcgen05.mma into a
// double-buffered TMEM accumulator; two groups of four epilogue warps (one per accumulator stage) apply the
// fused epilogue (folded-BN scale/bias, ReLU/SiLU, post-scale, residual) and store bf16 NHWC at a channel
// offset or fp32 NCHW.  Persistent CTAs, 10 warps:
//   warps 0-3 epilogue group 0, warp 4 TMA producer, warp 5 MMA issuer + TMEM alloc, warps 6-9 epilogue group 1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {

namespace {

constexpr int BM = 128;
constexpr int kTmaWarp = 4;
constexpr int kMmaWarp = 5;
constexpr int kThreads = 320;       // generic instance: two epilogue groups
#ifndef DCFA_FAST_GROUPS
#define DCFA_FAST_GROUPS 2   // four groups (96 registers per thread) measured 0.8 % slower per step than two (130)
#endif
constexpr int kFastGroups = DCFA_FAST_GROUPS;             // FAST instance: epilogue groups (warps 0-3, 6-9, 10-13, 14-17)
constexpr int kThreadsFast = 64 + 128 * kFastGroups;
constexpr int kMaxStages = 24;

struct FastDiv {
  uint32_t d, mul, shr;
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return mul ? (__umulhi(n, mul) >> shr) : n; }
};

FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{d, 0u, 0u};
  if (d > 1) {
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    const uint32_t p = 31 + l;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
  }
  return f;
}

struct TmaConvArgs {
  View<const __nv_bfloat16> res;
  View<void> y;
  View<__nv_bfloat16> y2;  // split output: channels >= split go to this view (channel 0 of y2 = channel `split`)
  int split;               // 0 = single destination
  const __nv_bfloat16* w;
  const float* scale;
  const float* bias;
  int64_t w_gstride;
  int64_t sb_gstride;
  int n_img, group_imgs;
  int Ho, Wo, Cout, ksize, stride, pad;
  int pair;               // DCFA_CONV_FLAG_PAIR: six pixel-pair k-blocks per tile (3x3 stride 2, Cin = 32), see launch_conv_tma
  int BN, n_tiles, k_blocks;
  int bk;                 // channels per k-block (64 / 32 / 16)
  int cblocks;            // Cin / bk
  int act, out_mode, out_ctot, out_coff;
  float post_scale;
  int tw, th;             // spatial M tile (tw*th <= 128)
  int tiles_x, tiles_img; // tiles per image row / per image
  int total_tiles;
  int stages;
  uint32_t a_stage_bytes, b_stage_bytes, a_tx_bytes, b_tx_bytes;
  uint32_t layout_type;   // UMMA descriptor swizzle code (2 / 4 / 6)
  uint32_t sbo;           // 8 rows * row pitch
  FastDiv div_tw, div_tiles_img, div_tiles_x, div_ntiles;
  int tma_store;          // bf16 NHWC output written with TMA stores from a swizzled smem staging tile
  int st256;              // bf16 NHWC output written with one 256-bit store per thread and 16-channel chunk: a whole
                          // 32-byte sector per instruction, no staging, no barriers inside the epilogue group
  int cbox;               // channels per store slab (64 / 32 / 16)
  uint32_t out_stage_bytes;  // 128 rows * cbox * 2
  uint32_t tmem_cols;
  uint32_t acc_stages;   // TMEM accumulator stages (2 or 4): the MMA issuer runs this many tiles ahead of the epilogues
  // DCFA_CONV_FLAG_DFL (fp32 NCHW head map [box 64 | cls nc]): the epilogue also emits DFL(box) and the class logits
  float* dfl_dbox;       // [n_img, 4, A] (nullptr: off)
  float* dfl_cls;        // [n_img, nc, A]
  int dfl_A, dfl_aoff, dfl_nc;   // total anchors, first anchor of this level, classes
  int mc;                // 1: CTA PAIRS (cluster of 2) on two M tiles of the same weight tile: each CTA fetches half of
                         // every W k-block and multicasts it into both CTAs' rings (W is 40-67 % of the bytes a k-block
                         // pulls through L2 -> SM, the delivery limit of the N >= 128 layers)
  int iters;             // loop range of the persistent roles: total_tiles, or the number of tile PAIRS with mc
  int ngroups;           // epilogue groups of four warps: 2 (generic instance) or 4 (FAST instance)
  int epi_split;         // 1: BOTH epilogue groups work on every tile, each on every other 16-channel chunk (wide
                         // tiles: halves the epilogue latency of a tile, which is exposed at the tail of every launch
                         // and is all there is when a CTA gets one tile); 0: the groups alternate tiles
};

#ifdef DCFA_TIMELINE
// debug build only (make EXTRA=-DDCFA_TIMELINE): clock64 timestamps of CTA 0, read back by tools/timeline.py
__device__ long long g_tl[8][2048];
#define TL(role, idx) do { if (blockIdx.x == 0 && (idx) < 2048) g_tl[role][idx] = clock64(); } while (0)
#else
#define TL(role, idx) do { } while (0)
#endif

__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// FAST: the configuration of most layers (SiLU, bf16 NHWC through 256-bit stores, no residual / split / post-scale) with
// those choices compiled in: the generic epilogue spends ~30 % of its samples on constant loads, compares and branches
// that re-derive them for every 16-channel chunk (ncu source page of dark2.0)
template <bool FAST>
__global__ void __launch_bounds__(FAST ? kThreadsFast : kThreads, 1) conv_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                               const __grid_constant__ CUtensorMap tmap_y,
                                                               const TmaConvArgs p) {
  extern __shared__ uint8_t smem_raw[];
  if (threadIdx.x == 0) TL(7, 0);   // kernel entry (debug timeline)
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_a + (uint32_t)S * p.a_stage_bytes;
  const uint32_t bars = smem_b + (uint32_t)S * p.b_stage_bytes;
  const uint32_t bar_full = bars;
  const uint32_t bar_empty = bars + 8u * kMaxStages;
  const uint32_t bar_tfull = bars + 16u * kMaxStages;
  const uint32_t bar_tempty = bar_tfull + 32u;
  const uint32_t tmem_slot = bar_tempty + 32u;
  const uint32_t sb_base = (tmem_slot + 4u + 15u) & ~15u;  // up to 4 groups x (256 scale + 256 bias) floats
  const uint32_t stage_out = (sb_base + 8192u + 1023u) & ~1023u;  // [group][2] output staging tiles (TMA store)
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // persistent schedule: unit u of `units` takes iterations u, u + units, ...  Without mc a unit is a CTA and an
  // iteration a tile; with mc a unit is a CTA PAIR and iteration i covers the M tiles 2 * (i / n_tiles) + {0, 1} of
  // n-tile i % n_tiles -- both CTAs walk the same (weight group, n-tile, k-block) sequence in lockstep.
  uint32_t crank = 0;
  if (p.mc) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int unit0 = p.mc ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int units = p.mc ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto tile_of = [&](int it) -> int {
    if (!p.mc) return it;
    const int j = (int)p.div_ntiles.div((uint32_t)it);
    return (2 * j + (int)crank) * p.n_tiles + (it - j * p.n_tiles);
  };

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        ptx::mbar_init(bar_full + 8u * s, 1);   // one arrive.expect_tx; A and W complete through tx bytes
        ptx::mbar_init(bar_empty + 8u * s, p.mc ? 2 : 1);  // one tcgen05.commit (mc: of each CTA of the pair)
      }
      for (int a = 0; a < (int)p.acc_stages; ++a) {
        ptx::mbar_init(bar_tfull + 8u * a, 1);
        ptx::mbar_init(bar_tempty + 8u * a, p.epi_split ? 4 * p.ngroups : 4);  // the warps of the epilogue group(s) reading the stage
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp == kTmaWarp && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
  }
  ptx::pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  if (p.mc) {   // the peer's barriers must exist before anything is multicast into them
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  ptx::pdl_wait();   // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  if (threadIdx.x == 0) TL(7, 1);   // dependency wait passed

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer (one elected thread)
    // Everything in the per-k-block path is incremental (no divisions): the single issuing thread runs a
    // serial instruction stream, and in round 1 that stream -- not memory -- was the bottleneck.
    if (ptx::elect_one()) {
      uint32_t s = 0, ph = 0;
      [[maybe_unused]] uint32_t tl_it = 0;
      uint32_t a_dst = smem_a, b_dst = smem_b, full = bar_full, empty = bar_empty;
      const uint32_t tx_bytes = p.a_tx_bytes + p.b_tx_bytes;
      const int64_t wstep = (int64_t)p.BN * p.bk;
      const uint32_t w_half = p.b_tx_bytes >> 1;   // mc: this CTA's half of every W tile
      for (int it = unit0; it < p.iters; it += units) {
        const int tile = tile_of(it);
        const uint32_t rest = p.div_ntiles.div((uint32_t)tile);
        const int nt = tile - (int)rest * p.n_tiles;
        const int n = (int)p.div_tiles_img.div(rest);
        const int timg = (int)rest - n * p.tiles_img;
        const int ty = (int)p.div_tiles_x.div((uint32_t)timg);
        const int tx = timg - ty * p.tiles_x;
        const int g = n / p.group_imgs;
        const int x0 = tx * p.tw * p.stride - p.pad;
        const int y0 = ty * p.th * p.stride - p.pad;
        const __nv_bfloat16* wp = p.w + (int64_t)g * p.w_gstride + (int64_t)nt * p.k_blocks * wstep;
        // pair mode: per kernel row the k-blocks are pair (ox - 1) and pair ox (x in pair units); else one per (tap, c)
        const int nx = p.pair ? 2 : p.ksize;
        const int nc = p.pair ? p.bk : p.cblocks * p.bk;
        for (int dy = 0; dy < p.ksize; ++dy) {
          for (int dx = 0; dx < nx; ++dx) {
            for (int c = 0; c < nc; c += p.bk) {
              ptx::mbar_wait(empty, ph ^ 1u);
#ifdef DCFA_EXP_HALFW
              ptx::mbar_arrive_expect_tx(full, p.mc ? p.a_tx_bytes + w_half : tx_bytes);
#else
              ptx::mbar_arrive_expect_tx(full, tx_bytes);
#endif
              tma_load_4d(a_dst, &tmap, c, p.pair ? tx * p.tw - 1 + dx : x0 + dx, y0 + dy, n, full);
#ifdef DCFA_EXP_HALFW   // timing experiment (WRONG results): every CTA of a pair receives only ITS half of W -- what a 2-CTA MMA would pull
              if (p.mc) {
                ptx::bulk_g2s(b_dst + crank * w_half, reinterpret_cast<const char*>(wp) + crank * w_half, w_half, full);
              } else
#endif
              if (p.mc) {
                // half of the W tile, into the same ring slot of BOTH CTAs; each copy signals the full barrier of the
                // CTA it lands in (a complete_tx that overtakes that CTA's expect_tx is fine: its arrival is still pending)
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                    ::"r"(b_dst + crank * w_half), "l"(reinterpret_cast<const char*>(wp) + crank * w_half), "r"(w_half), "r"(full),
                    "h"((uint16_t)3)
                    : "memory");
              } else {
                ptx::bulk_g2s(b_dst, wp, p.b_tx_bytes, full);
              }
              TL(0, tl_it); ++tl_it;
              wp += wstep;
              a_dst += p.a_stage_bytes; b_dst += p.b_stage_bytes; full += 8u; empty += 8u;
              if (++s == (uint32_t)S) {
                s = 0; ph ^= 1u;
                a_dst = smem_a; b_dst = smem_b; full = bar_full; empty = bar_empty;
              }
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (one elected thread)
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16_f32(BM, p.BN);
      const int kk = p.bk >> 4;  // MMAs (K = 16) per k-block
      // descriptors: only the 14-bit start-address field changes per stage / per K step
      const uint64_t desc_hi = make_kmajor_desc(0u, p.sbo, p.layout_type);
      const uint32_t a_step = p.a_stage_bytes >> 4, b_step = p.b_stage_bytes >> 4;
      const uint32_t a_lo0 = (smem_a & 0x3FFFFu) >> 4, b_lo0 = (smem_b & 0x3FFFFu) >> 4;
      uint32_t s = 0, ph = 0, a_lo = a_lo0, b_lo = b_lo0, full = bar_full, empty = bar_empty;
      uint32_t as = 0, aph = 0;
      [[maybe_unused]] uint32_t tl_it = 0, tl_tile = 0;
      for (int it = unit0; it < p.iters; it += units) {
        ptx::mbar_wait(bar_tempty + 8u * as, aph ^ 1u);
        TL(3, tl_tile); ++tl_tile;
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)p.BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          ptx::mbar_wait(full, ph);
          ptx::tc_fence_after();
          TL(1, tl_it);
          const uint64_t adesc = desc_hi | (uint64_t)a_lo, bdesc = desc_hi | (uint64_t)b_lo;
          ptx::umma_bf16(d_tmem, adesc, bdesc, idesc, kb > 0 ? 1u : 0u);
          for (int k = 1; k < kk; ++k)  // +32 bytes along K per step: +2 in the (addr >> 4) field
            ptx::umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
          if (p.mc) {   // the ring slot is free once BOTH CTAs have consumed it: arrive on the pair's two empty barriers
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(empty), "h"((uint16_t)3) : "memory");
          } else {
            ptx::umma_commit(empty);
          }
          TL(2, tl_it); ++tl_it;
          a_lo += a_step; b_lo += b_step; full += 8u; empty += 8u;
          if (++s == (uint32_t)S) {
            s = 0; ph ^= 1u;
            a_lo = a_lo0; b_lo = b_lo0; full = bar_full; empty = bar_empty;
          }
        }
        ptx::umma_commit(bar_tfull + 8u * as);
        if (++as == p.acc_stages) { as = 0u; aph ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: group 0 = warps 0-3, group 1 = warps 6-9
    const int group = warp < 4 ? 0 : 1 + ((warp - 6) >> 2);
    const uint32_t gmask = (uint32_t)p.ngroups - 1u;   // 2 or 4 groups
    const int q4 = warp & 3;           // TMEM lane quarter this warp may access
    const int r = q4 * 32 + lane;      // accumulator row
    const int gtid = q4 * 32 + lane;   // thread index inside the group
    const int iy = (int)p.div_tw.div((uint32_t)r);
    const int ix = r - iy * p.tw;
    const int HoWo = p.Ho * p.Wo;
    // per-group shared copy of the current (group, n-tile) scale/bias: with the whole 227 KB carved out as
    // shared memory there is no L1 left, so __ldg would pay an L2 round trip per 16-channel chunk
    float* sb = reinterpret_cast<float*>(smem_raw + (sb_base - ptx::smem_u32(smem_raw))) + group * 512;
    int sb_key = -1;
    const uint32_t acc_shift = p.acc_stages == 4u ? 2u : 1u;
    uint32_t slab = 0;   // store slabs issued by this group (selects the staging buffer)
    [[maybe_unused]] uint32_t tl_fine = 0;
    uint32_t tcount = 0;
    for (int it = unit0; it < p.iters; it += units, ++tcount) {
      if (!p.epi_split && (int)(tcount & gmask) != group) continue;
      const int tile = tile_of(it);
      const uint32_t as = tcount & (p.acc_stages - 1u);   // accumulator stage; its parity is the group's
      const uint32_t aph = (tcount >> acc_shift) & 1u;     // the stage's use count, mod 2
      const uint32_t rest = p.div_ntiles.div((uint32_t)tile);
      const int nt = tile - (int)rest * p.n_tiles;
      const int n = (int)p.div_tiles_img.div(rest);
      const int timg = (int)rest - n * p.tiles_img;
      const int ty = (int)p.div_tiles_x.div((uint32_t)timg);
      const int tx = timg - ty * p.tiles_x;
      const int g = n / p.group_imgs;
      const int oy = ty * p.th + iy, ox = tx * p.tw + ix;
      const bool rvalid = iy < p.th && oy < p.Ho && ox < p.Wo;
      const int pix = oy * p.Wo + ox;
      if (q4 == 0 && lane == 0 && group < 2) TL(4 + group, 4 * (tcount >> 1));
      if (g * p.n_tiles + nt != sb_key) {   // uniform across the group's 128 threads
        sb_key = g * p.n_tiles + nt;
        ptx::named_bar_sync(1 + group, 128);  // previous tile's readers are done
        const float* sc = p.scale + (int64_t)g * p.sb_gstride + nt * p.BN;
        const float* bi = p.bias + (int64_t)g * p.sb_gstride + nt * p.BN;
        // FAST (SiLU): the epilogue needs h = (acc * s + b) / 2 -- halve the vectors here (exact: a power of two)
        const float pre = FAST ? 0.5f : 1.0f;
        for (int c = gtid; c < p.BN; c += 128) { sb[c] = pre * __ldg(sc + c); sb[256 + c] = pre * __ldg(bi + c); }
        ptx::named_bar_sync(1 + group, 128);
      }
      __nv_bfloat16* yb = nullptr;
      __nv_bfloat16* yb2 = nullptr;   // split output: destination of channel c is yb2 + c for c >= split
      float* yf = nullptr;
      const __nv_bfloat16* rb = nullptr;
      if (rvalid) {
        if (p.out_mode == DCFA_OUT_BF16_NHWC) {
          yb = reinterpret_cast<__nv_bfloat16*>(p.y.p) + p.y.img_off(n) + (int64_t)pix * p.y.ld + nt * p.BN;
          if (p.split > 0) yb2 = p.y2.p + p.y2.img_off(n) + (int64_t)pix * p.y2.ld + nt * p.BN - p.split;
          if (p.res.p) rb = p.res.p + p.res.img_off(n) + (int64_t)pix * p.res.ld + nt * p.BN;
        } else {
          yf = reinterpret_cast<float*>(p.y.p) + (int64_t)n * p.y.img_stride + (int64_t)(p.out_coff + nt * p.BN) * HoWo + pix;
        }
      }
      const int cvalid = min(p.BN, p.Cout - nt * p.BN);  // valid channels of this n-tile
      ptx::mbar_wait(bar_tfull + 8u * as, aph);
      ptx::tc_fence_after();
      if (q4 == 0 && lane == 0 && group < 2) TL(4 + group, 4 * (tcount >> 1) + 1);
      const uint32_t taddr0 = tmem_base + as * (uint32_t)p.BN + ((uint32_t)(q4 * 32) << 16);
      const int nchunks = p.BN >> 4;
      // two register buffers with STATIC indexing (a dynamically indexed array would live in local memory,
      // and with the whole L1 carved out as shared memory every local access is an L2 round trip)
      uint32_t accA[16], accB[16];
      // A chunk = 16 channels of the thread's row.  compute(): accumulators -> BN -> activation -> post-scale, pure
      // register arithmetic, so the two chunks in flight interleave (one chunk alone is a ~700-cycle dependent chain:
      // LDS -> FFMA2 -> MUFU -> FFMA2; measured with tools/timeline.py); emit(): residual, pack, stores.
      auto compute = [&](const uint32_t (&a)[16], const int j, float (&v)[16]) {
        const int c0 = j * 16;
        // packed fp32x2 arithmetic (fma.rn.f32x2 / mul.rn.f32x2 are single instructions on sm_100): BN and the FMA
        // halves of SiLU take half the issue slots; the results are the same IEEE operations as the scalar forms
        F2 v2[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 s4 = *reinterpret_cast<const float4*>(sb + c0 + 4 * q);
          const float4 b4 = *reinterpret_cast<const float4*>(sb + 256 + c0 + 4 * q);
          v2[2 * q] = f2_make(b4.x, b4.y);
          v2[2 * q + 1] = f2_make(b4.z, b4.w);
          f2_fma(v2[2 * q], f2_make(__uint_as_float(a[4 * q + 0]), __uint_as_float(a[4 * q + 1])), f2_make(s4.x, s4.y));
          f2_fma(v2[2 * q + 1], f2_make(__uint_as_float(a[4 * q + 2]), __uint_as_float(a[4 * q + 3])), f2_make(s4.z, s4.w));
        }
        if (FAST || p.act == DCFA_ACT_SILU) {   // x * sigmoid(x) = h + h * tanh(h), h = x / 2
          const F2 half2 = f2_make(0.5f, 0.5f), zero2 = f2_make(0.0f, 0.0f);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            F2 h = zero2;
            if (FAST) h = v2[e];   // scale and bias were halved when they were staged
            else f2_fma(h, v2[e], half2);
            float h0, h1, t0, t1;
            f2_get(h, h0, h1);
#ifdef DCFA_EXP_NOMUFU
            t0 = h0 * 0.25f; t1 = h1 * 0.25f;
#else
            asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
#endif
            f2_fma(h, h, f2_make(t0, t1));   // h * t + h
            f2_get(h, v[2 * e], v[2 * e + 1]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) f2_get(v2[e], v[2 * e], v[2 * e + 1]);
          if (p.act == DCFA_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.0f);
          }
        }
        if (!FAST && p.post_scale != 1.0f) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] *= p.post_scale;
        }
      };
      auto emit = [&](float (&v)[16], const int j) {
        const int c0 = j * 16;
        if (FAST || p.st256) {
#ifdef DCFA_EXP_NOSTORE
          if (rvalid && c0 < cvalid && v[3] == 12345.678f) {
#else
          if (rvalid && c0 < cvalid) {
#endif
            uint4 lo, hi;
            if (!FAST && rb) {
              uint32_t q[8];
              asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                           : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
                           : "l"(rb + c0));
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float2 f = unpack_bf16x2(q[e]);
                v[2 * e] += f.x;
                v[2 * e + 1] += f.y;
              }
            }
            lo = pack8(v);
            hi = pack8(v + 8);
            __nv_bfloat16* dst = (!FAST && p.split > 0 && nt * p.BN + c0 >= p.split) ? yb2 + c0 : yb + c0;
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(lo.x), "r"(lo.y),
                         "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                         : "memory");
          }
        } else if (p.tma_store) {
          // ---- stage 16 channels of this row into the swizzled slab; a full slab leaves with one TMA store
          const int cs = c0 & (p.cbox - 1);      // channel offset inside the slab (cbox is 16, 32 or 64)
          if (cs == 0) {
            // the slab buffer about to be overwritten was read by the TMA store issued two slabs ago
            if (gtid == 0) bulk_wait_read<1>();
            ptx::named_bar_sync(1 + group, 128);
          }
          const uint32_t pitch = (uint32_t)p.cbox * 2u;
          const uint32_t rowb = stage_out + (uint32_t)(group * 2 + (int)(slab & 1u)) * p.out_stage_bytes + (uint32_t)r * pitch;
          // Swizzle<B,4,3>: 16-byte chunk index XOR (row bits above the 128-byte line)
          const uint32_t xr = p.cbox == 64 ? (uint32_t)(r & 7) : (p.cbox == 32 ? (uint32_t)((r >> 1) & 3) : (uint32_t)((r >> 2) & 1));
          const uint32_t ch = (uint32_t)(cs >> 3);
          const uint4 lo = pack8(v), hi = pack8(v + 8);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + (((ch) ^ xr) << 4)), "r"(lo.x), "r"(lo.y),
                       "r"(lo.z), "r"(lo.w)
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + (((ch + 1u) ^ xr) << 4)), "r"(hi.x),
                       "r"(hi.y), "r"(hi.z), "r"(hi.w)
                       : "memory");
          if (cs + 16 == p.cbox) {   // slab complete
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(1 + group, 128);
            if (gtid == 0 && nt * p.BN + c0 + 16 - p.cbox < p.Cout) {
              tma_store_4d(&tmap_y, stage_out + (uint32_t)(group * 2 + (int)(slab & 1u)) * p.out_stage_bytes,
                           nt * p.BN + c0 + 16 - p.cbox, tx * p.tw, ty * p.th, n);
            }
            if (gtid == 0) bulk_commit();
            ++slab;
          }
        } else if (rvalid && c0 < cvalid) {
          if (p.out_mode == DCFA_OUT_BF16_NHWC) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (c0 + 8 * h + 8 <= cvalid) {
                if (rb) {
                  float rr[8];
                  unpack8(ldg128(rb + c0 + 8 * h), rr);
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[8 * h + e] += rr[e];
                }
                stg128(yb + c0 + 8 * h, pack8(v + 8 * h));
              }
            }
          } else {
            // fp32 NCHW: lanes are consecutive pixels, so every store instruction writes whole 32-byte sectors.  One
            // running pointer (a 64-bit add per channel) instead of a 64-bit multiply per element; full chunks skip the
            // per-channel predicate (the head map's 64 + nc channels end inside the last chunk only)
            {
              float* q = yf + (int64_t)c0 * HoWo;
              if (c0 + 16 <= cvalid) {
#pragma unroll
                for (int e = 0; e < 16; ++e) { *q = v[e]; q += HoWo; }
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) { if (c0 + e < cvalid) *q = v[e]; q += HoWo; }
              }
            }
            if (p.dfl_dbox) {
              // DFL (nets/yolo_mul.py:312-322) straight from the fp32 accumulators: a 16-channel chunk of the box part is
              // one side's 16 bins -- softmax expectation in registers; the class logits are gathered into (B, nc, A)
              // (:459-460).  The thread's row is one anchor of this level.
              const int cg0 = nt * p.BN + c0;
              const int64_t a = (int64_t)p.dfl_aoff + pix;
              if (cg0 < 64) {
                // max by a tree (four dependent steps instead of fifteen), exp(v - mx) as ex2(v * log2e - mx * log2e): one
                // FFMA + one MUFU per bin, two independent accumulation chains each for the sum and the expectation
                float m01 = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), m23 = fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7]));
                float m45 = fmaxf(fmaxf(v[8], v[9]), fmaxf(v[10], v[11])), m67 = fmaxf(fmaxf(v[12], v[13]), fmaxf(v[14], v[15]));
                const float mx = fmaxf(fmaxf(m01, m23), fmaxf(m45, m67));
                const float kLog2e = 1.4426950408889634f;
                const float off = -mx * kLog2e;
                float den0 = 0.0f, den1 = 0.0f, num0 = 0.0f, num1 = 0.0f;
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                  float e0, e1;
                  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(v[e], kLog2e, off)));
                  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(v[e + 1], kLog2e, off)));
                  den0 += e0; den1 += e1;
                  num0 = fmaf((float)e, e0, num0);
                  num1 = fmaf((float)(e + 1), e1, num1);
                }
                p.dfl_dbox[((int64_t)n * 4 + (cg0 >> 4)) * p.dfl_A + a] = __fdividef(num0 + num1, den0 + den1);
              } else {
                float* q = p.dfl_cls + ((int64_t)n * p.dfl_nc + (cg0 - 64)) * p.dfl_A + a;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  if (cg0 - 64 + e < p.dfl_nc) *q = v[e];
                  q += p.dfl_A;
                }
              }
            }
          }
        }
      };
      // this group's chunks: all of them, or (split) every other one starting at `group`
      const int jstep = p.epi_split ? p.ngroups : 1, j0 = p.epi_split ? group : 0;
      ptx::tmem_ld_x16(taddr0 + (uint32_t)(j0 * 16), accA);
      if (j0 + jstep < nchunks) ptx::tmem_ld_x16(taddr0 + (uint32_t)((j0 + jstep) * 16), accB);
      for (int j = j0; j < nchunks; j += 2 * jstep) {
        const bool two = j + jstep < nchunks;
        float vA[16], vB[16];
#ifdef DCFA_TIMELINE
        if (group == 0 && q4 == 0 && lane == 0) { TL(6, tl_fine); ++tl_fine; }
#endif
        ptx::tmem_ld_wait();
#ifdef DCFA_TIMELINE
        if (group == 0 && q4 == 0 && lane == 0) { TL(6, tl_fine); ++tl_fine; }
#endif
        compute(accA, j, vA);
        if (two) compute(accB, j + jstep, vB);
#ifdef DCFA_TIMELINE
        if (group == 0 && q4 == 0 && lane == 0) { asm volatile("" :: "f"(vA[0]), "f"(vA[15]), "f"(vB[0]), "f"(vB[15])); TL(6, tl_fine); ++tl_fine; }
#endif
        // both register buffers are consumed: the next pair of chunks loads while this pair is stored
        if (j + 2 * jstep < nchunks) ptx::tmem_ld_x16(taddr0 + (uint32_t)((j + 2 * jstep) * 16), accA);
        if (j + 3 * jstep < nchunks) ptx::tmem_ld_x16(taddr0 + (uint32_t)((j + 3 * jstep) * 16), accB);
        emit(vA, j);
        __syncwarp();
        if (two) {
          emit(vB, j + jstep);
          __syncwarp();
        }
#ifdef DCFA_TIMELINE
        if (group == 0 && q4 == 0 && lane == 0) { TL(6, tl_fine); ++tl_fine; }
#endif
      }
      if (q4 == 0 && lane == 0 && group < 2) TL(4 + group, 4 * (tcount >> 1) + 2);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + 8u * as);
      if (q4 == 0 && lane == 0 && group < 2) TL(4 + group, 4 * (tcount >> 1) + 3);
    }
    if (p.tma_store && gtid == 0) bulk_wait<0>();   // staging smem must outlive the last TMA store
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (p.mc) {   // the peer may still arrive on this CTA's empty barriers: leave together
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
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

// spatial tile (tw, th) with tw*th <= 128, tw <= 128/stride... maximising useful rows per 128-row MMA tile
void pick_tile(int Ho, int Wo, int* tw_out, int* th_out) {
  double best = -1.0;
  int btw = 1, bth = 1;
  for (int tw = 1; tw <= 128 && tw <= Wo; ++tw) {
    int th = 128 / tw;
    if (th > Ho) th = Ho;
    if (th < 1) continue;
    const int64_t tiles = (int64_t)((Ho + th - 1) / th) * ((Wo + tw - 1) / tw);
    const double util = (double)Ho * Wo / (double)(tiles * 128);
    // prefer wider tiles on ties (longer contiguous runs per box row)
    if (util > best + 1e-9 || (util > best - 1e-9 && tw > btw)) {
      best = util;
      btw = tw;
      bth = th;
    }
  }
  *tw_out = btw;
  *th_out = bth;
}

}  // namespace

// Returns DCFA_OK after launching, or a negative code.  Called by launch_conv when the op was packed for TMA.
int launch_conv_tma(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  {   // 3x3 stride-1 SiLU layers: nine taps from one shared-memory strip (conv_strip.cu) when the shape qualifies
    bool taken = false;
    const int rc = launch_conv_strip(op, bufs, st, &taken);
    if (taken) return rc;
  }
  TmaConvArgs a;
  View<const __nv_bfloat16> x = resolve<const __nv_bfloat16>(op.x, bufs);
  a.res = resolve<const __nv_bfloat16>(op.x2, bufs);
  a.y = resolve<void>(op.y, bufs);
  a.y2 = resolve<__nv_bfloat16>(op.a0, bufs);
  a.split = op.parts;
  a.w = resolve_ptr<const __nv_bfloat16>(op.w, bufs);
  a.scale = resolve_ptr<const float>(op.scale, bufs);
  a.bias = resolve_ptr<const float>(op.bias, bufs);
  a.w_gstride = op.w_gstride;
  a.sb_gstride = op.sb_gstride;
  a.n_img = op.n_img;
  a.group_imgs = op.group_imgs > 0 ? op.group_imgs : op.n_img;
  int Hi = op.Hi, Wi = op.Wi;
  const int Cin = op.Cin;
  a.Ho = op.Ho; a.Wo = op.Wo; a.Cout = op.Cout;
  a.ksize = op.ksize; a.stride = op.stride; a.pad = op.ksize / 2;
  a.BN = op.BN; a.n_tiles = op.n_tiles; a.k_blocks = op.k_blocks;
  a.bk = op.flags & 0xff;
  a.pair = (op.flags & DCFA_CONV_FLAG_PAIR) ? 1 : 0;
  a.act = op.act; a.out_mode = op.out_mode; a.out_ctot = op.out_ctot; a.out_coff = op.out_coff;
  a.post_scale = op.f0;
  a.dfl_dbox = nullptr; a.dfl_cls = nullptr; a.dfl_A = a.dfl_aoff = a.dfl_nc = 0;
  if (op.flags & DCFA_CONV_FLAG_DFL) {
    a.dfl_dbox = resolve_ptr<float>(op.a1, bufs);
    a.dfl_cls = resolve_ptr<float>(op.a2, bufs);
    a.dfl_A = op.A; a.dfl_aoff = op.hidden; a.dfl_nc = op.nc;
    DCFA_REQUIRE(a.dfl_dbox && a.dfl_cls && op.out_mode == DCFA_OUT_F32_NCHW && op.out_coff == 0 && op.Cout == 64 + op.nc &&
                     op.nc >= 1 && op.act == DCFA_ACT_NONE && op.hidden >= 0 && op.hidden + op.Ho * op.Wo <= op.A,
                 "conv(tma): DCFA_CONV_FLAG_DFL needs the fp32 NCHW head map [64 box | nc cls] and its dbox / cls outputs");
  }

  DCFA_REQUIRE(x.p && a.y.p && a.w && a.scale && a.bias, "conv(tma): missing tensor");
  DCFA_REQUIRE(a.bk == 64 || a.bk == 32 || a.bk == 16, "conv(tma): bk %d unsupported", a.bk);
  if (a.pair)
    DCFA_REQUIRE(a.bk == 64 && Cin == 32 && op.ksize == 3 && op.stride == 2 && op.Wi % 2 == 0 && op.k_blocks == 6 &&
                     op.K_real == 288,
                 "conv(tma): pixel-pair packing needs a 3x3 stride-2 conv with Cin = 32 on an even-width input");
  else
    DCFA_REQUIRE(Cin % a.bk == 0, "conv(tma): Cin %d not a multiple of bk %d", Cin, a.bk);
  DCFA_REQUIRE(a.n_img > 0 && a.n_img % a.group_imgs == 0, "conv(tma): bad grouping");
  DCFA_REQUIRE(a.ksize == 1 || a.ksize == 3, "conv(tma): ksize %d unsupported", a.ksize);
  DCFA_REQUIRE(a.stride == 1 || a.stride == 2, "conv(tma): stride %d unsupported", a.stride);
  DCFA_REQUIRE(a.Ho == (Hi + 2 * a.pad - a.ksize) / a.stride + 1 && a.Wo == (Wi + 2 * a.pad - a.ksize) / a.stride + 1,
               "conv(tma): output size inconsistent");
  DCFA_REQUIRE(a.BN >= 16 && a.BN <= 256 && a.BN % 16 == 0, "conv(tma): BN %d invalid", a.BN);
  DCFA_REQUIRE(a.n_tiles >= 1 && a.n_tiles * a.BN >= a.Cout, "conv(tma): n_tiles*BN < Cout");
  a.cblocks = a.pair ? 1 : Cin / a.bk;
  DCFA_REQUIRE(a.pair || (a.k_blocks == a.ksize * a.ksize * a.cblocks && op.K_real == a.ksize * a.ksize * Cin),
               "conv(tma): k_blocks %d inconsistent with packing", a.k_blocks);
  DCFA_REQUIRE(x.gi <= 0 || x.gstride == (int64_t)x.gi * x.img_stride, "conv(tma): grouped input views unsupported");
  DCFA_REQUIRE(((uintptr_t)x.p % 16) == 0 && x.ld % 8 == 0 && x.img_stride % 8 == 0, "conv(tma): input must be 16-byte aligned");
  DCFA_REQUIRE(((uintptr_t)a.w % 16) == 0 && a.w_gstride % 8 == 0, "conv(tma): weights must be 16-byte aligned");
  DCFA_REQUIRE(((uintptr_t)a.scale % 16) == 0 && ((uintptr_t)a.bias % 16) == 0 && a.sb_gstride % 4 == 0,
               "conv(tma): scale/bias must be 16-byte aligned");
  if (a.out_mode == DCFA_OUT_BF16_NHWC) {
    DCFA_REQUIRE(a.Cout % 8 == 0, "conv(tma): bf16 output needs Cout %% 8 == 0");
    DCFA_REQUIRE(((uintptr_t)a.y.p % 16) == 0 && a.y.ld % 8 == 0 && a.y.img_stride % 8 == 0 && a.y.gstride % 8 == 0,
                 "conv(tma): output view must be 16-byte aligned");
    if (a.res.p)
      DCFA_REQUIRE(((uintptr_t)a.res.p % 16) == 0 && a.res.ld % 8 == 0 && a.res.img_stride % 8 == 0 && a.res.gstride % 8 == 0,
                   "conv(tma): residual view must be 16-byte aligned");
  } else {
    DCFA_REQUIRE(a.out_mode == DCFA_OUT_F32_NCHW, "conv(tma): bad out_mode");
    DCFA_REQUIRE(!a.res.p, "conv(tma): residual unsupported with fp32 NCHW output");
    DCFA_REQUIRE(a.out_coff >= 0 && a.out_coff + a.Cout <= a.out_ctot, "conv(tma): NCHW channel slot out of range");
    DCFA_REQUIRE(a.y.img_stride == (int64_t)a.out_ctot * a.Ho * a.Wo, "conv(tma): NCHW img_stride mismatch");
  }

  // ---- 1x1 stride-1 convolutions over dense tensors are plain GEMMs over pixels: describe each weight group as ONE
  // image of height 1 and width group_imgs * H * W, so that an M tile is 128 consecutive pixels whatever the map size.
  // At 40 x 40 a spatial tile holds 3 x 40 = 120 pixels: 448 tiles for 32 images, i.e. FOUR waves on 148 SMs for
  // 3.03 tiles per CTA; flattened it is 400 tiles and three waves (dn1.cv1: 23 -> 18 us).
  {
    const char* e = getenv("DCFA_CONV_FLAT");   // debug / tests: DCFA_CONV_FLAT=0 keeps the spatial tiling
    const int hw = a.Ho * a.Wo;
    auto uniform = [&](int64_t ld, int64_t img_stride, int gi) {   // pixels of a whole weight group are equidistant
      return img_stride == (int64_t)hw * ld && (gi <= 0 || gi == a.group_imgs);
    };
    bool flat = !(e && atoi(e) == 0) && a.ksize == 1 && a.stride == 1 && !a.pair && a.out_mode == DCFA_OUT_BF16_NHWC &&
                !a.dfl_dbox && (int64_t)a.group_imgs * hw < (1ll << 30) && uniform(x.ld, x.img_stride, x.gi) &&
                uniform(a.y.ld, a.y.img_stride, a.y.gi) && (!a.res.p || uniform(a.res.ld, a.res.img_stride, a.res.gi)) &&
                (a.split == 0 || uniform(a.y2.ld, a.y2.img_stride, a.y2.gi));
    if (flat) {
      int tw0, th0;
      pick_tile(a.Ho, a.Wo, &tw0, &th0);
      const int64_t spatial = (int64_t)a.n_img * ((a.Ho + th0 - 1) / th0) * ((a.Wo + tw0 - 1) / tw0);
      const int64_t P = (int64_t)a.group_imgs * hw;
      const int groups = a.n_img / a.group_imgs;
      if (groups * ((P + 127) / 128) < spatial) {
        auto regroup = [&](auto& v) {   // one "image" per weight group
          v.img_stride = v.gi > 0 ? v.gstride : (int64_t)a.group_imgs * v.img_stride;
          v.gi = 0; v.gstride = 0;
        };
        regroup(x); regroup(a.y);
        if (a.res.p) regroup(a.res);
        if (a.split > 0) regroup(a.y2);
        a.n_img = groups; a.group_imgs = 1;
        Hi = 1; Wi = (int)P; a.Ho = 1; a.Wo = (int)P;
      }
    }
  }

  pick_tile(a.Ho, a.Wo, &a.tw, &a.th);
  a.tiles_x = (a.Wo + a.tw - 1) / a.tw;
  const int tiles_y = (a.Ho + a.th - 1) / a.th;
  a.tiles_img = a.tiles_x * tiles_y;
  const int64_t total = (int64_t)a.n_img * a.tiles_img * a.n_tiles;
  DCFA_REQUIRE(total < (1ll << 31), "conv(tma): too many tiles");
  a.total_tiles = (int)total;
  a.div_tw = make_fastdiv((uint32_t)a.tw);
  a.div_tiles_img = make_fastdiv((uint32_t)a.tiles_img);
  a.div_tiles_x = make_fastdiv((uint32_t)a.tiles_x);
  a.div_ntiles = make_fastdiv((uint32_t)a.n_tiles);

  const uint32_t pitch = (uint32_t)a.bk * 2u;  // bytes per K row of a stage
  a.layout_type = a.bk == 64 ? 2u : (a.bk == 32 ? 4u : 6u);
  a.sbo = 8u * pitch;
  a.a_stage_bytes = 128u * pitch;                                  // multiple of 1024 for every bk
  a.b_tx_bytes = (uint32_t)a.BN * pitch;
  a.b_stage_bytes = (a.b_tx_bytes + 1023u) & ~1023u;
  a.a_tx_bytes = (uint32_t)(a.tw * a.th) * pitch;
  const int stage_bytes = (int)(a.a_stage_bytes + a.b_stage_bytes);
  const int max_smem = 227 * 1024;
  // epilogue store path for bf16 NHWC outputs
  a.st256 = 0;
  {
    const char* e = getenv("DCFA_ST256");   // debug: DCFA_ST256=0 forces the TMA-store / 128-bit paths
    const bool want = !(e && atoi(e) == 0);
    const bool ok = a.out_mode == DCFA_OUT_BF16_NHWC && a.Cout % 16 == 0 && ((uintptr_t)a.y.p % 32) == 0 && a.y.ld % 16 == 0 &&
                    a.y.img_stride % 16 == 0 && a.y.gstride % 16 == 0 &&
                    (!a.res.p || (((uintptr_t)a.res.p % 32) == 0 && a.res.ld % 16 == 0 && a.res.img_stride % 16 == 0 &&
                                  a.res.gstride % 16 == 0));
    if (want && ok) a.st256 = 1;
  }
  if (a.split > 0) {
    DCFA_REQUIRE(a.y2.p && a.split % 16 == 0 && a.split < a.Cout && !a.res.p, "conv(tma): bad split output");
    DCFA_REQUIRE(((uintptr_t)a.y2.p % 32) == 0 && a.y2.ld % 16 == 0 && a.y2.img_stride % 16 == 0 && a.y2.gstride % 16 == 0,
                 "conv(tma): split destination must be 32-byte aligned");
    DCFA_REQUIRE(a.st256, "conv(tma): split output needs the 256-bit store path (aligned bf16 NHWC destinations)");
  }
  a.tma_store = (!a.st256 && a.out_mode == DCFA_OUT_BF16_NHWC && !a.res.p && a.y.gi <= 0) ? 1 : 0;
  a.cbox = a.BN >= 64 ? 64 : a.BN;   // BN is 16, 32, 48 or a multiple of 64 below
  if (a.tma_store && (a.BN % a.cbox != 0 || (a.cbox != 64 && a.cbox != 32 && a.cbox != 16))) a.tma_store = 0;
  a.out_stage_bytes = 128u * (uint32_t)a.cbox * 2u;
  const int fixed = 1024 + 512 + 8192 + 1024 + (a.tma_store ? 4 * (int)a.out_stage_bytes : 0);
  int stages = (max_smem - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  { const char* e = getenv("DCFA_STAGES"); if (e && atoi(e) > 1 && atoi(e) < stages) stages = atoi(e); }
  DCFA_REQUIRE(stages >= 2, "conv(tma): not enough shared memory");
  a.stages = stages;
  const int smem = fixed + stages * stage_bytes;
  a.acc_stages = 4 * a.BN <= 512 ? 4u : 2u;
  {
    const char* e = getenv("DCFA_EPI_SPLIT");   // debug: DCFA_EPI_SPLIT=0 keeps the groups on alternating tiles
    a.epi_split = (a.BN >= 128 && !a.tma_store && !(e && atoi(e) == 0)) ? 1 : 0;
  }
  uint32_t cols = 32;
  while (cols < a.acc_stages * (uint32_t)a.BN) cols <<= 1;
  a.tmem_cols = cols;

  // ---- tensor map over the input view: dims (C, W, H, N), innermost first
  EncodeTiledFn enc = encode_tiled_fn();
  DCFA_REQUIRE(enc != nullptr, "conv(tma): cuTensorMapEncodeTiled entry point unavailable");
  alignas(64) CUtensorMap tmap;
  cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)a.n_img};
  cuuint64_t gstr[3] = {(cuuint64_t)x.ld * 2, (cuuint64_t)Wi * x.ld * 2, (cuuint64_t)x.img_stride * 2};
  cuuint32_t box[4] = {(cuuint32_t)a.bk, (cuuint32_t)(a.tw * a.stride), (cuuint32_t)(a.th * a.stride), 1u};
  cuuint32_t estr[4] = {1u, (cuuint32_t)a.stride, (cuuint32_t)a.stride, 1u};
  if (a.pair) {   // the W unit is a PAIR of pixels (64 channels, 128 contiguous bytes): dense along x, stride 2 along y only
    DCFA_REQUIRE(x.ld == Cin, "conv(tma): pixel-pair packing needs a dense input (ld == Cin)");
    gdim[0] = 2 * (cuuint64_t)Cin;
    gdim[1] = (cuuint64_t)Wi / 2;
    gstr[0] = (cuuint64_t)x.ld * 4;
    box[1] = (cuuint32_t)a.tw;
    estr[1] = 1u;
  }
  DCFA_REQUIRE(box[1] <= 256 && box[2] <= 256, "conv(tma): box too large");
  const CUtensorMapSwizzle swz = a.bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (a.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  // L2 fill granularity: a channel sub-view (Cin < ld) only touches Cin*2 contiguous bytes per pixel; promoting
  // beyond that run would fetch the neighbouring channels from DRAM for nothing
  const int64_t run = (Cin == x.ld) ? (int64_t)Cin * 2 * Wi : (int64_t)Cin * 2;
  const CUtensorMapL2promotion promo = run >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : (run >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B);
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x.p), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(DCFA_E_CUDA, "conv(tma): cuTensorMapEncodeTiled failed with %d", (int)cr);

  static DeviceOnce attr_set;
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(conv_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return fail(DCFA_E_CUDA, "conv(tma): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.mark();
  }
  bool fast = a.act == DCFA_ACT_SILU && a.st256 && !a.res.p && a.split == 0 && a.post_scale == 1.0f &&
              a.out_mode == DCFA_OUT_BF16_NHWC && !a.dfl_dbox;
  { const char* e = getenv("DCFA_CONV_FAST"); if (e && atoi(e) == 0) fast = false; }   // debug / tests: the generic instance
  int grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
  // ---- CTA pairs sharing every W tile by multicast.  Both CTAs of a pair must walk the same weight sequence: the M tiles
  // 2j and 2j + 1 have to belong to the same weight group, so the tiles of a group must come in an even number.
  a.mc = 0;
  a.iters = a.total_tiles;
  {
    // OFF by default: parity-green, but measured on B200 (s, B=32) head0.0 0.079 / 0.080 ms, head0.cls1 0.070 / 0.068,
    // dark4.0 0.067 / 0.066, head1.0 0.052 / 0.052 with / without -- halving the W bytes each SM pulls from L2 buys nothing,
    // like the resident-weight experiments before it: the k-block rate is set by the A side (profiles/README.md, round 2).
    const char* e = getenv("DCFA_CONV_MC");   // experiments / tests: n > 0 = CTA pairs for BN >= n
    const int min_bn = e ? atoi(e) : 0;
    const int64_t mtiles_group = (int64_t)a.group_imgs * a.tiles_img;
    if (min_bn > 0 && a.BN >= min_bn && mtiles_group % 2 == 0 && a.total_tiles >= 2 && (a.b_tx_bytes >> 1) % 16 == 0 &&
        sm_count() >= 2) {
      a.mc = 1;
      a.iters = a.total_tiles / 2;
      const int pairs = a.iters < sm_count() / 2 ? a.iters : sm_count() / 2;
      grid = 2 * pairs;
    }
  }
  alignas(64) CUtensorMap tmap_y;
  memset(&tmap_y, 0, sizeof(tmap_y));
  if (a.tma_store) {
    const cuuint64_t ydim[4] = {(cuuint64_t)a.Cout, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)a.n_img};
    const cuuint64_t ystr[3] = {(cuuint64_t)a.y.ld * 2, (cuuint64_t)a.Wo * a.y.ld * 2, (cuuint64_t)a.y.img_stride * 2};
    const cuuint32_t ybox[4] = {(cuuint32_t)a.cbox, (cuuint32_t)a.tw, (cuuint32_t)a.th, 1u};
    const cuuint32_t yes[4] = {1u, 1u, 1u, 1u};
    const CUtensorMapSwizzle yswz = a.cbox == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                 : (a.cbox == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cr = enc(&tmap_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.y.p, ydim, ystr, ybox, yes, CU_TENSOR_MAP_INTERLEAVE_NONE, yswz,
             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(DCFA_E_CUDA, "conv(tma): cuTensorMapEncodeTiled(output) failed with %d", (int)cr);
  }
  a.ngroups = 2;
  if (fast) {   // latency-bound epilogue (ncu: 16 % of its samples issue): four groups instead of two
    const char* e = getenv("DCFA_EPI_GROUPS");   // debug: DCFA_EPI_GROUPS=2 launches the FAST instance with two groups
    a.ngroups = (e && atoi(e) == 2) ? 2 : kFastGroups;
  }
  if (fast) launch_k(conv_tma_kernel<true>, dim3(grid), dim3(64 + 128 * a.ngroups), smem, st, a.mc ? 2 : 0, true, tmap, tmap_y, a);
  else launch_k(conv_tma_kernel<false>, dim3(grid), dim3(kThreads), smem, st, a.mc ? 2 : 0, true, tmap, tmap_y, a);
  DCFA_CHECK_LAUNCH("conv_tma_kernel");
  return DCFA_OK;
}

}  // namespace dcfa

#ifdef DCFA_TIMELINE
extern "C" int dcfa_debug_read_timeline(void* dst, int bytes) {
  return cudaMemcpyFromSymbol(dst, dcfa::g_tl, bytes) == cudaSuccess ? 0 : -2;
}
extern "C" int dcfa_debug_clear_timeline() {
  static long long zeros[8][2048];
  return cudaMemcpyToSymbol(dcfa::g_tl, zeros, sizeof(zeros)) == cudaSuccess ? 0 : -2;
}
#endif
