// DCFA_OP_STEM: Conv_maxpool (nets/yolo_mul.py:104-115) fused into one pass on tcgen05 tensor cores:
//   image -> conv3x3 s1 p1 (3 -> C0) -> BN scale/bias -> ReLU -> maxpool 3x3 s2 p1 -> bf16 NHWC.
// The full-resolution conv map (the largest tensor of the network) never leaves the SM.  Both modalities run in one
// launch (weight group = image / group_imgs).
//
// Round-1 version (ncu: 41 % issue-active, 0.97 TB/s): a generic-proxy im2col (27 LDS + converts + 4 STS per conv
// pixel) and four CTA-wide phases per tile serialised behind __syncthreads.  This version has NO im2col:
//
//   * the input patch of a tile (11 x 28 pixels) is converted ONCE to bf16 [y][x][4 channels] (8 bytes per pixel,
//     channel 3 = 0) -- 1.3 shared-memory operations per input pixel instead of ~35 per conv pixel;
//   * the tcgen05 shared-memory descriptor does the im2col: NO-SWIZZLE K-major layout with LBO = 16 B (next K chunk =
//     next pixel pair) and SBO = 128 B (next 8 rows), i.e. operand row n = the 16 bf16 starting at pixel 2n of the
//     LINEAR patch (pitch 28): four consecutive pixels x 4 channels.  Rows overlap in memory; the hardware only
//     computes addresses (tools/umma_overlap_test.cu proves it on a B200).  The even conv pixel 2n uses weights
//     [w(ky,0) | w(ky,1) | w(ky,2) | 0], the odd conv pixel 2n+1 uses [0 | w(ky,0) | w(ky,1) | w(ky,2)] over the SAME
//     operand rows, so both PARITIES are stacked on the M dimension of ONE K=16 MMA per kernel row: in every TMEM lane
//     quarter, lanes 0..15 hold the even conv columns of 16 channels and lanes 16..31 the odd ones (measured: the
//     stem's time is linear in its MMA count, 0.317 / 0.373 / 0.461 ms for 0 / 3 / 6 MMAs per tile);
//   * the GEMM is turned so that CHANNELS are MMA rows and pixels are TMEM columns: a thread holds three conv rows
//     (14 TMEM columns each) of ONE channel and one column parity; the vertical 3-max is pure register arithmetic
//     with three-input max, the two lanes of a parity pair (16 apart in the same warp) swap half of their column
//     maxima with seven shuffles and each finishes six pooled pixels (a first version that paired two WARPS through
//     shared memory and a named barrier per row was 0.1 ms slower than not stacking at all); BN + ReLU after pooling
//     (exact: the sign of the BN scale is folded into the weights, so the pooled quantity is monotone in the
//     accumulator); the 12 pooled pixels x 16 channels of a warp leave through a 384-byte shared-memory slab and ONE
//     TMA store (no global address arithmetic, edges clipped by the tensor map);
//   * warp-specialised pipeline, one CTA per SM, all of TMEM (4 accumulator stages x 128 columns):
//       warp 20      TMA producer: raw patch ring (fp32 NCHW planes, uint8 NHWC rows or a uint8 depth plane)
//       warps 16-19  converters, ONE TILE PER WARP (four tiles in flight: the dependent LDS -> convert -> STS chain and the
//                    proxy fence of one tile overlap the other three): raw -> bf16 [y][x][4], fence.proxy.async, arrive
//       warp 21      MMA issuer: 3 x (M128 N128 K16) per tile, commits free the converted slot / publish the accumulators
//       warps 0-15   four epilogue groups, one per accumulator stage, taking every fourth tile (ncu: an epilogue warp
//                    spends ~1 500 cycles per tile in dependent tcgen05.ld -> max -> shuffle -> convert -> st.shared ->
//                    proxy fence -> TMA store chains; four tiles in flight hide them).
// Tile: 4 x 12 pooled pixels = 9 x 25 conv pixels (linear index L = cy*28 + cx, parity = cx & 1, TMEM column L >> 1).
// Out-of-image conv positions are excluded from the max (reference: -inf pool padding); zero padding of the conv comes
// from the TMA's out-of-bounds fill.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {
namespace {

constexpr int TPH = 4, TPW = 12;              // pooled tile
constexpr int CH = 2 * TPH + 1;               // 9 conv rows
constexpr int CW = 2 * TPW + 1;               // 25 conv cols
constexpr int PH = CH + 2;                    // 11 patch rows
constexpr int PP = 28;                        // patch pitch in pixels (27 needed; even, so that parity(L) = parity(cx))
constexpr int HP = PP / 2;                    // TMEM columns per conv row and parity
constexpr int NCOL = 128;                     // MMA N = pixel pairs per parity (9 * 14 = 126 used)
constexpr int XOFF = 2;                       // fp32: the TMA box starts 2 floats left of the patch (16-byte aligned start)
constexpr int PWB = 32;                       // fp32 raw row pitch in floats (XOFF + 28, rounded to 16 bytes)
constexpr int RAW_F32_BYTES = 3 * PH * PWB * 4;   // 4224
constexpr int RAWB = 128;                     // uint8 NHWC raw row bytes: 3 * (14 + 28) = 126, rounded to 16
constexpr int RAW_U8_BYTES = RAWB * PH;       // 1408
constexpr int RAW1B = 48;                     // uint8 single plane raw row bytes: 14 + 28 = 42, rounded to 16
constexpr int RAW_C1_BYTES = RAW1B * PH;      // 528
constexpr int RAW_SLOT = 4224;                // bytes per raw ring slot (multiple of 128)
constexpr int NRAW = 8;                       // raw ring depth: the TMA producer runs 8 tiles ahead (34 KB in flight per SM --
                                              // with 4 the input stream sat near 1 TB/s: bytes in flight = bandwidth x latency)
constexpr int CVT_SLOT = 2560;                // bf16 [11][28][4] = 2464 + the rows the last MMA over-reads (columns 126, 127)
constexpr int NCVT = 4;                       // converted-patch ring depth = converter warps
constexpr int A_TILE_BYTES = 128 * 16 * 2;    // one kernel row's weight tile (both parities stacked on the rows)
constexpr int A_GROUP_BYTES = 3 * A_TILE_BYTES;
constexpr int HW6 = TPW / 2;                   // pooled pixels one lane finishes per tile row
constexpr int OUT_SLAB = 2 * TPW * 16 * 2;    // 768: two pooled rows of a warp, [2][12 pixels][16 channels] bf16
constexpr int NACC = 4;                       // TMEM accumulator stages (128 columns each)
constexpr int kEpiWarps = 16, kCvtWarps = 4;
constexpr int kCvtWarp0 = kEpiWarps, kTmaWarp = kEpiWarps + kCvtWarps, kMmaWarp = kTmaWarp + 1;
constexpr int kStemThreads = (kMmaWarp + 1) * 32;   // 704
constexpr uint32_t kTmemCols = 512;
static_assert(NCVT == kCvtWarps && NRAW % kCvtWarps == 0 && (NRAW & (NRAW - 1)) == 0,
              "converter warp w owns converted slot w and the raw slots w, w + 4, ...");
static_assert(XOFF == 2 && PWB == 32, "the converter reads 16 float pairs per raw row, pair 0 = the alignment padding");
static_assert(2 * PP * 8 + (NCOL - 1) * 16 + 32 <= CVT_SLOT, "MMA over-read must stay inside the slot");
static_assert((CH - 1) * HP + TPW + 1 <= NCOL, "tile does not fit the accumulator");

enum { MODE_F32 = 0, MODE_U8 = 1, MODE_U8_C1 = 2 };   // MODE_U8_C1: group 0 uint8 NHWC, group 1 a single uint8 plane

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

struct StemArgs {
  const void* x[2];        // fp32 NCHW / uint8 NHWC / (group 1, MODE_U8_C1) uint8 [N,H,W]
  const __nv_bfloat16* w;  // [G] x (this channel block's [3][128 x 16] canonical no-swizzle K-major tiles, pack.pack_stem)
  int64_t w_gstride;       // elements between the groups' tiles
  int chan_base;           // first channel of this launch's block (0, or 64 for the second block of C0 > 64)
  int cb;                  // row slots per parity that hold distinct channels: min(C0pad, 64) -> 32 (two replicas) or 64
  const float* scale;      // [G][C0pad]  (>= 0, sign folded into the weights)
  const float* bias;       // [G][C0pad]
  View<__nv_bfloat16> y;
  int n_img, group_imgs, groups, Hi, Wi, Ho, Wo, C0, C0pad;
  int tiles_x, tiles_y, total_tiles;
  FastDiv div_img, div_row; // tiles per image / per tile row
  int use_tma;             // inputs through TMA boxes (else: plain loads with explicit zero padding)
  int box_c;               // channels per output slab row (min(16, C0) rounded to 8)
  int dbg;                 // experiments only (DCFA_STEM_DBG): 1 no stores, 2 no input loads, 4 no MMAs, 16 half the MMAs
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
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
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x4s(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x2(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// no-swizzle K-major shared-memory descriptor (layout type 0): LBO = bytes between K chunks, SBO = between 8-row groups
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
// uint8 boxes must start on a 16-byte boundary: 3 * (2 * 12 * tx - L) = 72 tx - 3 L (and 24 tx - L for a single plane) is a
// multiple of 16 for L = 16 on even tile columns and L = 8 on odd ones.  L = pixels between the box start and pooled column
// px0's first conv input column + 2; the patch starts L - 2 pixels into the box.
__device__ __forceinline__ int u8_left(int tx) { return (tx & 1) ? 8 : 16; }

// Every CTA owns a CONTIGUOUS range of tiles (ordered image, tile row, tile column): consecutive tiles share their halo
// columns/rows through L2, and a tile's coordinates are two multiply-high divisions of the linear index.
struct Tile {
  int n, ty, tx;
};
__device__ __forceinline__ Tile tile_of(uint32_t t, const FastDiv& per_img, const FastDiv& per_row) {
  Tile r;
  r.n = (int)per_img.div(t);
  const uint32_t rem = t - (uint32_t)r.n * per_img.d;
  r.ty = (int)per_row.div(rem);
  r.tx = (int)(rem - (uint32_t)r.ty * per_row.d);
  return r;
}

__device__ __forceinline__ uint32_t u8x2_bf16(uint32_t lo, uint32_t hi) {   // two integers 0..255 -> packed bf16 (exact)
  return pack_bf16x2((float)lo, (float)hi);
}
// relu(lo), relu(hi) -> packed bf16 (one instruction)
__device__ __forceinline__ uint32_t relu_pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// lean spin on an mbarrier phase (suspend-time hint: the hardware parks the warp between polls); a pipeline bug traps
// after ~2^24 polls instead of hanging the GPU
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait_hint(bar, parity, 100000u)) {
    if (++spins > (1u << 24)) {
      printf("dcfa: stem mbarrier watchdog: block %d thread %d bar 0x%x parity %u\n", (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kStemThreads, 1) stem_kernel(const __grid_constant__ CUtensorMap map0,
                                                              const __grid_constant__ CUtensorMap map1,
                                                              const __grid_constant__ CUtensorMap map_y, const StemArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  // layout: A weight tiles (2 groups x 24 KB) | raw ring | converted ring | output slabs (16 warps x 2) | barriers | tmem slot
  const uint32_t s_a = base;
  const uint32_t s_raw = s_a + 2u * A_GROUP_BYTES;
  const uint32_t s_cvt = s_raw + (uint32_t)(NRAW * RAW_SLOT);
  const uint32_t s_out = s_cvt + (uint32_t)(NCVT * CVT_SLOT);
  const uint32_t bars = s_out + (uint32_t)(kEpiWarps * 2 * OUT_SLAB);
  const uint32_t bar_raw_full = bars, bar_raw_empty = bars + 8u * NRAW;
  const uint32_t bar_cvt_full = bars + 16u * NRAW, bar_cvt_empty = bar_cvt_full + 8u * NCVT;
  const uint32_t bar_tm_full = bar_cvt_empty + 8u * NCVT, bar_tm_empty = bar_tm_full + 8u * NACC;
  const uint32_t tmem_slot = bar_tm_empty + 8u * NACC;
  uint8_t* raw_ptr = gbase + 2 * A_GROUP_BYTES;
  uint8_t* cvt_ptr = raw_ptr + NRAW * RAW_SLOT;
  uint8_t* out_ptr = cvt_ptr + NCVT * CVT_SLOT;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(out_ptr + kEpiWarps * 2 * OUT_SLAB + (tmem_slot - bars));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int i = 0; i < NRAW; ++i) { ptx::mbar_init(bar_raw_full + 8u * i, 1); ptx::mbar_init(bar_raw_empty + 8u * i, 1); }
      for (int i = 0; i < NCVT; ++i) { ptx::mbar_init(bar_cvt_full + 8u * i, 1); ptx::mbar_init(bar_cvt_empty + 8u * i, 1); }
      for (int i = 0; i < NACC; ++i) { ptx::mbar_init(bar_tm_full + 8u * i, 1); ptx::mbar_init(bar_tm_empty + 8u * i, 4); }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  if (warp == kTmaWarp && lane == 0) {
    if (p.use_tma) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map0)) : "memory");
      if (p.groups > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map1)) : "memory");
    }
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_y)) : "memory");
  }
  {  // weight tiles of every group (constant parameters: no dependency on the previous kernel) and zeroed patch slots
    uint4* dst = reinterpret_cast<uint4*>(gbase);
    for (int g = 0; g < p.groups; ++g) {
      const uint4* src = reinterpret_cast<const uint4*>(p.w + (int64_t)g * p.w_gstride);
      for (int i = tid; i < A_GROUP_BYTES / 16; i += kStemThreads) dst[g * (A_GROUP_BYTES / 16) + i] = __ldg(src + i);
    }
    uint4* cz = reinterpret_cast<uint4*>(cvt_ptr);
    for (int i = tid; i < NCVT * CVT_SLOT / 16; i += kStemThreads) cz[i] = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's operand reads
  }
  ptx::pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  ptx::pdl_wait();

  // this CTA's tiles: [t_begin, t_end), a balanced contiguous split of the launch
  const uint32_t t_begin = (uint32_t)(((uint64_t)p.total_tiles * blockIdx.x) / gridDim.x);
  const uint32_t t_end = (uint32_t)(((uint64_t)p.total_tiles * (blockIdx.x + 1)) / gridDim.x);
  const uint32_t n_tiles = t_end - t_begin;

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer (whole warp runs the loop, one lane issues)
    if (p.use_tma && !(p.dbg & 2)) {
      const bool leader = ptx::elect_one();
      for (uint32_t i = 0; i < n_tiles; ++i) {
        const uint32_t s = i & (NRAW - 1), ph = (i / NRAW) & 1u;
        const Tile t = tile_of(t_begin + i, p.div_img, p.div_row);
        const uint32_t full = bar_raw_full + 8u * s, dst = s_raw + s * (uint32_t)RAW_SLOT;
        wait_bar(bar_raw_empty + 8u * s, ph ^ 1u);
        const int g = t.n >= p.group_imgs ? 1 : 0;
        const int nl = t.n - g * p.group_imgs;
        const int y = 2 * t.ty * TPH - 2;
        // (no pointer select between the two maps: that would copy a __grid_constant__ parameter to local memory)
        if (leader) {
          if (MODE == MODE_F32) {
            const int x = 2 * t.tx * TPW - 2 - XOFF;
            ptx::mbar_arrive_expect_tx(full, RAW_F32_BYTES);
            if (g == 0) tma_load_4d(dst, &map0, x, y, 0, nl, full);
            else tma_load_4d(dst, &map1, x, y, 0, nl, full);
          } else if (MODE == MODE_U8 || g == 0) {
            const int xb = 3 * (2 * t.tx * TPW - u8_left(t.tx));
            ptx::mbar_arrive_expect_tx(full, RAW_U8_BYTES);
            if (g == 0) tma_load_3d(dst, &map0, xb, y, nl, full);
            else tma_load_3d(dst, &map1, xb, y, nl, full);
          } else {
            const int xb = 2 * t.tx * TPW - u8_left(t.tx);
            ptx::mbar_arrive_expect_tx(full, RAW_C1_BYTES);
            tma_load_3d(dst, &map1, xb, y, nl, full);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (whole warp runs the loop, one lane issues)
    const bool leader = ptx::elect_one();
    const uint32_t idesc = ptx::make_idesc_bf16_f32(128, NCOL);
    // descriptors: only the 14-bit start-address field changes
    const uint64_t bd_hi = desc_nosw(0u, 16u, 128u);    // operand rows: row n = 16 bf16 at pixel 2n (linear): LBO 16, SBO 128
    const uint64_t ad_hi = desc_nosw(0u, 128u, 256u);   // canonical weight tiles
    const uint32_t a_lo0 = (s_a & 0x3FFFFu) >> 4, b_lo0 = (s_cvt & 0x3FFFFu) >> 4;
    const uint32_t t_group1 = (uint32_t)p.group_imgs * p.div_img.d;   // first tile of the second modality
    for (uint32_t i = 0; i < n_tiles; ++i) {
      const uint32_t ab = i & (NACC - 1), aph = (i / NACC) & 1u, cs = i & (NCVT - 1), cph = (i / NCVT) & 1u;
      const uint32_t g = (t_begin + i) >= t_group1 ? 1u : 0u;
      wait_bar(bar_tm_empty + 8u * ab, aph ^ 1u);
      wait_bar(bar_cvt_full + 8u * cs, cph);
      ptx::tc_fence_after();
      if (leader) {
        const uint32_t b0 = b_lo0 + cs * (uint32_t)(CVT_SLOT / 16);
        const uint32_t a0 = a_lo0 + g * (uint32_t)(A_GROUP_BYTES / 16);
        const uint32_t d0 = tmem_base + ab * (uint32_t)NCOL;
        if (!(p.dbg & 4)) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
            ptx::umma_bf16(d0, ad_hi | (uint64_t)(a0 + (uint32_t)(ky * A_TILE_BYTES / 16)),
                           bd_hi | (uint64_t)(b0 + (uint32_t)(ky * PP * 8 / 16)), idesc, ky > 0 ? 1u : 0u);
        }
        ptx::umma_commit(bar_cvt_empty + 8u * cs);   // the converted patch may be overwritten
        ptx::umma_commit(bar_tm_full + 8u * ab);     // the accumulator of this tile is complete
      }
      __syncwarp();
    }
  } else if (warp >= kCvtWarp0) {
    // ------------------------------------------------------------------ converters: raw patch -> bf16 [y][x][4]
    // Converter warp w owns this CTA's tiles w, w + 4, ... and converted slot w; tile i arrives in raw slot i % NRAW.
    const int cw = warp - kCvtWarp0;
    const bool tma_in = p.use_tma && !(p.dbg & 2);
    uint8_t* cvt = cvt_ptr + cw * CVT_SLOT;
    const uint32_t b_cvt_full = bar_cvt_full + 8u * cw, b_cvt_empty = bar_cvt_empty + 8u * cw;
    // tile-invariant item geometry of this lane.  fp32: item = (patch row r, pixel pair q), 154 items = 5 per lane;
    // uint8: item = (patch row r, 4 pixels q), 77 items = 3 per lane
    constexpr int NI_F = (PH * HP + 31) / 32, NI_U = (PH * (PP / 4) + 31) / 32;
    int rr[NI_F], qq[NI_F];
#pragma unroll
    for (int k = 0; k < NI_F; ++k) {
      const int item = lane + 32 * k;
      const int per = MODE == MODE_F32 ? HP : PP / 4;
      rr[k] = item / per;
      qq[k] = item - rr[k] * per;
    }
    uint32_t ph = 0;
    for (uint32_t i = (uint32_t)cw; i < n_tiles; i += kCvtWarps, ph ^= 1u) {
      const uint32_t rs = i & (NRAW - 1), rph = (i / NRAW) & 1u;   // raw slot of tile i and its use parity
      const uint8_t* raw = raw_ptr + rs * RAW_SLOT;
      const uint32_t b_raw_full = bar_raw_full + 8u * rs, b_raw_empty = bar_raw_empty + 8u * rs;
      wait_bar(b_cvt_empty, ph ^ 1u);
      if (tma_in) wait_bar(b_raw_full, rph);
      if (MODE == MODE_F32 && p.use_tma) {
        // three 64-bit loads (one per channel plane) and one 128-bit store per item; all loads of the lane's items are
        // issued before the first convert (independent chains).  item = (patch row, one of the row's 16 float pairs):
        // a half warp reads the 128 contiguous bytes of ONE raw row (conflict-free; with 14 pairs per row the half warps
        // straddled rows and every LDS.64 cost 4 wavefronts instead of 2); pairs 0 and 15 are the box's alignment padding.
        constexpr int NI_T = (PH * 16 + 31) / 32;   // 6
        float2 c[NI_T][3];
#pragma unroll
        for (int k = 0; k < NI_T; ++k) {
          const int item = lane + 32 * k, r16 = item >> 4, p16 = item & 15;
          if (k < NI_T - 1 || r16 < PH) {
            const float* src = reinterpret_cast<const float*>(raw) + r16 * PWB + 2 * p16;
            c[k][0] = *reinterpret_cast<const float2*>(src);
            c[k][1] = *reinterpret_cast<const float2*>(src + PH * PWB);
            c[k][2] = *reinterpret_cast<const float2*>(src + 2 * PH * PWB);
          }
        }
#pragma unroll
        for (int k = 0; k < NI_T; ++k) {
          const int item = lane + 32 * k, r16 = item >> 4, p16 = item & 15;
          if ((k < NI_T - 1 || r16 < PH) && p16 >= XOFF / 2 && p16 < XOFF / 2 + HP) {
            const uint4 o4 = make_uint4(pack_bf16x2(c[k][0].x, c[k][1].x), pack_bf16x2(c[k][2].x, 0.f),
                                        pack_bf16x2(c[k][0].y, c[k][1].y), pack_bf16x2(c[k][2].y, 0.f));
            *reinterpret_cast<uint4*>(cvt + (r16 * PP + 2 * (p16 - XOFF / 2)) * 8) = o4;
          }
        }
      } else {
        const Tile t = tile_of(t_begin + i, p.div_img, p.div_row);
        const int g = t.n >= p.group_imgs ? 1 : 0;
        const int nl = t.n - g * p.group_imgs;
        const int iy0 = 2 * t.ty * TPH - 2, ix0 = 2 * t.tx * TPW - 2;   // image coordinates of patch pixel (0, 0)
        if (MODE == MODE_F32) {
          // image rows are not 16-byte multiples: plain loads with explicit zero padding
          const float* img = reinterpret_cast<const float*>(g == 0 ? p.x[0] : p.x[1]) + (int64_t)nl * 3 * p.Hi * p.Wi;
          const int64_t pl = (int64_t)p.Hi * p.Wi;
#pragma unroll
          for (int k = 0; k < NI_F; ++k) {
            if (lane + 32 * k < PH * HP) {
              const int iy = iy0 + rr[k], ix = ix0 + 2 * qq[k];
              const bool yok = iy >= 0 && iy < p.Hi, a = yok && ix >= 0 && ix < p.Wi, b = yok && ix + 1 >= 0 && ix + 1 < p.Wi;
              const int64_t o = (int64_t)iy * p.Wi + ix;
              const float2 c0 = make_float2(a ? __ldg(img + o) : 0.f, b ? __ldg(img + o + 1) : 0.f);
              const float2 c1 = make_float2(a ? __ldg(img + pl + o) : 0.f, b ? __ldg(img + pl + o + 1) : 0.f);
              const float2 c2 = make_float2(a ? __ldg(img + 2 * pl + o) : 0.f, b ? __ldg(img + 2 * pl + o + 1) : 0.f);
              const uint4 o4 = make_uint4(pack_bf16x2(c0.x, c1.x), pack_bf16x2(c2.x, 0.f), pack_bf16x2(c0.y, c1.y), pack_bf16x2(c2.y, 0.f));
              *reinterpret_cast<uint4*>(cvt + (rr[k] * PP + 2 * qq[k]) * 8) = o4;
            }
          }
        } else if (MODE == MODE_U8 || g == 0) {
          // 12 bytes -> 32 bytes per item.  0..255 are exact bf16 integers; preprocess_input's 1/255 (utils/utils.py:76-79)
          // is folded into the BN scale.
          const int left = 3 * (u8_left(t.tx) - 2);
#pragma unroll
          for (int k = 0; k < NI_U; ++k) {
            if (lane + 32 * k < PH * (PP / 4)) {
              const int r = rr[k], q = qq[k];
              uint32_t b[12];
              if (p.use_tma) {
                const uint16_t* src = reinterpret_cast<const uint16_t*>(raw + r * RAWB + left + 12 * q);
#pragma unroll
                for (int j = 0; j < 6; ++j) { const uint32_t h = src[j]; b[2 * j] = h & 0xffu; b[2 * j + 1] = h >> 8; }
              } else {
                const uint8_t* img = reinterpret_cast<const uint8_t*>(g == 0 ? p.x[0] : p.x[1]) + (int64_t)nl * p.Hi * p.Wi * 3;
                const int iy = iy0 + r;
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                  const int ix = ix0 + 4 * q + j / 3;
                  b[j] = (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) ? (uint32_t)__ldg(img + ((int64_t)iy * p.Wi + ix) * 3 + j % 3) : 0u;
                }
              }
              uint4* dst = reinterpret_cast<uint4*>(cvt + (r * PP + 4 * q) * 8);
              dst[0] = make_uint4(u8x2_bf16(b[0], b[1]), u8x2_bf16(b[2], 0u), u8x2_bf16(b[3], b[4]), u8x2_bf16(b[5], 0u));
              dst[1] = make_uint4(u8x2_bf16(b[6], b[7]), u8x2_bf16(b[8], 0u), u8x2_bf16(b[9], b[10]), u8x2_bf16(b[11], 0u));
            }
          }
        } else {
          // single uint8 plane (the depth image before cvtColor replicates it, utils/utils.py:14-19): 4 bytes -> 32 bytes,
          // each value written to the three channel slots -- identical to uploading the replicated image
          const int left = u8_left(t.tx) - 2;
#pragma unroll
          for (int k = 0; k < NI_U; ++k) {
            if (lane + 32 * k < PH * (PP / 4)) {
              const int r = rr[k], q = qq[k];
              uint32_t b[4];
              if (p.use_tma) {
                const uint16_t* src = reinterpret_cast<const uint16_t*>(raw + r * RAW1B + left + 4 * q);
                const uint32_t h0 = src[0], h1 = src[1];
                b[0] = h0 & 0xffu; b[1] = h0 >> 8; b[2] = h1 & 0xffu; b[3] = h1 >> 8;
              } else {
                const uint8_t* img = reinterpret_cast<const uint8_t*>(p.x[1]) + (int64_t)nl * p.Hi * p.Wi;
                const int iy = iy0 + r;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int ix = ix0 + 4 * q + j;
                  b[j] = (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) ? (uint32_t)__ldg(img + (int64_t)iy * p.Wi + ix) : 0u;
                }
              }
              uint4* dst = reinterpret_cast<uint4*>(cvt + (r * PP + 4 * q) * 8);
              const uint32_t p0 = u8x2_bf16(b[0], b[0]), p1 = u8x2_bf16(b[1], b[1]), p2 = u8x2_bf16(b[2], b[2]), p3 = u8x2_bf16(b[3], b[3]);
              dst[0] = make_uint4(p0, p0 & 0xffffu, p1, p1 & 0xffffu);
              dst[1] = make_uint4(p2, p2 & 0xffffu, p3, p3 & 0xffffu);
            }
          }
        }
      }
      ptx::fence_proxy_async_smem();   // this thread's st.shared -> visible to the async proxy (MMA operand reads)
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(b_cvt_full);
        if (tma_in) ptx::mbar_arrive(b_raw_empty);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: group = tile parity, warp & 3 = TMEM lane quarter
    // Lane l of quarter q: column parity e = l >> 4, channel lane cl = l & 15, row slot j = 16 q + cl.  cb = 64: slot = channel
    // (all four pooled rows per warp); cb = 32: slots 32..63 are a replica that owns the other two pooled rows.
    const int q4 = warp & 3, grp = warp >> 2;
    const int e = lane >> 4, cl = lane & 15;
    const int slot0 = q4 * 16;
    const int chan0 = p.chan_base + (p.cb == 64 ? slot0 : (slot0 & 31));   // first channel of this warp (multiple of 16)
    const int rep = p.cb == 64 ? 0 : (q4 >> 1), nrep = p.cb == 64 ? 1 : 2;   // replica rep owns the pooled row pairs rep, rep + nrep
    const int ch = chan0 + cl;
    const bool ch_valid = ch < p.C0;
    const bool warp_valid = chan0 < p.C0;
    const bool leader = ptx::elect_one();
    float sc[2], bi[2];
    sc[0] = __ldg(p.scale + ch); bi[0] = __ldg(p.bias + ch);
    sc[1] = p.groups > 1 ? __ldg(p.scale + p.C0pad + ch) : sc[0];
    bi[1] = p.groups > 1 ? __ldg(p.bias + p.C0pad + ch) : bi[0];
    const uint32_t slab0 = s_out + (uint32_t)(warp * 2 * OUT_SLAB);
    const bool dense16 = p.box_c == 16;           // slab row pitch 32 bytes: compile-time store offsets
    uint8_t* slab_ptr0 = out_ptr + warp * 2 * OUT_SLAB + cl * 2 + HW6 * e * (p.box_c * 2);   // this lane's first pixel: 6 e
    const uint32_t b_full0 = bar_tm_full, b_empty0 = bar_tm_empty;
    uint32_t slab_sel = 0;
    for (uint32_t i = (uint32_t)grp; i < n_tiles; i += NACC) {   // group grp owns accumulator stage grp
      const uint32_t ab = i & (NACC - 1), aph = (i / NACC) & 1u;
      const Tile t = tile_of(t_begin + i, p.div_img, p.div_row);
      const int g = t.n >= p.group_imgs ? 1 : 0;
      const float s = g ? sc[1] : sc[0], b = g ? bi[1] : bi[0];
      const int py0 = t.ty * TPH, px0 = t.tx * TPW;
      const bool border = t.ty == 0 || t.tx == 0 || 2 * py0 - 1 + CH > p.Hi || 2 * px0 - 1 + CW > p.Wi;
      wait_bar(b_full0 + 8u * ab, aph);
      ptx::tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + ab * (uint32_t)NCOL;
      for (int rp = rep; rp < TPH / 2; rp += nrep) {
        // pooled rows 2rp, 2rp+1 = conv rows 4rp .. 4rp+4 = 70 consecutive TMEM columns (56 rp ..); this lane's parity:
        // conv columns 2k + e.  Both rows leave through ONE slab and ONE TMA store (the per-store proxy fence and bulk-group
        // bookkeeping were ~15 % of the epilogue's time).
        const uint32_t tr = t0 + (uint32_t)(4 * HP * rp);
        const uint32_t slab = slab0 + slab_sel * (uint32_t)OUT_SLAB;
        uint8_t* sp = slab_ptr0 + slab_sel * OUT_SLAB;
        int cy0 = 2 * py0 - 1 + 4 * rp, cx0 = 2 * px0 - 1 + e;   // image coordinates of this lane's first conv pixel
        // the column maxima of one pooled row -> swap with the lane of the other parity -> pool, BN, ReLU -> slab row h.
        // The even lane finishes pooled pixels 0..5 (needs the odd maxima 0..5), the odd lane pixels 6..11 (needs the even
        // maxima 6..12): with P = the even maxima and Q = the odd ones of the lane's six pixels, out = max3(P[pc], P[pc+1], Q[pc]).
        auto finish_row = [&](const float (&v)[TPW + 1], int h) {
          float rcv[7];
#pragma unroll
          for (int k = 0; k < 7; ++k) rcv[k] = __shfl_xor_sync(0xffffffffu, e ? v[k] : v[6 + k], 16);
          float o[HW6];
#pragma unroll
          for (int pc = 0; pc < HW6; ++pc) {
            const float p0 = e ? rcv[pc] : v[pc], p1 = e ? rcv[pc + 1] : v[pc + 1], q0 = e ? v[6 + pc] : rcv[pc];
            o[pc] = fmaf(max3(p0, p1, q0), s, b);
          }
          if (ch_valid) {
            if (dense16) {
#pragma unroll
              for (int pc = 0; pc < HW6; pc += 2) {
                const uint32_t w = relu_pack_bf16x2(o[pc], o[pc + 1]);   // ReLU fused into the convert
                *reinterpret_cast<uint16_t*>(sp + h * (TPW * 32) + pc * 32) = (uint16_t)(w & 0xffffu);
                *reinterpret_cast<uint16_t*>(sp + h * (TPW * 32) + (pc + 1) * 32) = (uint16_t)(w >> 16);
              }
            } else {
              const int pitch = p.box_c * 2;
#pragma unroll
              for (int pc = 0; pc < HW6; ++pc)
                *reinterpret_cast<__nv_bfloat16*>(sp + h * (TPW * pitch) + pc * pitch) = __float2bfloat16_rn(fmaxf(o[pc], 0.0f));
            }
          }
        };
        // out-of-image conv positions are excluded from the max below (reference: -inf pool padding); border tiles only
        if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last read this slab
        __syncwarp();
        float mid[TPW + 1];   // conv row 4rp + 2 belongs to both pooled rows
        {
          uint32_t V[3 * HP];   // conv rows 4rp .. 4rp+2
          tmem_ld_x32(tr, V); tmem_ld_x8(tr + 32u, V + 32); tmem_ld_x2(tr + 40u, V + 40);
          ptx::tmem_ld_wait();
          if (border) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const bool yok = cy0 + j >= 0 && cy0 + j < p.Hi;
#pragma unroll
              for (int k = 0; k < TPW + 1; ++k)
                if (!(yok && cx0 + 2 * k >= 0 && cx0 + 2 * k < p.Wi)) V[j * HP + k] = 0xff800000u;
            }
          }
          float v[TPW + 1];
#pragma unroll
          for (int k = 0; k < TPW + 1; ++k) {
            mid[k] = __uint_as_float(V[2 * HP + k]);
            v[k] = max3(__uint_as_float(V[k]), __uint_as_float(V[HP + k]), mid[k]);
          }
          finish_row(v, 0);
        }
        {
          uint32_t V[2 * HP];   // conv rows 4rp+3, 4rp+4
          ptx::tmem_ld_x16(tr + 42u, V); tmem_ld_x8(tr + 58u, V + 16); tmem_ld_x4s(tr + 66u, V + 24);
          ptx::tmem_ld_wait();
          if (rp + nrep >= TPH / 2) {   // last rows of this thread: the accumulator may be overwritten by a later tile
            ptx::tc_fence_before();
            __syncwarp();
            if (leader) ptx::mbar_arrive(b_empty0 + 8u * ab);
          }
          if (border) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const bool yok = cy0 + 3 + j >= 0 && cy0 + 3 + j < p.Hi;
#pragma unroll
              for (int k = 0; k < TPW + 1; ++k)
                if (!(yok && cx0 + 2 * k >= 0 && cx0 + 2 * k < p.Wi)) V[j * HP + k] = 0xff800000u;
            }
          }
          float v[TPW + 1];
#pragma unroll
          for (int k = 0; k < TPW + 1; ++k) v[k] = max3(mid[k], __uint_as_float(V[k]), __uint_as_float(V[HP + k]));
          finish_row(v, 1);
        }
        // [2 rows][12 pixels][box_c channels] slab -> one TMA store (clipped at the image edges and at C0 by the tensor map)
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (leader) {
          const int py = py0 + 2 * rp;
          if (warp_valid && py < p.Ho && !(p.dbg & 1)) tma_store_4d(&map_y, slab, chan0, px0, py, t.n);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        slab_sel ^= 1u;
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // slabs must outlive their stores
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn stem_encode_fn() {
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

}  // namespace

int launch_stem(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  StemArgs a;
  const bool u8 = (op.flags & DCFA_STEM_FLAG_U8) != 0;
  const bool c1 = (op.flags & DCFA_STEM_FLAG_X2_PLANE) != 0;
  a.x[0] = resolve_ptr<const uint8_t>(op.x, bufs);
  a.x[1] = resolve_ptr<const uint8_t>(op.x2, bufs);
  a.w = resolve_ptr<const __nv_bfloat16>(op.w, bufs);
  a.scale = resolve_ptr<const float>(op.scale, bufs);
  a.bias = resolve_ptr<const float>(op.bias, bufs);
  a.y = resolve<__nv_bfloat16>(op.y, bufs);
  a.n_img = op.n_img;
  a.group_imgs = op.group_imgs > 0 ? op.group_imgs : op.n_img;
  a.Hi = op.Hi; a.Wi = op.Wi; a.Ho = op.Ho; a.Wo = op.Wo; a.C0 = op.Cout; a.C0pad = op.BN;
  DCFA_REQUIRE(a.x[0] && a.w && a.scale && a.bias && a.y.p, "stem: missing tensor");
  DCFA_REQUIRE(a.n_img == a.group_imgs || (a.n_img == 2 * a.group_imgs && a.x[1]), "stem: needs 1 or 2 groups");
  DCFA_REQUIRE(a.Hi > 0 && a.Wi > 0 && a.Ho == (a.Hi - 1) / 2 + 1 && a.Wo == (a.Wi - 1) / 2 + 1,
               "stem: pooled size %dx%d inconsistent with %dx%d", a.Ho, a.Wo, a.Hi, a.Wi);
  DCFA_REQUIRE(a.C0 >= 1 && a.C0 <= 128, "stem: C0 %d unsupported", a.C0);
  DCFA_REQUIRE((a.C0pad == 32 || a.C0pad == 64 || a.C0pad == 128) && a.C0pad >= a.C0 && op.K_real == 27,
               "stem: weight packing mismatch (C0pad %d, C0 %d)", a.C0pad, a.C0);
  const int nblk = a.C0pad > 64 ? a.C0pad / 64 : 1;   // channel blocks of 64 (one launch each)
  DCFA_REQUIRE(op.w_gstride == (int64_t)nblk * 3 * 128 * 16 && op.sb_gstride == a.C0pad,
               "stem: weight tiles must be [G][%d][3][128x16]", nblk);
  a.w_gstride = op.w_gstride;
  a.cb = a.C0pad < 64 ? 32 : 64;
  DCFA_REQUIRE(((uintptr_t)a.w % 16) == 0, "stem: weights must be 16-byte aligned");
  a.groups = a.n_img / a.group_imgs;
  DCFA_REQUIRE(!c1 || (u8 && a.groups == 2), "stem: the single-plane flag needs uint8 inputs and two groups");
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("DCFA_STEM_DBG"); dbg = e ? atoi(e) : 0; }
    a.dbg = dbg;
  }
  a.tiles_x = ceil_div(a.Wo, TPW);
  a.tiles_y = ceil_div(a.Ho, TPH);
  const int64_t total = (int64_t)a.n_img * a.tiles_x * a.tiles_y;
  DCFA_REQUIRE(total < (1ll << 31), "stem: too many tiles");
  a.total_tiles = (int)total;
  a.div_img = make_fastdiv((uint32_t)(a.tiles_x * a.tiles_y));
  a.div_row = make_fastdiv((uint32_t)a.tiles_x);

  // ---- tensor maps, zero fill outside the image:
  //      fp32 NCHW : dims (W, H, C, N), box (PWB, 9, 3, 1)      uint8 NHWC : dims (3W bytes, H, N), box (RAWB, 9, 1)
  //      uint8 plane : dims (W, H, N), box (RAW1B, 9, 1)
  alignas(64) CUtensorMap maps[2];
  memset(maps, 0, sizeof(maps));
  bool ok = ((uintptr_t)a.x[0] % 16) == 0 && (a.groups == 1 || ((uintptr_t)a.x[1] % 16) == 0);
  if (u8) ok = ok && (3 * a.Wi) % 16 == 0 && (!c1 || a.Wi % 16 == 0);
  else ok = ok && a.Wi % 4 == 0;
  a.use_tma = ok ? 1 : 0;
  if (a.use_tma) {
    EncodeTiledFn enc = stem_encode_fn();
    DCFA_REQUIRE(enc != nullptr, "stem: cuTensorMapEncodeTiled entry point unavailable");
    for (int g = 0; g < a.groups; ++g) {
      const cuuint32_t es[4] = {1u, 1u, 1u, 1u};
      CUresult cr;
      if (u8 && !(c1 && g == 1)) {
        const cuuint64_t gdim[3] = {(cuuint64_t)a.Wi * 3, (cuuint64_t)a.Hi, (cuuint64_t)a.group_imgs};
        const cuuint64_t gstr[2] = {(cuuint64_t)a.Wi * 3, (cuuint64_t)a.Wi * a.Hi * 3};
        const cuuint32_t box[3] = {(cuuint32_t)RAWB, (cuuint32_t)PH, 1u};
        cr = enc(&maps[g], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(a.x[g]), gdim, gstr, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else if (u8) {
        const cuuint64_t gdim[3] = {(cuuint64_t)a.Wi, (cuuint64_t)a.Hi, (cuuint64_t)a.group_imgs};
        const cuuint64_t gstr[2] = {(cuuint64_t)a.Wi, (cuuint64_t)a.Wi * a.Hi};
        const cuuint32_t box[3] = {(cuuint32_t)RAW1B, (cuuint32_t)PH, 1u};
        cr = enc(&maps[g], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(a.x[g]), gdim, gstr, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {
        const cuuint64_t gdim[4] = {(cuuint64_t)a.Wi, (cuuint64_t)a.Hi, 3, (cuuint64_t)a.group_imgs};
        const cuuint64_t gstr[3] = {(cuuint64_t)a.Wi * 4, (cuuint64_t)a.Wi * a.Hi * 4, (cuuint64_t)a.Wi * a.Hi * 12};
        const cuuint32_t box[4] = {(cuuint32_t)PWB, (cuuint32_t)PH, 3u, 1u};
        cr = enc(&maps[g], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(a.x[g]), gdim, gstr, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      if (cr != CUDA_SUCCESS) return fail(DCFA_E_CUDA, "stem: cuTensorMapEncodeTiled failed with %d", (int)cr);
    }
  }
  // ---- output tensor map: dims (C0, Wo, Ho, N) over the NHWC view, box (box_c channels, 12 pixels, 1, 1)
  alignas(64) CUtensorMap map_y;
  memset(&map_y, 0, sizeof(map_y));
  a.box_c = a.C0 >= 16 ? 16 : (a.C0 + 7) / 8 * 8;   // channels per epilogue warp (16 channel lanes x 2 parities)
  DCFA_REQUIRE(a.C0 % 8 == 0 && ((uintptr_t)a.y.p % 16) == 0 && a.y.ld % 8 == 0 && a.y.img_stride % 8 == 0 &&
                   (a.y.gi <= 0 || a.y.gstride == (int64_t)a.y.gi * a.y.img_stride),
               "stem: the output view must be 16-byte aligned with C0 %% 8 == 0 and a uniform image stride (TMA store)");
  {
    EncodeTiledFn enc = stem_encode_fn();
    DCFA_REQUIRE(enc != nullptr, "stem: cuTensorMapEncodeTiled entry point unavailable");
    const cuuint64_t ydim[4] = {(cuuint64_t)a.C0, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)a.n_img};
    const cuuint64_t ystr[3] = {(cuuint64_t)a.y.ld * 2, (cuuint64_t)a.Wo * a.y.ld * 2, (cuuint64_t)a.y.img_stride * 2};
    const cuuint32_t ybox[4] = {(cuuint32_t)a.box_c, (cuuint32_t)TPW, 2u, 1u};
    const cuuint32_t yes[4] = {1u, 1u, 1u, 1u};
    CUresult cr = enc(&map_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.y.p, ydim, ystr, ybox, yes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(DCFA_E_CUDA, "stem: cuTensorMapEncodeTiled(output) failed with %d", (int)cr);
  }
  const size_t smem = 1024 + 2 * A_GROUP_BYTES + NRAW * RAW_SLOT + NCVT * CVT_SLOT + kEpiWarps * 2 * OUT_SLAB + 256;
  static DeviceOnce attr_set;
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(stem_kernel<MODE_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_kernel<MODE_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_kernel<MODE_U8_C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(DCFA_E_CUDA, "stem: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.mark();
  }
  int64_t grid = sm_count();   // one CTA per SM: the kernel owns all 512 TMEM columns
  if (grid > total) grid = total;
  const __nv_bfloat16* w0 = a.w;
  for (int blk = 0; blk < nblk; ++blk) {
    a.w = w0 + (int64_t)blk * 3 * 128 * 16;
    a.chan_base = 64 * blk;
    if (!u8) launch_pdl(stem_kernel<MODE_F32>, dim3((unsigned)grid), dim3(kStemThreads), smem, st, maps[0], maps[1], map_y, a);
    else if (!c1) launch_pdl(stem_kernel<MODE_U8>, dim3((unsigned)grid), dim3(kStemThreads), smem, st, maps[0], maps[1], map_y, a);
    else launch_pdl(stem_kernel<MODE_U8_C1>, dim3((unsigned)grid), dim3(kStemThreads), smem, st, maps[0], maps[1], map_y, a);
    if (blk + 1 < nblk) DCFA_CHECK_LAUNCH("stem_kernel");
  }
  DCFA_CHECK_LAUNCH("stem_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
