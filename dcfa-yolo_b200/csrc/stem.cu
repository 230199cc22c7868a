// DCFA_OP_STEM: Conv_maxpool (nets/yolo_mul.py:104-115) fused into one pass on tcgen05 tensor cores:
//   image -> conv3x3 s1 p1 (3 -> C0) -> BN scale/bias -> ReLU -> maxpool 3x3 s2 p1 -> bf16 NHWC.
// The full-resolution conv map (the largest tensor of the network) never leaves the SM.  Both modalities run in one
// launch (weight group = image / group_imgs).
//
// Round-1 version (ncu: 41 % issue-active, 0.97 TB/s): a generic-proxy im2col (27 LDS + converts + 4 STS per conv
// pixel) and four CTA-wide phases per tile serialised behind __syncthreads.  This version has NO im2col:
//
//   * the input patch of a tile (9 x 36 pixels) is converted ONCE to bf16 [y][x][4 channels] (8 bytes per pixel,
//     channel 3 = 0) -- 1.3 shared-memory operations per input pixel instead of ~35 per conv pixel;
//   * the tcgen05 shared-memory descriptor does the im2col: NO-SWIZZLE K-major layout with LBO = 16 B (next K chunk =
//     next pixel pair) and SBO = 128 B (next 8 rows), i.e. operand row n = the 16 bf16 starting at pixel 2n of the
//     LINEAR patch (pitch 36): four consecutive pixels x 4 channels.  Rows overlap in memory; the hardware only
//     computes addresses (tools/umma_overlap_test.cu proves it).  One K=16 MMA per kernel row and per pixel PARITY:
//     the even conv pixel 2n uses weights [w(ky,0) | w(ky,1) | w(ky,2) | 0], the odd conv pixel 2n+1 uses
//     [0 | w(ky,0) | w(ky,1) | w(ky,2)] over the SAME operand rows, into a second accumulator;
//   * the GEMM is turned so that CHANNELS are the 128 MMA rows (C0 replicated 128/C0pad times) and pixels are TMEM
//     columns: after tcgen05.ld a thread holds a window of ONE channel and the 3x3/2 max-pool is pure register
//     arithmetic with three-input max; BN + ReLU after pooling (exact: the sign of the BN scale is folded into the
//     weights, so the pooled quantity is monotone in the accumulator);
//   * warp-specialised pipeline, one CTA per SM, all of TMEM (2 tiles x 2 parities x 128 columns):
//       warp 12      TMA producer: raw patch ring (fp32 NCHW planes, uint8 NHWC rows or a uint8 depth plane)
//       warps 8-11   converters: raw -> bf16 [y][x][4] (ring of 3), fence.proxy.async, arrive
//       warp 13      MMA issuer: 6 x (M128 N128 K16) per tile, commits free the converted slot / publish the accumulators
//       warps 0-7    two epilogue groups alternating tiles: tcgen05.ld, pool, BN, ReLU, 64-byte-per-pixel stores.
// Tile: 3 x 16 pooled pixels = 7 x 33 conv pixels (linear index L = cy*36 + cx, parity = cx & 1, TMEM column L >> 1).
// Out-of-image conv positions are excluded from the max (reference: -inf pool padding); zero padding of the conv comes
// from the TMA's out-of-bounds fill.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {
namespace {

constexpr int TPH = 3, TPW = 16;              // pooled tile
constexpr int CH = 2 * TPH + 1;               // 7 conv rows
constexpr int CW = 2 * TPW + 1;               // 33 conv cols
constexpr int PH = CH + 2;                    // 9 patch rows
constexpr int PP = 36;                        // patch pitch in pixels (35 needed; even, so that parity(L) = parity(cx))
constexpr int NCOL = 128;                     // MMA N = pixel pairs per parity (7 * 18 = 126 used)
constexpr int XOFF = 2;                       // fp32: the TMA box starts 2 floats left of the patch (16-byte aligned start)
constexpr int PWB = 40;                       // fp32 raw row pitch in floats (XOFF + 36, rounded to 16 bytes)
constexpr int RAW_F32_BYTES = 3 * PH * PWB * 4;   // 4320
constexpr int U8_LEFT = 16;                   // uint8: the box starts 16 pixels left of the tile's first pooled pixel column * 2
constexpr int RAWB = 160;                     // uint8 NHWC raw row bytes: 3 * (14 + 36) = 150, rounded to 16
constexpr int RAW_U8_BYTES = RAWB * PH;       // 1440
constexpr int RAW1B = 64;                     // uint8 single plane raw row bytes: 14 + 36 = 50, rounded to 16
constexpr int RAW_C1_BYTES = RAW1B * PH;      // 576
constexpr int RAW_SLOT = 4352;                // bytes per raw ring slot (>= 4320, multiple of 128)
constexpr int NRAW = 4;                       // raw ring depth
constexpr int CVT_BYTES = PH * PP * 8;        // 2592: bf16 [9][36][4]
constexpr int CVT_SLOT = 2688;                // + the rows the last MMA over-reads (columns 126, 127: never used), mult. of 128
constexpr int NCVT = 3;                       // converted-patch ring depth
constexpr int A_TILE_BYTES = 128 * 16 * 2;    // one (kernel row, parity) weight tile
constexpr int A_GROUP_BYTES = 6 * A_TILE_BYTES;
constexpr int kEpiWarps = 8, kCvtWarps = 4;
constexpr int kCvtWarp0 = kEpiWarps, kTmaWarp = kEpiWarps + kCvtWarps, kMmaWarp = kTmaWarp + 1;
constexpr int kStemThreads = (kMmaWarp + 1) * 32;   // 448
constexpr uint32_t kTmemCols = 512;
static_assert(2 * PP * 8 + (NCOL - 1) * 16 + 32 <= CVT_SLOT, "MMA over-read must stay inside the slot");

enum { MODE_F32 = 0, MODE_U8 = 1, MODE_U8_C1 = 2 };   // MODE_U8_C1: group 0 uint8 NHWC, group 1 a single uint8 plane

struct StemArgs {
  const void* x[2];        // fp32 NCHW / uint8 NHWC / (group 1, MODE_U8_C1) uint8 [N,H,W]
  const __nv_bfloat16* w;  // [G][3][2][128 x 16] canonical no-swizzle K-major tiles (pack.pack_stem)
  const float* scale;      // [G][C0pad]  (>= 0, sign folded into the weights)
  const float* bias;       // [G][C0pad]
  View<__nv_bfloat16> y;
  int n_img, group_imgs, groups, Hi, Wi, Ho, Wo, C0, C0pad;
  int tiles_x, tiles_y;
  int use_tma;
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
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
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

// Tile coordinates advanced incrementally (tile index += gridDim.x): the per-tile path has no divisions.
// Tiles are ordered (image over both groups, tile row, tile column).
struct TileIter {
  int tx, ty, n;       // tile column, tile row, image index over both groups
  int sx, sy, sn;      // mixed-radix digits of the step gridDim.x
  __device__ __forceinline__ void init(int tile, int step, int tiles_x, int tiles_y) {
    const int per_img = tiles_x * tiles_y;
    n = tile / per_img;
    int r = tile - n * per_img;
    ty = r / tiles_x;
    tx = r - ty * tiles_x;
    sn = step / per_img;
    r = step - sn * per_img;
    sy = r / tiles_x;
    sx = r - sy * tiles_x;
  }
  __device__ __forceinline__ void advance(int tiles_x, int tiles_y) {
    tx += sx;
    if (tx >= tiles_x) { tx -= tiles_x; ty += 1; }
    ty += sy;
    if (ty >= tiles_y) { ty -= tiles_y; n += 1; }
    n += sn;
  }
};

__device__ __forceinline__ uint32_t u8x2_bf16(uint32_t lo, uint32_t hi) {   // two integers 0..255 -> packed bf16 (exact)
  return pack_bf16x2((float)lo, (float)hi);
}

template <int MODE>
__global__ void __launch_bounds__(kStemThreads, 1) stem_kernel(const __grid_constant__ CUtensorMap map0,
                                                              const __grid_constant__ CUtensorMap map1, const StemArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  // layout: A weight tiles (2 groups x 24 KB) | raw ring | converted ring | barriers | tmem slot
  const uint32_t s_a = base;
  const uint32_t s_raw = s_a + 2u * A_GROUP_BYTES;
  const uint32_t s_cvt = s_raw + (uint32_t)(NRAW * RAW_SLOT);
  const uint32_t bars = s_cvt + (uint32_t)(NCVT * CVT_SLOT);
  const uint32_t bar_raw_full = bars, bar_raw_empty = bars + 8u * NRAW;
  const uint32_t bar_cvt_full = bars + 16u * NRAW, bar_cvt_empty = bar_cvt_full + 8u * NCVT;
  const uint32_t bar_tm_full = bar_cvt_empty + 8u * NCVT, bar_tm_empty = bar_tm_full + 16u;
  const uint32_t tmem_slot = bar_tm_empty + 16u;
  uint8_t* raw_ptr = gbase + 2 * A_GROUP_BYTES;
  uint8_t* cvt_ptr = raw_ptr + NRAW * RAW_SLOT;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(cvt_ptr + NCVT * CVT_SLOT + (tmem_slot - bars));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int i = 0; i < NRAW; ++i) { ptx::mbar_init(bar_raw_full + 8u * i, 1); ptx::mbar_init(bar_raw_empty + 8u * i, kCvtWarps); }
      for (int i = 0; i < NCVT; ++i) { ptx::mbar_init(bar_cvt_full + 8u * i, kCvtWarps); ptx::mbar_init(bar_cvt_empty + 8u * i, 1); }
      for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar_tm_full + 8u * i, 1); ptx::mbar_init(bar_tm_empty + 8u * i, 4); }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  if (warp == kTmaWarp && lane == 0 && p.use_tma) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map0)) : "memory");
    if (p.groups > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map1)) : "memory");
  }
  {  // weight tiles of every group (constant parameters: no dependency on the previous kernel) and zeroed patch slots
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(gbase);
    const int n16 = p.groups * (A_GROUP_BYTES / 16);
    for (int i = tid; i < n16; i += kStemThreads) dst[i] = __ldg(src + i);
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

  TileIter cur;
  cur.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer (one thread)
    if (lane == 0 && p.use_tma) {
      uint32_t s = 0, ph = 0;
      for (; cur.n < p.n_img; cur.advance(p.tiles_x, p.tiles_y)) {
        const uint32_t full = bar_raw_full + 8u * s, dst = s_raw + s * (uint32_t)RAW_SLOT;
        ptx::mbar_wait(bar_raw_empty + 8u * s, ph ^ 1u);
        const int g = cur.n >= p.group_imgs ? 1 : 0;
        const int nl = cur.n - g * p.group_imgs;
        const int y = 2 * cur.ty * TPH - 2;
        // (no pointer select between the two maps: that would copy a __grid_constant__ parameter to local memory)
        if (MODE == MODE_F32) {
          const int x = 2 * cur.tx * TPW - 2 - XOFF;
          ptx::mbar_arrive_expect_tx(full, RAW_F32_BYTES);
          if (g == 0) tma_load_4d(dst, &map0, x, y, 0, nl, full);
          else tma_load_4d(dst, &map1, x, y, 0, nl, full);
        } else if (MODE == MODE_U8 || g == 0) {
          const int xb = 3 * (2 * cur.tx * TPW - U8_LEFT);
          ptx::mbar_arrive_expect_tx(full, RAW_U8_BYTES);
          if (g == 0) tma_load_3d(dst, &map0, xb, y, nl, full);
          else tma_load_3d(dst, &map1, xb, y, nl, full);
        } else {
          const int xb = 2 * cur.tx * TPW - U8_LEFT;
          ptx::mbar_arrive_expect_tx(full, RAW_C1_BYTES);
          tma_load_3d(dst, &map1, xb, y, nl, full);
        }
        if (++s == NRAW) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16_f32(128, NCOL);
      uint32_t cs = 0, cph = 0, it = 0;
      for (; cur.n < p.n_img; cur.advance(p.tiles_x, p.tiles_y), ++it) {
        const uint32_t ab = it & 1u, aph = (it >> 1) & 1u;
        const int g = cur.n >= p.group_imgs ? 1 : 0;
        ptx::mbar_wait(bar_tm_empty + 8u * ab, aph ^ 1u);
        ptx::mbar_wait(bar_cvt_full + 8u * cs, cph);
        ptx::tc_fence_after();
        const uint32_t b0 = s_cvt + cs * (uint32_t)CVT_SLOT;
        const uint32_t a0 = s_a + (uint32_t)g * A_GROUP_BYTES;
        const uint32_t d0 = tmem_base + ab * 256u;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          // operand rows: row n = 16 bf16 at pixel 2n of patch row ky (+ the tile's conv rows, linear): LBO 16, SBO 128
          const uint64_t bd = desc_nosw(b0 + (uint32_t)(ky * PP * 8), 16u, 128u);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint64_t ad = desc_nosw(a0 + (uint32_t)((ky * 2 + e) * A_TILE_BYTES), 128u, 256u);
            ptx::umma_bf16(d0 + (uint32_t)(e * NCOL), ad, bd, idesc, ky > 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(bar_cvt_empty + 8u * cs);   // the converted patch may be overwritten
        ptx::umma_commit(bar_tm_full + 8u * ab);     // both accumulators of this tile are complete
        if (++cs == NCVT) { cs = 0; cph ^= 1u; }
      }
    }
  } else if (warp >= kCvtWarp0) {
    // ------------------------------------------------------------------ converters: raw patch -> bf16 [y][x][4]
    const int ctid = tid - kCvtWarp0 * 32;
    uint32_t rs = 0, rph = 0, cs = 0, cph = 0;
    for (; cur.n < p.n_img; cur.advance(p.tiles_x, p.tiles_y)) {
      const int g = cur.n >= p.group_imgs ? 1 : 0;
      const int nl = cur.n - g * p.group_imgs;
      ptx::mbar_wait(bar_cvt_empty + 8u * cs, cph ^ 1u);
      if (p.use_tma) ptx::mbar_wait(bar_raw_full + 8u * rs, rph);
      const uint8_t* raw = raw_ptr + rs * RAW_SLOT;
      uint8_t* cvt = cvt_ptr + cs * CVT_SLOT;
      const int iy0 = 2 * cur.ty * TPH - 2, ix0 = 2 * cur.tx * TPW - 2;   // image coordinates of patch pixel (0, 0)
      if (MODE == MODE_F32) {
        // item = (patch row r, pixel pair q): three 64-bit loads (one per channel plane), one 128-bit store
        for (int item = ctid; item < PH * (PP / 2); item += kCvtWarps * 32) {
          const int r = item / (PP / 2), q = item - r * (PP / 2);
          float2 c0, c1, c2;
          if (p.use_tma) {
            const float* src = reinterpret_cast<const float*>(raw) + r * PWB + XOFF + 2 * q;
            c0 = *reinterpret_cast<const float2*>(src);
            c1 = *reinterpret_cast<const float2*>(src + PH * PWB);
            c2 = *reinterpret_cast<const float2*>(src + 2 * PH * PWB);
          } else {   // image rows are not 16-byte multiples: plain loads with explicit zero padding
            const float* img = reinterpret_cast<const float*>(g == 0 ? p.x[0] : p.x[1]) + (int64_t)nl * 3 * p.Hi * p.Wi;
            const int iy = iy0 + r, ix = ix0 + 2 * q;
            const bool yok = iy >= 0 && iy < p.Hi, a = yok && ix >= 0 && ix < p.Wi, b = yok && ix + 1 >= 0 && ix + 1 < p.Wi;
            const int64_t o = (int64_t)iy * p.Wi + ix, pl = (int64_t)p.Hi * p.Wi;
            c0 = make_float2(a ? __ldg(img + o) : 0.f, b ? __ldg(img + o + 1) : 0.f);
            c1 = make_float2(a ? __ldg(img + pl + o) : 0.f, b ? __ldg(img + pl + o + 1) : 0.f);
            c2 = make_float2(a ? __ldg(img + 2 * pl + o) : 0.f, b ? __ldg(img + 2 * pl + o + 1) : 0.f);
          }
          const uint4 o4 = make_uint4(pack_bf16x2(c0.x, c1.x), pack_bf16x2(c2.x, 0.f), pack_bf16x2(c0.y, c1.y), pack_bf16x2(c2.y, 0.f));
          *reinterpret_cast<uint4*>(cvt + (r * PP + 2 * q) * 8) = o4;
        }
      } else if (MODE == MODE_U8 || g == 0) {
        // item = (patch row r, 4 pixels): 12 bytes -> 32 bytes.  0..255 are exact bf16 integers; preprocess_input's 1/255
        // (utils/utils.py:76-79) is folded into the BN scale.
        for (int item = ctid; item < PH * (PP / 4); item += kCvtWarps * 32) {
          const int r = item / (PP / 4), q = item - r * (PP / 4);
          uint32_t b[12];
          if (p.use_tma) {
            const uint16_t* src = reinterpret_cast<const uint16_t*>(raw + r * RAWB + 3 * (U8_LEFT - 2) + 12 * q);
#pragma unroll
            for (int k = 0; k < 6; ++k) { const uint32_t h = src[k]; b[2 * k] = h & 0xffu; b[2 * k + 1] = h >> 8; }
          } else {
            const uint8_t* img = reinterpret_cast<const uint8_t*>(g == 0 ? p.x[0] : p.x[1]) + (int64_t)nl * p.Hi * p.Wi * 3;
            const int iy = iy0 + r;
#pragma unroll
            for (int k = 0; k < 12; ++k) {
              const int ix = ix0 + 4 * q + k / 3;
              b[k] = (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) ? (uint32_t)__ldg(img + ((int64_t)iy * p.Wi + ix) * 3 + k % 3) : 0u;
            }
          }
          uint4* dst = reinterpret_cast<uint4*>(cvt + (r * PP + 4 * q) * 8);
          dst[0] = make_uint4(u8x2_bf16(b[0], b[1]), u8x2_bf16(b[2], 0u), u8x2_bf16(b[3], b[4]), u8x2_bf16(b[5], 0u));
          dst[1] = make_uint4(u8x2_bf16(b[6], b[7]), u8x2_bf16(b[8], 0u), u8x2_bf16(b[9], b[10]), u8x2_bf16(b[11], 0u));
        }
      } else {
        // single uint8 plane (the depth image before cvtColor replicates it, utils/utils.py:14-19): 4 bytes -> 32 bytes,
        // each value written to the three channel slots -- identical to uploading the replicated image
        for (int item = ctid; item < PH * (PP / 4); item += kCvtWarps * 32) {
          const int r = item / (PP / 4), q = item - r * (PP / 4);
          uint32_t b[4];
          if (p.use_tma) {
            const uint16_t* src = reinterpret_cast<const uint16_t*>(raw + r * RAW1B + (U8_LEFT - 2) + 4 * q);
            const uint32_t h0 = src[0], h1 = src[1];
            b[0] = h0 & 0xffu; b[1] = h0 >> 8; b[2] = h1 & 0xffu; b[3] = h1 >> 8;
          } else {
            const uint8_t* img = reinterpret_cast<const uint8_t*>(p.x[1]) + (int64_t)nl * p.Hi * p.Wi;
            const int iy = iy0 + r;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int ix = ix0 + 4 * q + k;
              b[k] = (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) ? (uint32_t)__ldg(img + (int64_t)iy * p.Wi + ix) : 0u;
            }
          }
          uint4* dst = reinterpret_cast<uint4*>(cvt + (r * PP + 4 * q) * 8);
          const uint32_t p0 = u8x2_bf16(b[0], b[0]), p1 = u8x2_bf16(b[1], b[1]), p2 = u8x2_bf16(b[2], b[2]), p3 = u8x2_bf16(b[3], b[3]);
          dst[0] = make_uint4(p0, p0 & 0xffffu, p1, p1 & 0xffffu);
          dst[1] = make_uint4(p2, p2 & 0xffffu, p3, p3 & 0xffffu);
        }
      }
      ptx::fence_proxy_async_smem();   // this thread's st.shared -> visible to the async proxy (MMA operand reads)
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(bar_cvt_full + 8u * cs);
        if (p.use_tma) ptx::mbar_arrive(bar_raw_empty + 8u * rs);
      }
      if (++rs == NRAW) { rs = 0; rph ^= 1u; }
      if (++cs == NCVT) { cs = 0; cph ^= 1u; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: group = tile parity, warp & 3 = TMEM lane quarter
    const int q4 = warp & 3, grp = warp >> 2;
    const int mrow = q4 * 32 + lane;              // accumulator row (TMEM lane)
    const int ch = mrow & (p.C0pad - 1);          // C0pad is 32, 64 or 128
    const int rep = mrow / p.C0pad;               // replica: handles the units rep, rep + nrep, ...
    const int nrep = 128 / p.C0pad;
    const bool ch_valid = ch < p.C0;
    float sc[2], bi[2];
    sc[0] = __ldg(p.scale + ch); bi[0] = __ldg(p.bias + ch);
    sc[1] = p.groups > 1 ? __ldg(p.scale + p.C0pad + ch) : sc[0];
    bi[1] = p.groups > 1 ? __ldg(p.bias + p.C0pad + ch) : bi[0];
    uint32_t it = 0;
    for (; cur.n < p.n_img; cur.advance(p.tiles_x, p.tiles_y), ++it) {
      if ((int)(it & 1u) != grp) continue;
      const uint32_t aph = (it >> 1) & 1u;
      const int g = cur.n >= p.group_imgs ? 1 : 0;
      const float s = g ? sc[1] : sc[0], b = g ? bi[1] : bi[0];
      const int py0 = cur.ty * TPH, px0 = cur.tx * TPW;
      const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;   // image coordinates of conv pixel (0, 0) of the tile
      const bool border = cy0 < 0 || cx0 < 0 || cy0 + CH > p.Hi || cx0 + CW > p.Wi;
      __nv_bfloat16* ybase = p.y.p + p.y.img_off(cur.n) + ch;
      ptx::mbar_wait(bar_tm_full + 8u * grp, aph);
      ptx::tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)grp * 256u;
      for (int u = rep; u < TPW / 4; u += nrep) {
        // unit u: pooled columns 4u..4u+3 = conv columns 8u..8u+8: per conv row five even columns (accumulator 0,
        // TMEM columns cy*18 + 4u + 0..4) and four odd ones (accumulator 1, columns cy*18 + 4u + 0..3)
        uint32_t E[CH][8], O[CH][4];
#pragma unroll
        for (int cy = 0; cy < CH; ++cy) {
          tmem_ld_x8(t0 + (uint32_t)(cy * (PP / 2) + 4 * u), E[cy]);
          ptx::tmem_ld_x4(t0 + (uint32_t)(NCOL + cy * (PP / 2) + 4 * u), O[cy]);
        }
        ptx::tmem_ld_wait();
        if (u + nrep >= TPW / 4) {   // last unit of this thread: the accumulators may be overwritten by the next tile
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_tm_empty + 8u * grp);
        }
        float ve[5][CH], vo[4][CH];
#pragma unroll
        for (int cy = 0; cy < CH; ++cy) {
#pragma unroll
          for (int i = 0; i < 5; ++i) ve[i][cy] = __uint_as_float(E[cy][i]);
#pragma unroll
          for (int i = 0; i < 4; ++i) vo[i][cy] = __uint_as_float(O[cy][i]);
        }
        if (border) {   // exclude out-of-image conv positions from the max (reference: -inf pool padding)
#pragma unroll
          for (int cy = 0; cy < CH; ++cy) {
            const bool yok = cy0 + cy >= 0 && cy0 + cy < p.Hi;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              const int gx = cx0 + 8 * u + 2 * i;
              if (!(yok && gx >= 0 && gx < p.Wi)) ve[i][cy] = -INFINITY;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int gx = cx0 + 8 * u + 2 * i + 1;
              if (!(yok && gx >= 0 && gx < p.Wi)) vo[i][cy] = -INFINITY;
            }
          }
        }
        // vertical 3-max at stride 2 per conv column, then horizontal 3-max at stride 2
        float he[5][TPH], ho[4][TPH];
#pragma unroll
        for (int pr = 0; pr < TPH; ++pr) {
#pragma unroll
          for (int i = 0; i < 5; ++i) he[i][pr] = max3(ve[i][2 * pr], ve[i][2 * pr + 1], ve[i][2 * pr + 2]);
#pragma unroll
          for (int i = 0; i < 4; ++i) ho[i][pr] = max3(vo[i][2 * pr], vo[i][2 * pr + 1], vo[i][2 * pr + 2]);
        }
        if (ch_valid) {
#pragma unroll
          for (int pr = 0; pr < TPH; ++pr) {
            const int py = py0 + pr;
            if (py < p.Ho) {
              __nv_bfloat16* yrow = ybase + (int64_t)(py * p.Wo + px0 + 4 * u) * p.y.ld;
#pragma unroll
              for (int pc = 0; pc < 4; ++pc) {
                const float m = max3(he[pc][pr], ho[pc][pr], he[pc + 1][pr]);
                const float o = fmaxf(fmaf(m, s, b), 0.0f);
                if (px0 + 4 * u + pc < p.Wo) yrow[pc * p.y.ld] = __float2bfloat16_rn(o);
              }
            }
          }
        }
      }
    }
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
  DCFA_REQUIRE(op.w_gstride == 6 * 128 * 16 && op.sb_gstride == a.C0pad, "stem: weight tiles must be [G][3][2][128x16]");
  DCFA_REQUIRE(((uintptr_t)a.w % 16) == 0, "stem: weights must be 16-byte aligned");
  a.groups = a.n_img / a.group_imgs;
  DCFA_REQUIRE(!c1 || (u8 && a.groups == 2), "stem: the single-plane flag needs uint8 inputs and two groups");
  a.tiles_x = ceil_div(a.Wo, TPW);
  a.tiles_y = ceil_div(a.Ho, TPH);
  const int64_t total = (int64_t)a.n_img * a.tiles_x * a.tiles_y;
  DCFA_REQUIRE(total < (1ll << 31), "stem: too many tiles");

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
  const size_t smem = 1024 + 2 * A_GROUP_BYTES + NRAW * RAW_SLOT + NCVT * CVT_SLOT + 256;
  cudaError_t e = cudaFuncSetAttribute(stem_kernel<MODE_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_kernel<MODE_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_kernel<MODE_U8_C1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(DCFA_E_CUDA, "stem: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  int64_t grid = sm_count();   // one CTA per SM: the kernel owns all 512 TMEM columns
  if (grid > total) grid = total;
  if (!u8) launch_pdl(stem_kernel<MODE_F32>, dim3((unsigned)grid), dim3(kStemThreads), smem, st, maps[0], maps[1], a);
  else if (!c1) launch_pdl(stem_kernel<MODE_U8>, dim3((unsigned)grid), dim3(kStemThreads), smem, st, maps[0], maps[1], a);
  else launch_pdl(stem_kernel<MODE_U8_C1>, dim3((unsigned)grid), dim3(kStemThreads), smem, st, maps[0], maps[1], a);
  DCFA_CHECK_LAUNCH("stem_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
