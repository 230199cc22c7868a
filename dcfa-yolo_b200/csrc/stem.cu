// DCFA_OP_STEM: Conv_maxpool (nets/yolo_mul.py:104-115) fused into one pass on tcgen05 tensor cores:
//   fp32 NCHW image -> conv3x3 s1 p1 (3 -> C0) -> BN scale/bias -> ReLU -> maxpool 3x3 s2 p1 -> bf16 NHWC.
// The full-resolution conv map (the largest tensor of the network) never leaves the SM.
// Both modalities run in one launch (weight group = image / group_imgs).
//
// Round-1 profile of the first tensor-core version: 1 740 instructions per thread per tile, IPC 1.6 -- bound
// by the CUDA-core work around the MMA (im2col, per-conv-pixel epilogue, pooling through shared memory).
// This version turns the GEMM around:
//     D[channel, pixel] = W[channel, K=27] * im2col[pixel, K]^T        (M = 128 channel rows, N = 240 pixels)
//   * the accumulator of one CHANNEL is one TMEM lane and the conv PIXELS are its columns, so after
//     tcgen05.ld a thread holds a 7 x 5 window of one channel in registers and the 3x3/2 max-pool is pure
//     register arithmetic (no shared-memory round trip for the conv map);
//   * C0 < 128 channels are replicated over the 128 MMA rows, so every TMEM lane quarter (= every epilogue
//     warp) works, each replica pooling a different strip of the tile;
//   * BN + ReLU are applied AFTER pooling (4x fewer elements).  Exact: with scale >= 0, x -> fma(x, s, b) and ReLU
//     are monotone, so they commute with max; channels with negative scale are packed with negated weights and
//     |scale| (pack time), so the pooled quantity is always monotone-increasing in the accumulator;
//   * the fp32 input patch arrives by TMA (cp.async.bulk.tensor over the NCHW image): out-of-image pixels are
//     zero-filled by the hardware (= conv zero padding), the copy is asynchronous and double-buffered.
// Tile: 3 x 16 pooled pixels = 7 x 33 conv pixels = 9 x 35 x 3 input patch; 256 threads, 256 TMEM columns
// -> two CTAs per SM overlap each other's phases.  Out-of-image conv positions are excluded from the max
// (reference: -inf pool padding).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {
namespace {

#ifndef DCFA_STEM_TPW
#define DCFA_STEM_TPW 16
#endif
constexpr int TPH = 3, TPW = DCFA_STEM_TPW;   // pooled tile
constexpr int CH = 2 * TPH + 1;               // 7 conv rows
constexpr int CW = 2 * TPW + 1;               // 33 conv cols
constexpr int NPIX = CH * CW;                 // 231 conv pixels, GEMM column m = cx * CH + cy
constexpr int UMMA_N = (NPIX + 15) / 16 * 16;  // 128 (TPW 8) or 240 (TPW 16)
constexpr int PH = CH + 2;                    // 9 patch rows
#ifndef DCFA_STEM_PWB
#define DCFA_STEM_PWB (2 * DCFA_STEM_TPW + 8)
#endif
constexpr int PWB = DCFA_STEM_PWB;            // patch row pitch in floats: 35 needed + XOFF, 16-byte multiple for the TMA box
constexpr int XOFF = 2;                       // the TMA box must start on a 16-byte boundary of the innermost (x) dimension:
                                              // it starts at image column 2*px0 - 4, two columns left of the patch
constexpr int PATCH_BUF = (3 * (2 * TPH + 3) * PWB * 4 + 1023) / 1024 * 1024;   // bytes reserved per patch buffer
constexpr int NBUF = 4;                       // patch ring: loads run NBUF-1 tiles ahead (round 1: with one box in flight
                                              // per CTA the kernel was bound by the TMA round-trip latency)
constexpr int PATCH_FLOATS = 3 * PH * PWB;    // 972
constexpr int PATCH_BYTES = PATCH_FLOATS * 4; // 3888
constexpr int kStemThreads = TPW == 16 ? 256 : 128;
constexpr int kStemCtasPerSm = TPW == 16 ? 2 : 4;
constexpr int KROW = 64;                      // bytes per K row: 32 bf16 (27 used), SWIZZLE_64B
constexpr int B_BYTES = (TPW == 16 ? 256 : 128) * KROW;   // im2col tile
constexpr int A_BYTES = 128 * KROW;           // replicated weight tile
constexpr uint32_t kTmemCols = TPW == 16 ? 256 : 128;

// uint8 NHWC input (DCFA_STEM_FLAG_U8): the raw patch is 9 rows of RAWB bytes starting 16 pixels (48 bytes, a
// 16-byte multiple as the TMA box start requires) left of the tile's first conv column; bytes [32, 160) of each
// row are converted once per tile to bf16 (exact: 0..255 are bf16 integers; the 1/255 of preprocess_input,
// utils/utils.py:76-79, is folded into the BN scale) at U8_PITCH values per row.
constexpr int RAWB = (3 * (16 + CW + 2) + 15) / 16 * 16;   // 160 for TPW 16
constexpr int RAW_BYTES = RAWB * PH;                        // 1440
constexpr int U8_SKIP = 32;                                 // first converted byte of a row
constexpr int U8_CVT = RAWB - U8_SKIP;                      // 128 bytes converted per row
constexpr int U8_PITCH = U8_CVT + 8;                        // converted row pitch in bf16 values: 68 words, so that the
                                                            // 7 rows x 5 columns a warp touches spread over the banks
constexpr int U8_CVT_OFF = (RAW_BYTES + 127) / 128 * 128;   // converted patch offset inside the ring slot
static_assert(U8_CVT_OFF + PH * U8_PITCH * 2 <= PATCH_BUF, "u8 patch does not fit the ring slot");
static_assert(U8_CVT % 16 == 0 && U8_PITCH % 8 == 0 && RAWB <= 256, "u8 patch geometry");

struct StemArgs {
  const void* x[2];        // fp32 NCHW, or uint8 NHWC
  const __nv_bfloat16* w;  // [G][128*32] swizzled (SWIZZLE_64B) replicated weight tile
  const float* scale;      // [G][C0pad]  (>= 0, sign folded into the weights)
  const float* bias;       // [G][C0pad]
  View<__nv_bfloat16> y;
  int n_img, group_imgs, Hi, Wi, Ho, Wo, C0, C0pad;
  int tiles_x, tiles_y, tiles_per_group, total_tiles;
  int use_tma;
};

__device__ __forceinline__ void tma_load_patch_u8(uint32_t dst, const CUtensorMap* map, int xb, int y, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(xb), "r"(y), "r"(n), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void tma_load_patch(uint32_t dst, const CUtensorMap* map, int x, int y, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(0), "r"(n), "r"(bar)
      : "memory");
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

template <bool U8>
__global__ void __launch_bounds__(kStemThreads, kStemCtasPerSm) stem_kernel(const __grid_constant__ CUtensorMap map0,
                                                               const __grid_constant__ CUtensorMap map1,
                                                               const StemArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  // layout: B im2col (16 KB) | A weights (8 KB) | patch[2] (2 x 3888, padded to 4096) | staging (48 px * C0pad bf16) | barriers
  const uint32_t s_b = base;
  const uint32_t s_a = s_b + B_BYTES;
  const uint32_t s_patch = s_a + A_BYTES;
  float* patch_ptr = reinterpret_cast<float*>(gbase + B_BYTES + A_BYTES);
  __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(gbase + B_BYTES + A_BYTES + NBUF * PATCH_BUF);
  const uint32_t bars = s_patch + (uint32_t)(NBUF * PATCH_BUF) + (uint32_t)(TPH * TPW) * 128u * 2u;
  const uint32_t bar_patch = bars;            // NBUF barriers
  const uint32_t bar_mma = bars + 8u * NBUF;
  const uint32_t tmem_slot = bar_mma + 8u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < NBUF; ++i) ptx::mbar_init(bar_patch + 8u * i, 1);
      ptx::mbar_init(bar_mma, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t idesc = ptx::make_idesc_bf16_f32(128, UMMA_N);
  ptx::pdl_wait();

  // ---- static per-thread roles
  // im2col: GEMM column (row of the K-major B tile) m = cx*CH + cy; one thread builds columns 2q, 2q+1 of a conv row
  // epilogue: TMEM lane = 32*(warp%4) + lane -> MMA row; channel = row % C0pad; replica = row / C0pad
  const int q4 = warp & 3, half = warp >> 2;
  const int mrow = q4 * 32 + lane;
  const int ch = mrow % p.C0pad;
  const int rep = mrow / p.C0pad;
  const int nrep = 128 / p.C0pad;             // 4, 2 or 1 replicas
  const int nsub = nrep * (kStemThreads / 128); // strips: (replica, warp half)
  const int units_per_sub = (TPW / 2) / nsub; // units of 2 pooled columns per strip: 1, 2 or 4
  const int sub = rep * (kStemThreads / 128) + half;
  const bool ch_valid = ch < p.C0;

  // pooled-tile store mapping (tile-invariant): 16-byte chunk idx -> (pooled pixel pp, channel chunk c8)
  constexpr int kMaxChunks = (TPH * TPW * 16 + kStemThreads - 1) / kStemThreads;   // C0 <= 128 -> 16 chunks / pixel
  const int c8n = p.C0 >> 3;
  const int nchunks = TPH * TPW * c8n;
  int st_pp[kMaxChunks], st_c8[kMaxChunks];
#pragma unroll
  for (int k = 0; k < kMaxChunks; ++k) {
    const int i = tid + k * kStemThreads;
    st_pp[k] = i < nchunks ? i / c8n : -1;
    st_c8[k] = i < nchunks ? i - (i / c8n) * c8n : 0;
  }

  int cur_group = -1;
  float sc = 0.0f, bi = 0.0f;
  uint32_t it = 0;            // tiles processed by this CTA: ring slot = it % NBUF, its mbarrier parity = (it / NBUF) & 1
  uint32_t mma_phase = 0u;

  TileIter cur;               // the tile being processed (all threads)
  cur.init(blockIdx.x, gridDim.x, p.tiles_x, p.tiles_y);
  TileIter pre = cur;         // the tile whose patch is requested next (thread 32 only), NBUF-1 tiles ahead

  // issue the TMA load of one tile's patch into ring slot `slot` (one thread)
  auto issue_patch = [&](const TileIter& tc, int slot) {
    const uint32_t bar = bar_patch + 8u * slot;
    const uint32_t dst = s_patch + (uint32_t)slot * (uint32_t)PATCH_BUF;
    const int g = tc.n >= p.group_imgs ? 1 : 0;
    const int y = 2 * tc.ty * TPH - 2;
    // (no pointer select between the two maps: that would copy a __grid_constant__ parameter to local memory)
    if (U8) {
      const int xb = 3 * (2 * tc.tx * TPW - 16);
      ptx::mbar_arrive_expect_tx(bar, RAW_BYTES);
      if (g == 0) tma_load_patch_u8(dst, &map0, xb, y, tc.n, bar);
      else tma_load_patch_u8(dst, &map1, xb, y, tc.n - p.group_imgs, bar);
    } else {
      const int x = 2 * tc.tx * TPW - 2 - XOFF;
      ptx::mbar_arrive_expect_tx(bar, PATCH_BYTES);
      if (g == 0) tma_load_patch(dst, &map0, x, y, tc.n, bar);
      else tma_load_patch(dst, &map1, x, y, tc.n - p.group_imgs, bar);
    }
  };
  if (p.use_tma && tid == 32) {   // prologue: request the first NBUF-1 patches
    for (int d = 0; d < NBUF - 1; ++d) {
      if (pre.n < p.n_img) issue_patch(pre, d);
      pre.advance(p.tiles_x, p.tiles_y);
    }
  }

  int buf = 0;
  for (; cur.n < p.n_img; cur.advance(p.tiles_x, p.tiles_y), buf = (buf + 1 == NBUF ? 0 : buf + 1), ++it) {
    const int g = cur.n >= p.group_imgs ? 1 : 0;
    const int nl = cur.n - g * p.group_imgs;
    const int py0 = cur.ty * TPH, px0 = cur.tx * TPW;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;  // conv-map origin of the tile (patch origin is one less)
    float* s_in = patch_ptr + buf * (PATCH_BUF / 4);   // [3][PH][PWB]

    if (g != cur_group) {  // (re)load this modality's weight tile and this thread's scale/bias
      const uint4* src = reinterpret_cast<const uint4*>(p.w + (int64_t)g * 128 * 32);   // packing matches the input type
      uint4* dst = reinterpret_cast<uint4*>(gbase + B_BYTES);
      for (int i = tid; i < A_BYTES / 16; i += kStemThreads) dst[i] = __ldg(src + i);
      sc = __ldg(p.scale + (int64_t)g * p.C0pad + ch);
      bi = __ldg(p.bias + (int64_t)g * p.C0pad + ch);
      cur_group = g;
    }
    if (p.use_tma) {
      ptx::mbar_wait(bar_patch + 8u * buf, (it / NBUF) & 1u);
    } else if (U8) {   // image rows are not 16-byte multiples: plain loads with explicit zero padding
      const uint8_t* img = reinterpret_cast<const uint8_t*>(g == 0 ? p.x[0] : p.x[1]) + (int64_t)nl * p.Hi * p.Wi * 3;
      uint8_t* raw = reinterpret_cast<uint8_t*>(s_in);
      const int xb0 = 3 * (2 * px0 - 16), rowb = 3 * p.Wi;
      for (int i = tid; i < RAW_BYTES; i += kStemThreads) {
        const int r = i / RAWB, q = i - r * RAWB;
        const int iy = cy0 - 1 + r, xb = xb0 + q;
        raw[i] = (iy >= 0 && iy < p.Hi && xb >= 0 && xb < rowb) ? __ldg(img + (int64_t)iy * rowb + xb) : (uint8_t)0;
      }
      __syncthreads();
    } else {   // image rows are not 16-byte multiples: plain loads with explicit zero padding
      const float* img = reinterpret_cast<const float*>(g == 0 ? p.x[0] : p.x[1]) + (int64_t)nl * 3 * p.Hi * p.Wi;
      constexpr int PC = CW + 2;   // patch columns actually used
      for (int i = tid; i < 3 * PH * PC; i += kStemThreads) {
        const int c = i / (PH * PC);
        const int r = (i - c * PH * PC) / PC;
        const int q = i - c * PH * PC - r * PC;
        const int iy = cy0 - 1 + r, ix = cx0 - 1 + q;
        float v = 0.0f;
        if (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) v = __ldg(img + ((int64_t)c * p.Hi + iy) * p.Wi + ix);
        s_in[(c * PH + r) * PWB + q + XOFF] = v;
      }
      __syncthreads();
    }

    if (U8) {
      // ---- uint8 -> bf16, once per patch byte (each byte feeds up to 9 taps of ~2 conv pixels)
      constexpr int CHUNKS = PH * (U8_CVT / 16);
      if (tid < CHUNKS) {
        const int r = tid / (U8_CVT / 16), c = tid - r * (U8_CVT / 16);
        const uint4 b = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(s_in) + r * RAWB + U8_SKIP + c * 16);
        const uint32_t w[4] = {b.x, b.y, b.z, b.w};
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          o[2 * k] = pack_bf16x2((float)(w[k] & 0xffu), (float)((w[k] >> 8) & 0xffu));
          o[2 * k + 1] = pack_bf16x2((float)((w[k] >> 16) & 0xffu), (float)(w[k] >> 24));
        }
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(s_in) + U8_CVT_OFF + (r * U8_PITCH + c * 16) * 2);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
      __syncthreads();
      // ---- im2col: K index = ky*10 + kx*3 + ci (slots 9, 19, 29 carry a neighbouring value under a zero weight;
      //      30, 31 are zero).  The 9 taps of one kernel row are 9 consecutive bf16 of the converted patch row,
      //      starting at value 10 + 3*cx.  One thread builds columns 2q and 2q+1 of conv row cy from seven words per
      //      kernel row: the even column's five words as they are, the odd column's shifted by a word and a half.
      if (tid < (CW + 1) / 2 * CH) {
        const int q = tid / CH, cy = tid - q * CH;
        const int cx = 2 * q;
        const uint32_t* prow = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(s_in) + U8_CVT_OFF) +
                               cy * (U8_PITCH / 2) + 5 + 3 * q;
        uint32_t w[3][7];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int k = 0; k < 7; ++k) w[ky][k] = prow[ky * (U8_PITCH / 2) + k];
#pragma unroll
        for (int px = 0; px < 2; ++px) {
          if (px == 1 && cx + 1 >= CW) break;
          uint32_t pk[16];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int k = 0; k < 5; ++k) pk[ky * 5 + k] = px == 0 ? w[ky][k] : __funnelshift_r(w[ky][k + 1], w[ky][k + 2], 16u);
          pk[15] = 0u;
          const int m = (cx + px) * CH + cy;
          const uint32_t rowb = s_b + (uint32_t)m * KROW;
          const uint32_t xr = (uint32_t)((m >> 1) & 3);   // SWIZZLE_64B: chunk ^= (row >> 1) & 3
#pragma unroll
          for (int c = 0; c < 4; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + ((((uint32_t)c) ^ xr) << 4)), "r"(pk[4 * c]),
                         "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                         : "memory");
        }
      }
    } else if (tid < (CW + 1) / 2 * CH) {
      // ---- im2col: two horizontally adjacent conv pixels per thread (columns 2q, 2q+1 of conv row cy): the 4 floats
      //      of a (channel, kernel row) feeding both pixels are two aligned 64-bit loads instead of 2 x 3 scalar ones.
      //      K index = (ky*3 + kx)*3 + ci, 27 taps + 5 zeros -> 4 x 16 bytes per pixel.
      const int q = tid / CH, cy = tid - q * CH;
      const int cx = 2 * q;
      const float* pin = s_in + cy * PWB + cx + XOFF;   // even float index: 8-byte aligned
      float f[3][3][4];
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float2 a = *reinterpret_cast<const float2*>(pin + (ci * PH + ky) * PWB);
          const float2 b = *reinterpret_cast<const float2*>(pin + (ci * PH + ky) * PWB + 2);
          f[ci][ky][0] = a.x; f[ci][ky][1] = a.y; f[ci][ky][2] = b.x; f[ci][ky][3] = b.y;
        }
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        if (px == 1 && cx + 1 >= CW) break;
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float v0 = 0.0f, v1 = 0.0f;
          const int k0 = 2 * j, k1 = 2 * j + 1;   // compile-time after unrolling
          if (k0 < 27) v0 = f[k0 % 3][k0 / 9][(k0 / 3) % 3 + px];
          if (k1 < 27) v1 = f[k1 % 3][k1 / 9][(k1 / 3) % 3 + px];
          pk[j] = pack_bf16x2(v0, v1);
        }
        const int m = (cx + px) * CH + cy;
        const uint32_t rowb = s_b + (uint32_t)m * KROW;
        const uint32_t xr = (uint32_t)((m >> 1) & 3);   // SWIZZLE_64B: chunk ^= (row >> 1) & 3
#pragma unroll
        for (int c = 0; c < 4; ++c)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + ((((uint32_t)c) ^ xr) << 4)), "r"(pk[4 * c]),
                       "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                       : "memory");
      }
    }
    __syncthreads();   // B tile (and A tile) written; the patch slot `buf` has been consumed

    if (tid == 32 && p.use_tma) {
      // request the patch NBUF-1 tiles ahead into the slot the previous tile released.  Issued by a different
      // thread than the MMA issuer: fence.proxy.async waits for the executing thread's own outstanding bulk
      // copies, which would serialise the prefetch ring behind every MMA.
      if (pre.n < p.n_img) issue_patch(pre, buf == 0 ? NBUF - 1 : buf - 1);
      pre.advance(p.tiles_x, p.tiles_y);
    }
    if (tid == 0) {
      // consumer-side proxy fence: the other threads' st.shared (ordered before this point by the barrier)
      // become visible to the async proxy that reads the MMA operands
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_after();
      // SWIZZLE_64B K-major descriptors: SBO = 8 rows * 64 B, layout code 4
      const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
      const uint64_t adesc = desc_hi | (uint64_t)((s_a & 0x3FFFFu) >> 4);
      const uint64_t bdesc = desc_hi | (uint64_t)((s_b & 0x3FFFFu) >> 4);
      ptx::umma_bf16(tmem_base, adesc, bdesc, idesc, 0u);
      ptx::umma_bf16(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);   // K 16..31: +32 bytes
      ptx::umma_commit(bar_mma);
    }
    ptx::mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1u;
    ptx::tc_fence_after();

    // ---- epilogue: per unit of 2 pooled columns, 5 conv columns x 7 conv rows = 35 consecutive TMEM columns
    const bool border = cy0 < 0 || cx0 < 0 || cy0 + CH > p.Hi || cx0 + CW > p.Wi;
    for (int u = 0; u < units_per_sub; ++u) {
      const int pc0 = 2 * (sub * units_per_sub + u);    // first pooled column of the unit (tile-local)
      const int j0 = 2 * pc0;                           // first conv column
      uint32_t raw[36];
      const uint32_t taddr = tmem_base + (uint32_t)(j0 * CH) + ((uint32_t)(q4 * 32) << 16);
      ptx::tmem_ld_x16(taddr, raw);
      ptx::tmem_ld_x16(taddr + 16u, raw + 16);
      ptx::tmem_ld_x4(taddr + 32u, raw + 32);           // columns 32..34 are needed; stay inside the allocation
      ptx::tmem_ld_wait();
      float v[5][CH];
#pragma unroll
      for (int jj = 0; jj < 5; ++jj)
#pragma unroll
        for (int i = 0; i < CH; ++i) v[jj][i] = __uint_as_float(raw[jj * CH + i]);
      if (border) {   // exclude out-of-image conv positions from the max (reference: -inf pool padding)
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) {
          const int gx = cx0 + j0 + jj;
          const bool xok = gx >= 0 && gx < p.Wi;
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            const int gy = cy0 + i;
            if (!(xok && gy >= 0 && gy < p.Hi)) v[jj][i] = -INFINITY;
          }
        }
      }
      // vertical 3-max at stride 2 for each conv column, then horizontal 3-max at stride 2
      float vm[5][TPH];
#pragma unroll
      for (int jj = 0; jj < 5; ++jj)
#pragma unroll
        for (int pr = 0; pr < TPH; ++pr) vm[jj][pr] = fmaxf(fmaxf(v[jj][2 * pr], v[jj][2 * pr + 1]), v[jj][2 * pr + 2]);
      if (ch_valid) {
#pragma unroll
        for (int pr = 0; pr < TPH; ++pr)
#pragma unroll
          for (int pc = 0; pc < 2; ++pc) {
            const float m = fmaxf(fmaxf(vm[2 * pc][pr], vm[2 * pc + 1][pr]), vm[2 * pc + 2][pr]);
            const float o = fmaxf(fmaf(m, sc, bi), 0.0f);
            stage[(pr * TPW + pc0 + pc) * p.C0pad + ch] = __float2bfloat16_rn(o);
          }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();   // staging complete; TMEM and the B tile may be overwritten by the next tile

    // ---- pooled NHWC tile -> global, 16 bytes per thread, channel-contiguous
    __nv_bfloat16* ybase = p.y.p + p.y.img_off(cur.n);
#pragma unroll
    for (int k = 0; k < kMaxChunks; ++k) {
      const int pp = st_pp[k];
      if (pp >= 0) {
        const int pyl = pp / TPW, pxl = pp - pyl * TPW;   // TPW is a compile-time power of two
        const int py = py0 + pyl, px = px0 + pxl;
        if (py < p.Ho && px < p.Wo) {
          const uint4 val = *reinterpret_cast<const uint4*>(stage + pp * p.C0pad + st_c8[k] * 8);
          stg128(ybase + (int64_t)(py * p.Wo + px) * p.y.ld + st_c8[k] * 8, val);
        }
      }
    }
    __syncthreads();   // staging is rewritten by the next tile's epilogue
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
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
  DCFA_REQUIRE(a.C0 % 8 == 0 && a.C0 >= 8 && a.C0 <= 128, "stem: C0 %d unsupported", a.C0);
  DCFA_REQUIRE((a.C0pad == 32 || a.C0pad == 64 || a.C0pad == 128) && a.C0pad >= a.C0 && op.K_real == 27,
               "stem: weight packing mismatch (C0pad %d, C0 %d)", a.C0pad, a.C0);
  DCFA_REQUIRE(((uintptr_t)a.y.p % 16) == 0 && a.y.ld % 8 == 0 && a.y.img_stride % 8 == 0 && a.y.gstride % 8 == 0,
               "stem: output view must be 16-byte aligned");
  DCFA_REQUIRE(((uintptr_t)a.w % 16) == 0, "stem: weights must be 16-byte aligned");
  a.tiles_x = ceil_div(a.Wo, TPW);
  a.tiles_y = ceil_div(a.Ho, TPH);
  const int64_t per_group = (int64_t)a.group_imgs * a.tiles_x * a.tiles_y;
  const int groups = a.n_img / a.group_imgs;
  const int64_t total = per_group * groups;
  DCFA_REQUIRE(total < (1ll << 31), "stem: too many tiles");
  a.tiles_per_group = (int)per_group;
  a.total_tiles = (int)total;

  // ---- tensor maps, zero fill outside the image:
  //      fp32 NCHW : dims (W, H, C, N), box (PWB, 9, 3, 1)          uint8 NHWC : dims (3W bytes, H, N), box (RAWB, 9, 1)
  alignas(64) CUtensorMap maps[2];
  memset(maps, 0, sizeof(maps));
  const bool ptr_ok = ((uintptr_t)a.x[0] % 16) == 0 && (groups == 1 || ((uintptr_t)a.x[1] % 16) == 0);
  a.use_tma = (ptr_ok && (u8 ? (3 * a.Wi) % 16 == 0 : a.Wi % 4 == 0)) ? 1 : 0;
  if (a.use_tma) {
    EncodeTiledFn enc = stem_encode_fn();
    DCFA_REQUIRE(enc != nullptr, "stem: cuTensorMapEncodeTiled entry point unavailable");
    for (int g = 0; g < groups; ++g) {
      const cuuint32_t es[4] = {1u, 1u, 1u, 1u};
      CUresult cr;
      if (u8) {
        const cuuint64_t gdim[3] = {(cuuint64_t)a.Wi * 3, (cuuint64_t)a.Hi, (cuuint64_t)a.group_imgs};
        const cuuint64_t gstr[2] = {(cuuint64_t)a.Wi * 3, (cuuint64_t)a.Wi * a.Hi * 3};
        const cuuint32_t box[3] = {(cuuint32_t)RAWB, (cuuint32_t)PH, 1u};
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
  const size_t smem = 1024 + B_BYTES + A_BYTES + NBUF * PATCH_BUF + (size_t)TPH * TPW * 128 * 2 + 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(stem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(stem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(DCFA_E_CUDA, "stem: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int64_t grid = (int64_t)sm_count() * kStemCtasPerSm;   // CTAs per SM bounded by TMEM columns (512 / kTmemCols)
  if (grid > total) grid = total;
  if (u8) launch_pdl(stem_kernel<true>, dim3((unsigned)grid), dim3(kStemThreads), smem, st, maps[0], maps[1], a);
  else launch_pdl(stem_kernel<false>, dim3((unsigned)grid), dim3(kStemThreads), smem, st, maps[0], maps[1], a);
  DCFA_CHECK_LAUNCH("stem_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
