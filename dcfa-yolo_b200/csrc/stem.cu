// DCFA_OP_STEM: Conv_maxpool (nets/yolo_mul.py:104-115) fused into one pass on tcgen05 tensor cores:
//   fp32 NCHW image -> conv3x3 s1 p1 (3 -> C0) -> BN scale/bias -> ReLU -> maxpool 3x3 s2 p1 -> bf16 NHWC.
// The full-resolution conv map (the largest tensor of the network) only ever exists as a shared-memory tile.
// Both modalities run in one launch (weight group = image / group_imgs).
//
// Persistent CTAs; per tile of 4 x 32 pooled pixels (= 9 x 65 conv pixels = 11 x 67 x 3 input patch):
//   1. the fp32 patch is staged in shared memory (zero outside the image = conv zero padding);
//   2. im2col: for 5 M-tiles of 128 conv pixels, threads build the K-major, 128B-swizzled A tile
//      (27 taps -> bf16, K padded to 32) in a double-buffered 16 KB buffer and one thread issues
//      2 x tcgen05.mma (M=128, N=BN, K=16) per M-tile into TMEM columns [t*BN, (t+1)*BN);
//   3. epilogue: tcgen05.ld, fp32 scale/bias, ReLU (0 outside the image), bf16 conv tile in shared memory;
//   4. 3x3/2 max-pool from shared memory, 128-bit stores of the pooled NHWC tile.
// Post-ReLU values are >= 0 and every pool window holds a valid pixel, so writing 0 for out-of-image conv
// positions is equivalent to the reference's -inf pool padding.
#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {
namespace {

constexpr int TPH = 4, TPW = 32;              // pooled tile
constexpr int CH = 2 * TPH + 1;               // 9 conv rows
constexpr int CW = 2 * TPW + 1;               // 65 conv cols
constexpr int NPIX = CH * CW;                 // 585 conv pixels
constexpr int MT = (NPIX + 127) / 128;        // 5 M-tiles
constexpr int PH = CH + 2, PW = CW + 2;       // 11 x 67 input patch
constexpr int PWP = 68;                       // padded patch row
constexpr int kStemThreads = 256;
constexpr int A_BYTES = 128 * 128;            // one M-tile of A: 128 rows x 128 B (64 bf16, 32 used)

struct StemArgs {
  const float* x0;
  const float* x1;
  const __nv_bfloat16* w;  // [G][BN*64] swizzled tile image (pack_conv_weight of [C0,3,3,3])
  const float* scale;      // [G][BN]
  const float* bias;       // [G][BN]
  View<__nv_bfloat16> y;
  int n_img, group_imgs, Hi, Wi, Ho, Wo, C0, BN;
  int tiles_x, tiles_y, tiles_per_group, total_tiles;
  uint32_t tmem_cols;
};

// 16 consecutive im2col entries k = HALF*16 .. HALF*16+15 of one conv pixel, k = (ky*3 + kx)*3 + ci (27..31 = 0)
template <int HALF>
__device__ __forceinline__ void gather16(const float* pin, float* v) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    constexpr int dummy = 0;
    (void)dummy;
    const int k = HALF * 16 + j;
    if (k < 27) {
      const int ci = k % 3, kk = k / 3;
      const int kx = kk % 3, ky = kk / 3;
      v[j] = pin[(ci * PH + ky) * PWP + kx];
    } else {
      v[j] = 0.0f;
    }
  }
}

__global__ void __launch_bounds__(kStemThreads) stem_kernel(const StemArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  // layout: A[2] (32 KB) | B (BN*128, <= 16 KB) | patch (3*11*68 fp32) | conv tile (NPIX*C0 bf16) | barriers
  const uint32_t s_a = base;
  const uint32_t s_b = s_a + 2 * A_BYTES;
  const uint32_t b_bytes = (uint32_t)p.BN * 128u;
  float* s_in = reinterpret_cast<float*>(gbase + 2 * A_BYTES + 16384);
  __nv_bfloat16* s_conv = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(s_in) + 3 * PH * PWP * 4);
  const uint32_t bars = s_b + 16384u + 3u * PH * PWP * 4u + (uint32_t)NPIX * p.C0 * 2u;
  const uint32_t bar_buf = (bars + 7u) & ~7u;          // 2 barriers: A buffer free
  const uint32_t bar_done = bar_buf + 16u;             // all MMAs of the tile complete
  const uint32_t tmem_slot = bar_done + 8u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_init(bar_buf, 1);
      ptx::mbar_init(bar_buf + 8u, 1);
      ptx::mbar_init(bar_done, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t idesc = ptx::make_idesc_bf16_f32(128, p.BN);

  int cur_group = -1;
  uint32_t buf_phase[2] = {0u, 0u};   // parity of the next completion to wait for, per A buffer
  uint32_t done_phase = 0u;
  uint32_t buf_uses[2] = {0u, 0u};

  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    // tiles are ordered group-major so that a CTA reloads its weights at most once
    const int g = tile / p.tiles_per_group;
    int t = tile - g * p.tiles_per_group;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int nl = t / p.tiles_y;
    const int n = g * p.group_imgs + nl;
    const float* img = (g == 0 ? p.x0 : p.x1) + (int64_t)nl * 3 * p.Hi * p.Wi;
    const int py0 = ty * TPH, px0 = tx * TPW;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;  // conv-map origin of the tile
    const int iy0 = cy0 - 1, ix0 = cx0 - 1;          // input origin of the patch

    if (g != cur_group) {  // (re)load the pre-swizzled weight tile of this modality
      const uint4* src = reinterpret_cast<const uint4*>(p.w + (int64_t)g * p.BN * 64);
      uint4* dst = reinterpret_cast<uint4*>(gbase + 2 * A_BYTES);
      for (int i = tid; i < (int)(b_bytes / 16); i += kStemThreads) dst[i] = __ldg(src + i);
      cur_group = g;
    }
    for (int i = tid; i < 3 * PH * PW; i += kStemThreads) {
      const int c = i / (PH * PW);
      const int r = (i - c * PH * PW) / PW;
      const int q = i - c * PH * PW - r * PW;
      const int iy = iy0 + r, ix = ix0 + q;
      float v = 0.0f;
      if (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) v = __ldg(img + ((int64_t)c * p.Hi + iy) * p.Wi + ix);
      s_in[(c * PH + r) * PWP + q] = v;
    }
    __syncthreads();

    // ---- im2col + MMA over the 5 M-tiles
    for (int mt = 0; mt < MT; ++mt) {
      const int b = mt & 1;
      if (buf_uses[b] > 0) {  // wait until the MMAs that read this buffer have completed
        ptx::mbar_wait(bar_buf + 8u * b, buf_phase[b]);
        buf_phase[b] ^= 1u;
      }
      buf_uses[b]++;
      // two threads per row: half 0 builds k = 0..15, half 1 builds k = 16..31 (27..31 are zero)
      const int r = tid >> 1, half = tid & 1;
      const int m = mt * 128 + r;
      uint32_t pk[8];
      if (m < NPIX) {
        const int cy = m / CW, cx = m - cy * CW;
        float v[16];
        const float* pin = s_in + cy * PWP + cx;
        if (half == 0) gather16<0>(pin, v); else gather16<1>(pin, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = 0u;
      }
      uint8_t* arow = gbase + b * A_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
      const int c0 = half * 2;  // 16-byte chunks c0, c0+1 of the 128-byte row, swizzled by (r & 7)
      *reinterpret_cast<uint4*>(arow + (((c0) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(arow + (((c0 + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      ptx::fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        ptx::tc_fence_after();
        const uint64_t adesc = ptx::make_sw128_kmajor_desc(s_a + b * A_BYTES);
        const uint64_t bdesc = ptx::make_sw128_kmajor_desc(s_b);
        const uint32_t d = tmem_base + (uint32_t)(mt * p.BN);
        ptx::umma_bf16(d, adesc, bdesc, idesc, 0u);
        ptx::umma_bf16(d, adesc + 2, bdesc + 2, idesc, 1u);
        ptx::umma_commit(bar_buf + 8u * b);
        if (mt == MT - 1) ptx::umma_commit(bar_done);
      }
    }

    // ---- epilogue: TMEM -> scale/bias/ReLU -> bf16 conv tile in shared memory
    ptx::mbar_wait(bar_done, done_phase);
    done_phase ^= 1u;
    ptx::tc_fence_after();
    {
      const float* sc = p.scale + (int64_t)g * p.BN;
      const float* bi = p.bias + (int64_t)g * p.BN;
      const int q4 = warp & 3;                       // TMEM lane quarter this warp may read
      for (int mt = (warp >> 2); mt < MT; mt += 2) { // warps 0-3: tiles 0,2,4; warps 4-7: tiles 1,3
        const int m = mt * 128 + q4 * 32 + lane;
        bool inside = false;
        if (m < NPIX) {
          const int cy = m / CW, cx = m - cy * CW;
          const int gy = cy0 + cy, gx = cx0 + cx;
          inside = gy >= 0 && gy < p.Hi && gx >= 0 && gx < p.Wi;
        }
        for (int j = 0; j < p.BN / 16; ++j) {
          uint32_t acc[16];
          ptx::tmem_ld_x16(tmem_base + (uint32_t)(mt * p.BN + j * 16) + ((uint32_t)(q4 * 32) << 16), acc);
          ptx::tmem_ld_wait();
          if (m < NPIX && j * 16 < p.C0) {
            float v[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + j * 16) + q);
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bi + j * 16) + q);
              v[4 * q + 0] = fmaxf(fmaf(__uint_as_float(acc[4 * q + 0]), s4.x, b4.x), 0.0f);
              v[4 * q + 1] = fmaxf(fmaf(__uint_as_float(acc[4 * q + 1]), s4.y, b4.y), 0.0f);
              v[4 * q + 2] = fmaxf(fmaf(__uint_as_float(acc[4 * q + 2]), s4.z, b4.z), 0.0f);
              v[4 * q + 3] = fmaxf(fmaf(__uint_as_float(acc[4 * q + 3]), s4.w, b4.w), 0.0f);
            }
            if (!inside) {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = 0.0f;
            }
            __nv_bfloat16* dst = s_conv + (int64_t)m * p.C0 + j * 16;
            *reinterpret_cast<uint4*>(dst) = pack8(v);
            if (j * 16 + 8 < p.C0) *reinterpret_cast<uint4*>(dst + 8) = pack8(v + 8);
          }
          __syncwarp();
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();

    // ---- 3x3 stride-2 max-pool of the conv tile, pooled NHWC store
    const int c8n = p.C0 >> 3;
    for (int i = tid; i < TPH * TPW * c8n; i += kStemThreads) {
      const int c8 = i % c8n;
      const int pp = i / c8n;
      const int pxl = pp % TPW, pyl = pp / TPW;
      const int py = py0 + pyl, px = px0 + pxl;
      if (py >= p.Ho || px >= p.Wo) continue;
      __nv_bfloat162 mx[4];
      {
        const uint4 v = *reinterpret_cast<const uint4*>(s_conv + (int64_t)((2 * pyl) * CW + 2 * pxl) * p.C0 + c8 * 8);
        mx[0] = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
        mx[1] = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
        mx[2] = *reinterpret_cast<const __nv_bfloat162*>(&v.z);
        mx[3] = *reinterpret_cast<const __nv_bfloat162*>(&v.w);
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          if (r == 0 && q == 0) continue;
          const uint4 v =
              *reinterpret_cast<const uint4*>(s_conv + (int64_t)((2 * pyl + r) * CW + 2 * pxl + q) * p.C0 + c8 * 8);
          mx[0] = __hmax2(mx[0], *reinterpret_cast<const __nv_bfloat162*>(&v.x));
          mx[1] = __hmax2(mx[1], *reinterpret_cast<const __nv_bfloat162*>(&v.y));
          mx[2] = __hmax2(mx[2], *reinterpret_cast<const __nv_bfloat162*>(&v.z));
          mx[3] = __hmax2(mx[3], *reinterpret_cast<const __nv_bfloat162*>(&v.w));
        }
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&mx[0]);
      o.y = *reinterpret_cast<uint32_t*>(&mx[1]);
      o.z = *reinterpret_cast<uint32_t*>(&mx[2]);
      o.w = *reinterpret_cast<uint32_t*>(&mx[3]);
      stg128(p.y.p + p.y.img_off(n) + (int64_t)(py * p.Wo + px) * p.y.ld + c8 * 8, o);
    }
    __syncthreads();  // conv tile and patch are reused by the next tile
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace

int launch_stem(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  StemArgs a;
  a.x0 = resolve_ptr<const float>(op.x, bufs);
  a.x1 = resolve_ptr<const float>(op.x2, bufs);
  a.w = resolve_ptr<const __nv_bfloat16>(op.w, bufs);
  a.scale = resolve_ptr<const float>(op.scale, bufs);
  a.bias = resolve_ptr<const float>(op.bias, bufs);
  a.y = resolve<__nv_bfloat16>(op.y, bufs);
  a.n_img = op.n_img;
  a.group_imgs = op.group_imgs > 0 ? op.group_imgs : op.n_img;
  a.Hi = op.Hi; a.Wi = op.Wi; a.Ho = op.Ho; a.Wo = op.Wo; a.C0 = op.Cout; a.BN = op.BN;
  DCFA_REQUIRE(a.x0 && a.w && a.scale && a.bias && a.y.p, "stem: missing tensor");
  DCFA_REQUIRE(a.n_img == a.group_imgs || (a.n_img == 2 * a.group_imgs && a.x1), "stem: needs 1 or 2 groups");
  DCFA_REQUIRE(a.Hi > 0 && a.Wi > 0 && a.Ho == (a.Hi - 1) / 2 + 1 && a.Wo == (a.Wi - 1) / 2 + 1,
               "stem: pooled size %dx%d inconsistent with %dx%d", a.Ho, a.Wo, a.Hi, a.Wi);
  DCFA_REQUIRE(a.C0 % 8 == 0 && a.C0 >= 8 && a.C0 <= 96, "stem: C0 %d unsupported", a.C0);
  DCFA_REQUIRE(a.BN % 16 == 0 && a.BN >= a.C0 && a.BN <= 96 && op.k_blocks == 1 && op.n_tiles == 1 && op.K_real == 27,
               "stem: weight packing mismatch (BN %d, C0 %d)", a.BN, a.C0);
  DCFA_REQUIRE(((uintptr_t)a.y.p % 16) == 0 && a.y.ld % 8 == 0 && a.y.img_stride % 8 == 0 && a.y.gstride % 8 == 0,
               "stem: output view must be 16-byte aligned");
  DCFA_REQUIRE(((uintptr_t)a.w % 16) == 0 && ((uintptr_t)a.scale % 16) == 0 && ((uintptr_t)a.bias % 16) == 0,
               "stem: parameters must be 16-byte aligned");
  a.tiles_x = ceil_div(a.Wo, TPW);
  a.tiles_y = ceil_div(a.Ho, TPH);
  const int64_t per_group = (int64_t)a.group_imgs * a.tiles_x * a.tiles_y;
  const int64_t total = per_group * (a.n_img / a.group_imgs);
  DCFA_REQUIRE(total < (1ll << 31), "stem: too many tiles");
  a.tiles_per_group = (int)per_group;
  a.total_tiles = (int)total;
  uint32_t cols = 32;
  while (cols < (uint32_t)(MT * a.BN)) cols <<= 1;
  a.tmem_cols = cols;
  const size_t smem = 1024 + 2 * A_BYTES + 16384 + (size_t)3 * PH * PWP * 4 + (size_t)NPIX * a.C0 * 2 + 64;
  static int max_set = 0;
  if ((int)smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(DCFA_E_CUDA, "stem: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    max_set = (int)smem;
  }
  // CTAs per SM: limited by shared memory and by TMEM columns (512 per SM)
  int per_sm = (int)((227 * 1024) / smem);
  if (per_sm > (int)(512 / cols)) per_sm = (int)(512 / cols);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > total) grid = total;
  stem_kernel<<<(unsigned)grid, kStemThreads, smem, st>>>(a);
  DCFA_CHECK_LAUNCH("stem_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
