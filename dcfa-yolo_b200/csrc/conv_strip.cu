// DCFA_OP_CONV, halo-strip path: 3x3 stride-1 convolutions (the head's cv2 / cv3 branches, nets/yolo_mul.py:387-391) whose nine
// taps are served from ONE shared-memory copy of the input rows instead of nine TMA boxes.
//
// conv_tma_kernel fetches, per output tile, one 128-row box per (tap, channel block): the same input rows travel from L2
// to the SM nine times, and that A-operand delivery -- not the tensor pipe, not the weight tiles (profiles/README.md,
// round 2, items 3, 11, 14) -- sets the k-block rate of every 3x3 layer.  Here an M tile is 128 CONSECUTIVE positions of
// the image flattened with a pitch of W + 2 (one zero column left and right, produced by the TMA's out-of-bounds fill).
// With the rows of the padded image lying back to back in shared memory ("strip": pixel rows of 128 bytes = 64
// channels, SWIZZLE_128B), tap (dy, dx) of such a tile is again 128 consecutive pixel rows of the strip, dy * (W + 2) + dx
// further down: nine UMMA descriptors that differ only in their start row (tools/umma_rowoffset_test.cu: a SWIZZLE_128B
// descriptor may start any number of rows into a tile).  Per tile and 64-channel block the SM receives
// NR * (W + 2) <= 410 pixel rows instead of 9 * 128 = 1 152; the two pad columns cost 2 / (W + 2) of the MMA rows
// (their outputs are skipped by the epilogue).
//
// Roles (11 warps): warps 0-3 / 6-9 epilogue groups (as conv_tma_kernel's FAST instance: SiLU, bf16 NHWC, 256-bit stores),
// warp 4 weight producer (one bulk copy per tap and channel block into a ring), warp 5 MMA issuer + TMEM, warp 10 strip
// producer (one TMA box per tile and channel block into a ring of two strips).  Weights are the tensor-map path's own
// packed tiles ([n_tile][k_block = tap * cblocks + cb][BN x 64]); K is summed channel block by channel block, tap by tap.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {
namespace {

constexpr int kStripThreads = 352;
constexpr int kWProducerWarp = 4, kMmaWarpS = 5, kStripWarp = 10;
constexpr int kStripUnits = 2;
constexpr int kMaxWStages = 12;

struct FastDivS {
  uint32_t d, mul, shr;
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return mul ? (__umulhi(n, mul) >> shr) : n; }
};
FastDivS make_fastdiv_s(uint32_t d) {
  FastDivS f{d, 0u, 0u};
  if (d > 1) {
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    const uint32_t p = 31 + l;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
  }
  return f;
}

#ifdef DCFA_TIMELINE
// debug build only: clock64 timestamps of CTA 0 (rows: 0 strip issue, 1 strip landed (seen by MMA), 2 W issue, 3 W landed, 4 MMA
// k-block issued, 5 tile start (MMA), 6 epilogue tfull / done pairs), read back by tools/tl_strip.py
__device__ long long g_tls[8][4096];
#define TLS(role, idx) do { if (blockIdx.x == 0 && (idx) < 4096) g_tls[role][idx] = clock64(); } while (0)
#else
#define TLS(role, idx) do { } while (0)
#endif

struct StripArgs {
  View<__nv_bfloat16> y;
  const __nv_bfloat16* w;
  const float* scale;
  const float* bias;
  int64_t w_gstride, sb_gstride;
  int n_img, group_imgs;
  int H, W, P;            // P = W + 2: pitch of the flattened padded image
  int Cout, BN, n_tiles, cblocks;
  int NR;                 // padded rows per strip box
  int MT;                 // M tiles (128 positions each) per pass: every weight k-block feeds MT accumulators
  int tiles_img, total_tiles;
  int wstages;
  uint32_t strip_bytes, strip_tx, w_stage_bytes, w_tx_bytes;   // strip_bytes: ring pitch (multiple of 1024); strip_tx: bytes one box delivers
  uint32_t tmem_cols, acc_stages;
  int epi_split;
  FastDivS div_P, div_tiles_img, div_ntiles;
};

__device__ __forceinline__ void strip_tma_load(uint32_t dst, const CUtensorMap* map, int c, int x, int y, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c), "r"(x), "r"(y), "r"(n), "r"(bar)
      : "memory");
}

// MT = M tiles per pass, compile time: the MMA-issuing thread is a serial instruction stream, a runtime loop there costs more than the MMAs
template <int MT>
__global__ void __launch_bounds__(kStripThreads, 1) conv_strip_kernel(const __grid_constant__ CUtensorMap tmap, const StripArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_strip = base;                                              // [2] strips
  const uint32_t s_w = s_strip + kStripUnits * p.strip_bytes;                  // [wstages] weight tiles
  const uint32_t bars = s_w + (uint32_t)p.wstages * p.w_stage_bytes;
  const uint32_t bar_sfull = bars, bar_sempty = bars + 16u;
  const uint32_t bar_wfull = bars + 32u, bar_wempty = bar_wfull + 8u * kMaxWStages;
  const uint32_t bar_tfull = bar_wempty + 8u * kMaxWStages, bar_tempty = bar_tfull + 32u;
  const uint32_t tmem_slot = bar_tempty + 32u;
  const uint32_t sb_base = (tmem_slot + 4u + 15u) & ~15u;                      // 2 groups x (256 scale + 256 bias) floats
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == kMmaWarpS) {
    if (lane == 0) {
      for (int u = 0; u < kStripUnits; ++u) { ptx::mbar_init(bar_sfull + 8u * u, 1); ptx::mbar_init(bar_sempty + 8u * u, 1); }
      for (int s = 0; s < p.wstages; ++s) { ptx::mbar_init(bar_wfull + 8u * s, 1); ptx::mbar_init(bar_wempty + 8u * s, 1); }
      for (int a = 0; a < (int)p.acc_stages; ++a) {
        ptx::mbar_init(bar_tfull + 8u * a, 1);
        ptx::mbar_init(bar_tempty + 8u * a, p.epi_split ? 8 : 4);
      }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (warp == kStripWarp && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
  ptx::pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  ptx::pdl_wait();

  // tile -> (n-tile, image, first flattened position); strip geometry of a tile
  auto decode = [&](int tile, int& nt, int& n, int& q0) {
    const uint32_t rest = p.div_ntiles.div((uint32_t)tile);
    nt = tile - (int)rest * p.n_tiles;
    n = (int)p.div_tiles_img.div(rest);
    q0 = ((int)rest - n * p.tiles_img) * 128 * MT;
  };

  if (warp == kStripWarp) {
    // ------------------------------------------------------------------ strip producer
    if (ptx::elect_one()) {
      uint32_t u = 0, ph = 0;
      [[maybe_unused]] uint32_t tl = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nt, n, q0;
        decode(tile, nt, n, q0);
        const int r0 = q0 == 0 ? -1 : (int)p.div_P.div((uint32_t)(q0 - 1));   // first padded row of the strip (padded row r = image row r - 1)
        for (int cb = 0; cb < p.cblocks; ++cb) {
          ptx::mbar_wait(bar_sempty + 8u * u, ph ^ 1u);
          TLS(0, tl); ++tl;
          ptx::mbar_arrive_expect_tx(bar_sfull + 8u * u, p.strip_tx);
          strip_tma_load(s_strip + u * p.strip_bytes, &tmap, cb * 64, -1, r0 - 1, n, bar_sfull + 8u * u);
          if (++u == kStripUnits) { u = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == kWProducerWarp) {
    // ------------------------------------------------------------------ weight producer
    if (ptx::elect_one()) {
      uint32_t s = 0, ph = 0;
      [[maybe_unused]] uint32_t tl = 0;
      const int64_t wstep = (int64_t)p.BN * 64;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nt, n, q0;
        decode(tile, nt, n, q0);
        const int g = n / p.group_imgs;
        const __nv_bfloat16* wt = p.w + (int64_t)g * p.w_gstride + (int64_t)nt * (9 * p.cblocks) * wstep;
        for (int cb = 0; cb < p.cblocks; ++cb)
          for (int tap = 0; tap < 9; ++tap) {
            ptx::mbar_wait(bar_wempty + 8u * s, ph ^ 1u);
            TLS(2, tl); ++tl;
            ptx::mbar_arrive_expect_tx(bar_wfull + 8u * s, p.w_tx_bytes);
            ptx::bulk_g2s(s_w + s * p.w_stage_bytes, wt + (int64_t)(tap * p.cblocks + cb) * wstep, p.w_tx_bytes, bar_wfull + 8u * s);
            if (++s == (uint32_t)p.wstages) { s = 0; ph ^= 1u; }
          }
      }
    }
  } else if (warp == kMmaWarpS) {
    // ------------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16_f32(128, p.BN);
      // K-major SWIZZLE_128B descriptor: LBO field 1, SBO = 1024 bytes (8 rows), version bit 46, layout type 2
      const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      uint32_t u = 0, uph = 0, s = 0, sph = 0, as = 0, aph = 0;
      [[maybe_unused]] uint32_t tl_u = 0, tl_w = 0, tl_t = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nt, n, q0;
        decode(tile, nt, n, q0);
        const int r0 = q0 == 0 ? -1 : (int)p.div_P.div((uint32_t)(q0 - 1));
        const uint32_t row0 = (uint32_t)(q0 - 1 - r0 * p.P);   // strip row of tap (0, 0) of the tile's first position
        ptx::mbar_wait(bar_tempty + 8u * as, aph ^ 1u);
        TLS(5, tl_t); ++tl_t;
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * (uint32_t)(MT * p.BN);
        for (int cb = 0; cb < p.cblocks; ++cb) {
          ptx::mbar_wait(bar_sfull + 8u * u, uph);
          TLS(1, tl_u); ++tl_u;
          ptx::tc_fence_after();
          const uint32_t strip = s_strip + u * p.strip_bytes;
          const uint32_t a_row0 = strip + row0 * 128u;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {   // unrolled: dy, dx are constants, the tap offset is one add
            const int dy = tap / 3, dx = tap - dy * 3;
            ptx::mbar_wait(bar_wfull + 8u * s, sph);
            TLS(3, tl_w);
            ptx::tc_fence_after();
            const uint32_t a_addr = a_row0 + (uint32_t)(dy * p.P + dx) * 128u;
            const uint32_t b_addr = s_w + s * p.w_stage_bytes;
            const uint64_t adesc = desc_hi | (uint64_t)((a_addr & 0x3FFFFu) >> 4);
            const uint64_t bdesc = desc_hi | (uint64_t)((b_addr & 0x3FFFFu) >> 4);
#pragma unroll
            for (int m = 0; m < MT; ++m) {   // +128 strip rows = +1024 in the (address >> 4) field
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_bf16(d_tmem + (uint32_t)(m * p.BN), adesc + (uint64_t)(1024 * m + 2 * k), bdesc + (uint64_t)(2 * k), idesc,
                               (cb | tap | k) ? 1u : 0u);
            }
            ptx::umma_commit(bar_wempty + 8u * s);
            TLS(4, tl_w); ++tl_w;
            if (++s == (uint32_t)p.wstages) { s = 0; sph ^= 1u; }
          }
          ptx::umma_commit(bar_sempty + 8u * u);
          if (++u == kStripUnits) { u = 0; uph ^= 1u; }
        }
        ptx::umma_commit(bar_tfull + 8u * as);
        if (++as == p.acc_stages) { as = 0u; aph ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: group 0 = warps 0-3, group 1 = warps 6-9
    const int group = warp < 4 ? 0 : 1;
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;       // accumulator row = position q0 + r of the flattened padded image
    const int gtid = r;
    float* sb = reinterpret_cast<float*>(smem_raw + (sb_base - ptx::smem_u32(smem_raw))) + group * 512;
    int sb_key = -1;
    const uint32_t acc_shift = p.acc_stages == 4u ? 2u : (p.acc_stages == 2u ? 1u : 0u);
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      if (!p.epi_split && (int)(tcount & 1u) != group) continue;
      const uint32_t as = tcount & (p.acc_stages - 1u);
      const uint32_t aph = (tcount >> acc_shift) & 1u;
      int nt, n, q0;
      decode(tile, nt, n, q0);
      const int g = n / p.group_imgs;
      if (g * p.n_tiles + nt != sb_key) {
        sb_key = g * p.n_tiles + nt;
        ptx::named_bar_sync(1 + group, 128);
        const float* sc = p.scale + (int64_t)g * p.sb_gstride + nt * p.BN;
        const float* bi = p.bias + (int64_t)g * p.sb_gstride + nt * p.BN;
        // SiLU: the epilogue needs h = (acc * s + b) / 2 -- stage the halved vectors (exact)
        for (int c = gtid; c < p.BN; c += 128) { sb[c] = 0.5f * __ldg(sc + c); sb[256 + c] = 0.5f * __ldg(bi + c); }
        ptx::named_bar_sync(1 + group, 128);
      }
      const int cvalid = min(p.BN, p.Cout - nt * p.BN);
      ptx::mbar_wait(bar_tfull + 8u * as, aph);
      if (group == 0 && q4 == 0 && lane == 0) TLS(6, 2 * tcount);
      ptx::tc_fence_after();
      const int nchunks = p.BN >> 4;
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
      const int q = q0 + m * 128 + r;
      const int yy = (int)p.div_P.div((uint32_t)q);
      const int xp = q - yy * p.P;
      const bool rvalid = xp >= 1 && xp <= p.W && yy < p.H;   // the two pad columns and the tail of the last tile are not outputs
      __nv_bfloat16* yb = nullptr;
      if (rvalid) yb = p.y.p + p.y.img_off(n) + (int64_t)(yy * p.W + xp - 1) * p.y.ld + nt * p.BN;
      const uint32_t taddr0 = tmem_base + as * (uint32_t)(MT * p.BN) + (uint32_t)(m * p.BN) + ((uint32_t)(q4 * 32) << 16);
      uint32_t accA[16], accB[16];
      auto compute = [&](const uint32_t (&a)[16], const int j, float (&v)[16]) {
        const int c0 = j * 16;
        F2 v2[8];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const float4 s4 = *reinterpret_cast<const float4*>(sb + c0 + 4 * qq);
          const float4 b4 = *reinterpret_cast<const float4*>(sb + 256 + c0 + 4 * qq);
          v2[2 * qq] = f2_make(b4.x, b4.y);
          v2[2 * qq + 1] = f2_make(b4.z, b4.w);
          f2_fma(v2[2 * qq], f2_make(__uint_as_float(a[4 * qq + 0]), __uint_as_float(a[4 * qq + 1])), f2_make(s4.x, s4.y));
          f2_fma(v2[2 * qq + 1], f2_make(__uint_as_float(a[4 * qq + 2]), __uint_as_float(a[4 * qq + 3])), f2_make(s4.z, s4.w));
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {   // x * sigmoid(x) = h + h * tanh(h), h = x / 2
          F2 h = v2[e];
          float h0, h1, t0, t1;
          f2_get(h, h0, h1);
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
          f2_fma(h, h, f2_make(t0, t1));
          f2_get(h, v[2 * e], v[2 * e + 1]);
        }
      };
      auto emit = [&](float (&v)[16], const int j) {
        const int c0 = j * 16;
        if (rvalid && c0 < cvalid) {
          const uint4 lo = pack8(v), hi = pack8(v + 8);
          asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yb + c0), "r"(lo.x), "r"(lo.y), "r"(lo.z),
                       "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                       : "memory");
        }
      };
      const int jstep = p.epi_split ? 2 : 1, j0 = p.epi_split ? group : 0;
      ptx::tmem_ld_x16(taddr0 + (uint32_t)(j0 * 16), accA);
      if (j0 + jstep < nchunks) ptx::tmem_ld_x16(taddr0 + (uint32_t)((j0 + jstep) * 16), accB);
      for (int j = j0; j < nchunks; j += 2 * jstep) {
        const bool two = j + jstep < nchunks;
        float vA[16], vB[16];
        ptx::tmem_ld_wait();
        compute(accA, j, vA);
        if (two) compute(accB, j + jstep, vB);
        if (j + 2 * jstep < nchunks) ptx::tmem_ld_x16(taddr0 + (uint32_t)((j + 2 * jstep) * 16), accA);
        if (j + 3 * jstep < nchunks) ptx::tmem_ld_x16(taddr0 + (uint32_t)((j + 3 * jstep) * 16), accB);
        emit(vA, j);
        __syncwarp();
        if (two) {
          emit(vB, j + jstep);
          __syncwarp();
        }
      }
      }   // m
      if (group == 0 && q4 == 0 && lane == 0) TLS(6, 2 * tcount + 1);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + 8u * as);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarpS) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn strip_encode_fn() {
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

// Takes the op (returns its status with *taken = true) when it is a 3x3 stride-1 SiLU convolution with 64-channel
// k-blocks and a plain bf16 NHWC output; otherwise leaves it to conv_tma_kernel (*taken = false).
int launch_conv_strip(const dcfa_op& op, void* const* bufs, cudaStream_t st, bool* taken) {
  *taken = false;
  {
    const char* e = getenv("DCFA_CONV_STRIP");   // debug / tests: DCFA_CONV_STRIP=0 keeps every conv on the tap-box path
    if (e && atoi(e) == 0) return DCFA_OK;
  }
  if (op.ksize != 3 || op.stride != 1 || (op.flags & 0xff) != 64 || (op.flags & (DCFA_CONV_FLAG_PAIR | DCFA_CONV_FLAG_DFL)) ||
      op.act != DCFA_ACT_SILU || op.out_mode != DCFA_OUT_BF16_NHWC || op.parts != 0 || op.f0 != 1.0f || op.Cin % 64 != 0 ||
      op.Wi + 2 > 256 || op.Hi != op.Ho || op.Wi != op.Wo)
    return DCFA_OK;
  View<const __nv_bfloat16> x = resolve<const __nv_bfloat16>(op.x, bufs);
  View<const __nv_bfloat16> res = resolve<const __nv_bfloat16>(op.x2, bufs);
  StripArgs a;
  a.y = resolve<__nv_bfloat16>(op.y, bufs);
  a.w = resolve_ptr<const __nv_bfloat16>(op.w, bufs);
  a.scale = resolve_ptr<const float>(op.scale, bufs);
  a.bias = resolve_ptr<const float>(op.bias, bufs);
  if (res.p || !x.p || !a.y.p || !a.w || !a.scale || !a.bias) return DCFA_OK;
  if (x.gi > 0 && x.gstride != (int64_t)x.gi * x.img_stride) return DCFA_OK;
  const bool aligned = ((uintptr_t)x.p % 16) == 0 && x.ld % 8 == 0 && x.img_stride % 8 == 0 && ((uintptr_t)a.y.p % 32) == 0 &&
                       a.y.ld % 16 == 0 && a.y.img_stride % 16 == 0 && a.y.gstride % 16 == 0 && op.Cout % 16 == 0 &&
                       ((uintptr_t)a.w % 16) == 0 && ((uintptr_t)a.scale % 16) == 0 && ((uintptr_t)a.bias % 16) == 0;
  if (!aligned) return DCFA_OK;
  a.w_gstride = op.w_gstride; a.sb_gstride = op.sb_gstride;
  a.n_img = op.n_img;
  a.group_imgs = op.group_imgs > 0 ? op.group_imgs : op.n_img;
  a.H = op.Hi; a.W = op.Wi; a.P = op.Wi + 2;
  a.Cout = op.Cout; a.BN = op.BN; a.n_tiles = op.n_tiles;
  a.cblocks = op.Cin / 64;
  if (a.BN < 16 || a.BN > 256 || a.BN % 16 != 0 || op.k_blocks != 9 * a.cblocks || a.n_img % a.group_imgs != 0) return DCFA_OK;
  // rows of the padded image a strip must hold: positions q0 - 1 ... q0 + 128 + 2P, starting anywhere inside a row
  a.w_tx_bytes = (uint32_t)a.BN * 128u;
  a.w_stage_bytes = (a.w_tx_bytes + 1023u) & ~1023u;
  const int max_smem = 227 * 1024;
  const int fixed = 1024 + 512 + 4096 + 1024;
  // Two M tiles per pass when their accumulators fit TMEM and the bigger strips still leave a weight ring: with the A side
  // served from the strip, the weight tiles are what an SM pulls from L2 (9 * cblocks * BN * 128 bytes per pass), and
  // every k-block then feeds twice the MMAs.
  // Measured (B200, s, B=32): two tiles per pass pay for BN <= 128 (two accumulator stages remain: head0.cls1 0.080 -> 0.065 ms,
  // head0.box1 0.044 -> 0.035) and lose for BN = 192 (one stage: the epilogue is exposed, head0.0 0.086 -> 0.094) and when
  // the halved number of passes no longer fills the SMs (20 x 20 maps: head2.0 0.033 -> 0.047).  Packing head*.0 as two
  // n-tiles of 96 columns so that it could take two tiles per pass was measured too: head0.0 0.082 (unchanged), head1.0 0.049 -> 0.057.
  const int64_t passes2 = (int64_t)a.n_img * a.n_tiles * (((int64_t)a.H * a.P + 255) / 256);
  int mt_max = (4 * a.BN <= 512 && passes2 >= sm_count()) ? 2 : 1;
  { const char* e = getenv("DCFA_STRIP_MT"); if (e && (atoi(e) == 1 || atoi(e) == 2)) mt_max = 2 * a.BN <= 512 ? atoi(e) : 1; }   // debug
  a.MT = 0;
  for (int mt = mt_max; mt >= 1; --mt) {
    const int nr = (a.P - 1 + 128 * mt + 2 + 2 * a.P) / a.P + 1;
    const uint32_t tx = (uint32_t)nr * a.P * 128u;
    const uint32_t pitch = (tx + 1023u) & ~1023u;
    int wst = (max_smem - fixed - kStripUnits * (int)pitch) / (int)a.w_stage_bytes;
    if (wst > kMaxWStages) wst = kMaxWStages;
    if (nr <= 256 && wst >= 3) {
      a.MT = mt; a.NR = nr; a.strip_tx = tx; a.strip_bytes = pitch; a.wstages = wst;
      break;
    }
  }
  if (a.MT == 0) return DCFA_OK;   // not enough shared memory for a useful weight ring: leave it to the tap-box kernel
  a.tiles_img = (a.H * a.P + 128 * a.MT - 1) / (128 * a.MT);
  const int64_t total = (int64_t)a.n_img * a.tiles_img * a.n_tiles;
  if (total >= (1ll << 31)) return DCFA_OK;
  a.total_tiles = (int)total;
  const int pass_cols = a.MT * a.BN;
  a.acc_stages = 4 * pass_cols <= 512 ? 4u : (2 * pass_cols <= 512 ? 2u : 1u);
  a.epi_split = pass_cols >= 128 ? 1 : 0;   // both groups share every pass (the alternate mode needs >= 2 accumulator stages)
  uint32_t cols = 32;
  while (cols < a.acc_stages * (uint32_t)pass_cols) cols <<= 1;
  a.tmem_cols = cols;
  a.div_P = make_fastdiv_s((uint32_t)a.P);
  a.div_tiles_img = make_fastdiv_s((uint32_t)a.tiles_img);
  a.div_ntiles = make_fastdiv_s((uint32_t)a.n_tiles);

  EncodeTiledFn enc = strip_encode_fn();
  if (!enc) return DCFA_OK;
  *taken = true;
  alignas(64) CUtensorMap tmap;
  const cuuint64_t gdim[4] = {(cuuint64_t)op.Cin, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.n_img};
  const cuuint64_t gstr[3] = {(cuuint64_t)x.ld * 2, (cuuint64_t)a.W * x.ld * 2, (cuuint64_t)x.img_stride * 2};
  const cuuint32_t box[4] = {64u, (cuuint32_t)a.P, (cuuint32_t)a.NR, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const int64_t run = (op.Cin == x.ld) ? (int64_t)op.Cin * 2 * a.W : (int64_t)op.Cin * 2;
  const CUtensorMapL2promotion promo = run >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : (run >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B);
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x.p), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(DCFA_E_CUDA, "conv(strip): cuTensorMapEncodeTiled failed with %d", (int)cr);
  static DeviceOnce attr_set;
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(conv_strip_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_strip_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return fail(DCFA_E_CUDA, "conv(strip): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.mark();
  }
  const int smem = fixed + kStripUnits * (int)a.strip_bytes + a.wstages * (int)a.w_stage_bytes;
  const int grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
  if (a.MT == 2) launch_pdl(conv_strip_kernel<2>, dim3(grid), dim3(kStripThreads), smem, st, tmap, a);
  else launch_pdl(conv_strip_kernel<1>, dim3(grid), dim3(kStripThreads), smem, st, tmap, a);
  DCFA_CHECK_LAUNCH("conv_strip_kernel");
  return DCFA_OK;
}

}  // namespace dcfa

#ifdef DCFA_TIMELINE
extern "C" int dcfa_debug_read_strip_timeline(void* dst, int bytes) {
  return cudaMemcpyFromSymbol(dst, dcfa::g_tls, bytes) == cudaSuccess ? 0 : -2;
}
extern "C" int dcfa_debug_clear_strip_timeline() {
  static long long zeros[8][4096];
  return cudaMemcpyToSymbol(dcfa::g_tls, zeros, sizeof(zeros)) == cudaSuccess ? 0 : -2;
}
#endif
