// DCFA_OP_CONV: implicit-GEMM convolution (1x1 / 3x3, stride 1 / 2, pad k/2) on tcgen05 tensor cores.
//
// Replaces every dense nn.Conv2d + BatchNorm2d + activation of the reference hot path
// (nets/yolo_mul.py:190-204 Conv, :138-151 ShuffleNetV2 pointwise convs, :388-391 head;
//  nets/repghost.py:80-84 RepGhostModule.primary_conv, :291-305 Conv).
//
// GEMM view:  D[M, N] = A[M, K] * W[N, K]^T,  M = images*Ho*Wo (per weight group), N = Cout, K = ks*ks*Cin.
//   A is never materialised: producer warps gather 16-byte channel runs of the NHWC bf16 input straight
//   into the 128B-swizzled K-major shared-memory tile the UMMA descriptor expects (zero fill for padding,
//   K tail and M tail).  W tiles are pre-swizzled at pack time and arrive with one cp.async.bulk per stage.
//   One elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a double-buffered TMEM accumulator;
//   four epilogue warps read it back with tcgen05.ld and apply scale/bias (folded BN), activation,
//   optional post-scale and residual, then store bf16 NHWC at a channel offset (concat-free) or fp32 NCHW.
//
// Persistent CTAs (grid = min(tiles, SMs)), 9 warps:
//   warps 0-3 epilogue (TMEM lane quarter = warp index), warp 4 MMA issue + TMEM alloc, warps 5-8 gather.
//
// The gather is the critical loop (ncu, round 1: producer warps were issue-latency bound at ~2400 cycles per
// k-block).  Per tile each producer thread now precomputes, for its 8 rows, a base pointer and a 9-bit mask of
// the taps that fall inside the image; per k-block a row costs one mask test, one 64-bit add and the cp.async.
// 1x1 stride-1 convs over pixel-linear tensors skip the (n, y, x) decomposition entirely; all divisions use
// host-computed magic numbers.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {

namespace {

constexpr int BM = 128;           // UMMA M
constexpr int BK = 64;            // K elements per pipeline stage (one 128-byte swizzle row)
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kProdWarp0 = 5;
constexpr int kProdWarps = 4;
constexpr int kThreads = 32 * (kEpiWarps + 1 + kProdWarps);

constexpr int kMaxStages = 8;
constexpr int kRowsPerProdThread = BM / (kProdWarps * 32 / 8);  // 8

// n / d for 0 <= n < 2^31 with a host-computed magic multiplier (d == 1 -> mul == 0)
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

struct ConvArgs {
  View<const __nv_bfloat16> x;
  View<const __nv_bfloat16> res;
  View<void> y;
  const __nv_bfloat16* w;
  const float* scale;
  const float* bias;
  int64_t w_gstride;
  int64_t sb_gstride;
  int n_img, group_imgs, n_groups;
  int Hi, Wi, Cin, Ho, Wo, Cout, ksize, stride, pad;
  int BN, n_tiles, k_blocks, K_real;
  int act, out_mode, out_ctot, out_coff;
  float post_scale;
  int Mg;                 // GEMM rows per group
  int m_tiles;            // M tiles per group
  int total_tiles;
  int stages;

  int x_linear;           // 1x1 stride-1 conv over a pixel-linear input: row m of group g is pixel g*Mg + m
  int y_linear;           // bf16 output (and residual) are pixel-linear
  FastDiv div_howo, div_wo, div_cin;
  uint32_t tmem_cols;
  int dbg;
};

__global__ void __launch_bounds__(kThreads, 1) conv_gemm_kernel(const ConvArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  const uint32_t b_stage_bytes = (uint32_t)p.BN * 128u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_a + (uint32_t)S * A_STAGE_BYTES;
  const uint32_t bars = smem_b + (uint32_t)S * b_stage_bytes;
  const uint32_t bar_full = bars;                       // S barriers
  const uint32_t bar_empty = bars + 8u * kMaxStages;    // S barriers
  const uint32_t bar_tfull = bars + 16u * kMaxStages;   // 2 barriers
  const uint32_t bar_tempty = bar_tfull + 16u;          // 2 barriers
  const uint32_t tmem_slot = bar_tempty + 16u;          // 4 bytes
  // generic pointer to the TMEM-address slot
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        ptx::mbar_init(bar_full + 8u * s, kProdWarps * 32 + 1);  // 128 cp.async completions + 1 expect_tx arrival
        ptx::mbar_init(bar_empty + 8u * s, 1);              // one tcgen05.commit
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(bar_tfull + 8u * a, 1);              // one tcgen05.commit
        ptx::mbar_init(bar_tempty + 8u * a, kEpiWarps);     // one arrival per epilogue warp
      }
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

  const int HoWo = p.Ho * p.Wo;

  if (warp >= kProdWarp0) {
    // ------------------------------------------------------------------ producers (gather A, bulk-load W)
    const int ptid = threadIdx.x - kProdWarp0 * 32;  // 0..127
    const int chunk = ptid & 7;                      // 16-byte chunk (8 channels) inside the 128-byte K row
    const int row_base = ptid >> 3;                  // 0..15 ; rows row_base + 16*i
    const uint32_t dst_thread =
        (uint32_t)(row_base >> 3) * 1024u + (uint32_t)(row_base & 7) * 128u + (uint32_t)((chunk ^ (row_base & 7)) << 4);
    const int ntaps = p.ksize * p.ksize;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int rest = tile / p.n_tiles;
      const int mt = rest % p.m_tiles;
      const int g = rest / p.m_tiles;
      const __nv_bfloat16* rowptr[kRowsPerProdThread];  // address of tap (0,0), channel 0 of the row's window
      uint32_t tapmask[kRowsPerProdThread];             // bit t: tap t of this row lies inside the image
#pragma unroll
      for (int i = 0; i < kRowsPerProdThread; ++i) {
        const int m = mt * BM + row_base + 16 * i;
        rowptr[i] = p.x.p;
        tapmask[i] = 0u;
        if (m < p.Mg) {
          if (p.x_linear) {
            rowptr[i] = p.x.p + ((int64_t)g * p.Mg + m) * p.x.ld;
            tapmask[i] = 1u;
          } else {
            const int nl = (int)p.div_howo.div((uint32_t)m);
            const int rem = m - nl * HoWo;
            const int oy = (int)p.div_wo.div((uint32_t)rem);
            const int ox = rem - oy * p.Wo;
            const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
            rowptr[i] = p.x.p + p.x.img_off(g * p.group_imgs + nl) + ((int64_t)iy0 * p.Wi + ix0) * p.x.ld;
            uint32_t mk = 0u;
            for (int t = 0; t < ntaps; ++t) {
              const int dy = (t * 11) >> 5, dx = t - 3 * dy;  // t / 3, t % 3 for t < 9 (ksize 1: t == 0)
              const int iy = iy0 + dy, ix = ix0 + dx;
              if (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) mk |= 1u << t;
            }
            tapmask[i] = mk;
          }
        }
      }
      const __nv_bfloat16* wtile = p.w + (int64_t)g * p.w_gstride + (int64_t)nt * p.k_blocks * (p.BN * BK);
      for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
        const uint32_t s = it % (uint32_t)S;
        const uint32_t ph = (it / (uint32_t)S) & 1u;
        ptx::mbar_wait(bar_empty + 8u * s, ph ^ 1u);
        if (ptid == 0) {
          ptx::mbar_arrive_expect_tx(bar_full + 8u * s, b_stage_bytes);
          ptx::bulk_g2s(smem_b + s * b_stage_bytes, wtile + (int64_t)kb * (p.BN * BK), b_stage_bytes,
                        bar_full + 8u * s);
        }
        const int k = kb * BK + chunk * 8;
        int tap = 0, tap_off = 0;
        uint32_t kbit = 0u;
        if (k < p.K_real) {
          tap = (int)p.div_cin.div((uint32_t)k);
          const int ch = k - tap * p.Cin;
          const int dy = (tap * 11) >> 5, dx = tap - 3 * dy;
          tap_off = (dy * p.Wi + dx) * p.x.ld + ch;
          kbit = 1u << tap;
        }
        const uint32_t dst = smem_a + s * A_STAGE_BYTES + dst_thread;
#pragma unroll
        for (int i = 0; i < kRowsPerProdThread; ++i) {
          const bool ok = (tapmask[i] & kbit) != 0u;
          ptx::cp_async_16(dst + (uint32_t)i * 2048u, ok ? rowptr[i] + tap_off : p.x.p, ok ? 16u : 0u);
        }
        // arrives on the stage's full barrier when all cp.asyncs of this thread have landed; the producer
        // never blocks on its own loads, so up to `stages` k-blocks are in flight per SM
        ptx::cp_async_mbar_arrive_noinc(bar_full + 8u * s);
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = ptx::make_idesc_bf16_f32(BM, p.BN);
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t as = tcount & 1u;
      const uint32_t aph = (tcount >> 1) & 1u;
      ptx::mbar_wait(bar_tempty + 8u * as, aph ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * (uint32_t)p.BN;
      for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
        const uint32_t s = it % (uint32_t)S;
        const uint32_t ph = (it / (uint32_t)S) & 1u;
        ptx::mbar_wait(bar_full + 8u * s, ph);
        if (lane == 0) {
          ptx::fence_proxy_async_smem();  // cp.async (generic proxy) writes -> tcgen05.mma (async proxy) reads
          ptx::tc_fence_after();
          const uint64_t adesc = ptx::make_sw128_kmajor_desc(smem_a + s * A_STAGE_BYTES);
          const uint64_t bdesc = ptx::make_sw128_kmajor_desc(smem_b + s * b_stage_bytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
            ptx::umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(bar_empty + 8u * s);
          if (kb == p.k_blocks - 1) ptx::umma_commit(bar_tfull + 8u * as);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    uint32_t tcount = 0;
    const int r = warp * 32 + lane;  // accumulator row == TMEM lane
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const int nt = tile % p.n_tiles;
      const int rest = tile / p.n_tiles;
      const int mt = rest % p.m_tiles;
      const int g = rest / p.m_tiles;
      const uint32_t as = tcount & 1u;
      const uint32_t aph = (tcount >> 1) & 1u;

      const int m = mt * BM + r;
      const bool rvalid = m < p.Mg;
      const float* sc = p.scale + (int64_t)g * p.sb_gstride + nt * p.BN;
      const float* bi = p.bias + (int64_t)g * p.sb_gstride + nt * p.BN;
      __nv_bfloat16* yb = nullptr;
      float* yf = nullptr;
      const __nv_bfloat16* rb = nullptr;
      if (rvalid) {
        if (p.out_mode == DCFA_OUT_BF16_NHWC && p.y_linear) {
          const int64_t gm = (int64_t)g * p.Mg + m;
          yb = reinterpret_cast<__nv_bfloat16*>(p.y.p) + gm * p.y.ld;
          if (p.res.p) rb = p.res.p + gm * p.res.ld;
        } else {
          const int nl = (int)p.div_howo.div((uint32_t)m);
          const int pix = m - nl * HoWo;
          const int n = g * p.group_imgs + nl;
          if (p.out_mode == DCFA_OUT_BF16_NHWC) {
            yb = reinterpret_cast<__nv_bfloat16*>(p.y.p) + p.y.img_off(n) + (int64_t)pix * p.y.ld;
            if (p.res.p) rb = p.res.p + p.res.img_off(n) + (int64_t)pix * p.res.ld;
          } else {
            yf = reinterpret_cast<float*>(p.y.p) + (int64_t)n * p.y.img_stride + (int64_t)p.out_coff * HoWo + pix;
          }
        }
      }

      ptx::mbar_wait(bar_tfull + 8u * as, aph);
      ptx::tc_fence_after();
      const uint32_t taddr0 = tmem_base + as * (uint32_t)p.BN + ((uint32_t)(warp * 32) << 16);
      for (int j = 0; j < p.BN / 16; ++j) {
        uint32_t acc[16];
        ptx::tmem_ld_x16(taddr0 + (uint32_t)(j * 16), acc);
        ptx::tmem_ld_wait();
        const int c0 = nt * p.BN + j * 16;  // first output channel of this chunk
        if (rvalid && c0 < p.Cout) {
          float v[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + j * 16) + q);
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bi + j * 16) + q);
            v[4 * q + 0] = fmaf(__uint_as_float(acc[4 * q + 0]), s4.x, b4.x);
            v[4 * q + 1] = fmaf(__uint_as_float(acc[4 * q + 1]), s4.y, b4.y);
            v[4 * q + 2] = fmaf(__uint_as_float(acc[4 * q + 2]), s4.z, b4.z);
            v[4 * q + 3] = fmaf(__uint_as_float(acc[4 * q + 3]), s4.w, b4.w);
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = apply_act(v[e], p.act) * p.post_scale;
          if (p.out_mode == DCFA_OUT_BF16_NHWC) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (c0 + 8 * h + 8 <= p.Cout) {
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
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (c0 + e < p.Cout) yf[(int64_t)(c0 + e) * HoWo] = v[e];
          }
        }
        __syncwarp();  // reconverge before the next warp-collective tcgen05.ld
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + 8u * as);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

template <typename T>
bool pixel_linear(const View<T>& v, int64_t hw) {
  return v.img_stride == hw * v.ld && (v.gi <= 0 || v.gstride == (int64_t)v.gi * v.img_stride);
}

}  // namespace

int launch_conv(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  if ((op.flags & 0xff) != 0) return launch_conv_tma(op, bufs, st);  // packed per (tap, channel block): TMA path
  DCFA_REQUIRE(op.parts == 0, "conv(gather): split outputs need the TMA path");
  DCFA_REQUIRE(!(op.flags & DCFA_CONV_FLAG_DFL), "conv(gather): the fused DFL epilogue needs the TMA path");
  ConvArgs a;
  a.x = resolve<const __nv_bfloat16>(op.x, bufs);
  a.res = resolve<const __nv_bfloat16>(op.x2, bufs);
  a.y = resolve<void>(op.y, bufs);
  a.w = resolve_ptr<const __nv_bfloat16>(op.w, bufs);
  a.scale = resolve_ptr<const float>(op.scale, bufs);
  a.bias = resolve_ptr<const float>(op.bias, bufs);
  a.w_gstride = op.w_gstride;
  a.sb_gstride = op.sb_gstride;
  a.n_img = op.n_img;
  a.group_imgs = op.group_imgs > 0 ? op.group_imgs : op.n_img;
  a.Hi = op.Hi; a.Wi = op.Wi; a.Cin = op.Cin;
  a.Ho = op.Ho; a.Wo = op.Wo; a.Cout = op.Cout;
  a.ksize = op.ksize; a.stride = op.stride; a.pad = op.ksize / 2;
  a.BN = op.BN; a.n_tiles = op.n_tiles; a.k_blocks = op.k_blocks; a.K_real = op.K_real;
  a.act = op.act; a.out_mode = op.out_mode; a.out_ctot = op.out_ctot; a.out_coff = op.out_coff;
  a.post_scale = op.f0;

  DCFA_REQUIRE(a.x.p && a.y.p && a.w && a.scale && a.bias, "conv: missing tensor");
  DCFA_REQUIRE(a.n_img > 0 && a.n_img % a.group_imgs == 0, "conv: n_img %d not a multiple of group_imgs %d",
               a.n_img, a.group_imgs);
  DCFA_REQUIRE(a.ksize == 1 || a.ksize == 3, "conv: ksize %d unsupported", a.ksize);
  DCFA_REQUIRE(a.stride == 1 || a.stride == 2, "conv: stride %d unsupported", a.stride);
  DCFA_REQUIRE(a.Cin > 0 && a.Cin % 8 == 0, "conv: Cin %d must be a positive multiple of 8", a.Cin);
  DCFA_REQUIRE(a.Ho == (a.Hi + 2 * a.pad - a.ksize) / a.stride + 1 && a.Wo == (a.Wi + 2 * a.pad - a.ksize) / a.stride + 1,
               "conv: output size %dx%d inconsistent with input %dx%d k%d s%d", a.Ho, a.Wo, a.Hi, a.Wi, a.ksize, a.stride);
  DCFA_REQUIRE(a.BN >= 16 && a.BN <= 256 && a.BN % 16 == 0, "conv: BN %d must be a multiple of 16 in [16,256]", a.BN);
  DCFA_REQUIRE(a.n_tiles >= 1 && a.n_tiles * a.BN >= a.Cout, "conv: n_tiles*BN < Cout");
  DCFA_REQUIRE(a.K_real == a.ksize * a.ksize * a.Cin, "conv: K_real mismatch");
  DCFA_REQUIRE(a.k_blocks == (a.K_real + BK - 1) / BK, "conv: k_blocks mismatch");
  DCFA_REQUIRE(((uintptr_t)a.x.p % 16) == 0 && a.x.ld % 8 == 0 && a.x.img_stride % 8 == 0 && a.x.gstride % 8 == 0,
               "conv: input view must be 16-byte aligned");
  DCFA_REQUIRE((int64_t)a.Hi * a.Wi * a.x.ld < (1ll << 31), "conv: image too large for 32-bit tap offsets");
  DCFA_REQUIRE(((uintptr_t)a.w % 16) == 0 && a.w_gstride % 8 == 0, "conv: weights must be 16-byte aligned");
  DCFA_REQUIRE(((uintptr_t)a.scale % 16) == 0 && ((uintptr_t)a.bias % 16) == 0 && a.sb_gstride % 4 == 0,
               "conv: scale/bias must be 16-byte aligned");
  if (a.out_mode == DCFA_OUT_BF16_NHWC) {
    DCFA_REQUIRE(a.Cout % 8 == 0, "conv: bf16 output needs Cout %% 8 == 0 (got %d)", a.Cout);
    DCFA_REQUIRE(((uintptr_t)a.y.p % 16) == 0 && a.y.ld % 8 == 0 && a.y.img_stride % 8 == 0 && a.y.gstride % 8 == 0,
                 "conv: output view must be 16-byte aligned");
    if (a.res.p)
      DCFA_REQUIRE(((uintptr_t)a.res.p % 16) == 0 && a.res.ld % 8 == 0 && a.res.img_stride % 8 == 0 && a.res.gstride % 8 == 0,
                   "conv: residual view must be 16-byte aligned");
  } else {
    DCFA_REQUIRE(a.out_mode == DCFA_OUT_F32_NCHW, "conv: bad out_mode %d", a.out_mode);
    DCFA_REQUIRE(!a.res.p, "conv: residual unsupported with fp32 NCHW output");
    DCFA_REQUIRE(a.out_coff >= 0 && a.out_coff + a.Cout <= a.out_ctot, "conv: NCHW channel slot out of range");
    DCFA_REQUIRE(a.y.img_stride == (int64_t)a.out_ctot * a.Ho * a.Wo, "conv: NCHW img_stride mismatch");
  }

  a.n_groups = a.n_img / a.group_imgs;
  const int64_t Mg = (int64_t)a.group_imgs * a.Ho * a.Wo;
  DCFA_REQUIRE(Mg < (1ll << 31) - BM, "conv: GEMM M too large");
  a.Mg = (int)Mg;
  a.m_tiles = (a.Mg + BM - 1) / BM;
  const int64_t total = (int64_t)a.n_groups * a.m_tiles * a.n_tiles;
  DCFA_REQUIRE(total < (1ll << 31), "conv: too many tiles");
  a.total_tiles = (int)total;
  a.div_howo = make_fastdiv((uint32_t)(a.Ho * a.Wo));
  a.div_wo = make_fastdiv((uint32_t)a.Wo);
  a.div_cin = make_fastdiv((uint32_t)a.Cin);
  // group g's images are [g*group_imgs, (g+1)*group_imgs): with a pixel-linear tensor, GEMM row m of group g
  // is simply pixel g*Mg + m (valid for 1x1 stride-1 inputs and for any output)
  a.x_linear = (a.ksize == 1 && a.stride == 1 && pixel_linear(a.x, (int64_t)a.Hi * a.Wi)) ? 1 : 0;
  a.y_linear = (a.out_mode == DCFA_OUT_BF16_NHWC && pixel_linear(a.y, (int64_t)a.Ho * a.Wo) &&
                (!a.res.p || pixel_linear(a.res, (int64_t)a.Ho * a.Wo)))
                   ? 1
                   : 0;

  const int stage_bytes = A_STAGE_BYTES + a.BN * 128;
  const int max_smem = 227 * 1024;
  const int fixed = 1024 /*alignment slack*/ + 256 /*barriers*/;
  int stages = (max_smem - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  DCFA_REQUIRE(stages >= 3, "conv: not enough shared memory for the pipeline");
  a.stages = stages;

  const int smem = fixed + stages * stage_bytes;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * a.BN)) cols <<= 1;
  a.tmem_cols = cols;
  a.dbg = 0;

  static DeviceOnce attr_set;
  if (attr_set.needed()) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return fail(DCFA_E_CUDA, "conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set.mark();
  }
  int grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
  launch_k(conv_gemm_kernel, dim3(grid), dim3(kThreads), smem, st, 0, false, a);
  DCFA_CHECK_LAUNCH("conv_gemm_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
