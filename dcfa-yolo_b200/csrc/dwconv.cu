// DCFA_OP_DWCONV: depthwise 3x3 stride 1 pad 1 on bf16 NHWC with folded BatchNorm.
//   ShuffleNetV2 branch2 depthwise (bias=True) + BN           nets/yolo_mul.py:144-146
//   RepGhostModule cheap_operation + fusion_bn identity branch nets/repghost.py:98-115 (the identity-BN
//   is folded into the centre tap exactly as switch_to_deploy does, :117-123), optional SiLU (:105-106,115)
//   and the bottleneck's identity shortcut (:279).
//
// Memory-bound (9 MAC per element).  Two paths:
//   dwconv_tma_kernel   plain [N,H,W,C] views: persistent CTAs; per tile one cp.async.bulk.tensor box load of
//       the (TH+2) x (TW+2) x 64-channel halo window (out-of-image pixels zero-filled by the TMA = conv
//       padding), double-buffered; threads read 128-bit channel vectors from shared memory, accumulate in
//       fp32, stage the bf16 result and one thread issues the TMA store (partial tiles clipped by hardware).
//       Round-1 profile of the first version (direct global loads, 4 pixels x 8 channels per thread):
//       1.4 TB/s, 16 warps/SM at 91 registers, latency-bound.
//   dwconv3x3_kernel    fallback for grouped / oddly strided views (direct 128-bit global loads).
#include <cuda.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace dcfa {
namespace {

// ------------------------------------------------------------------------------------------ TMA path
constexpr int TH = 8, TW = 16;                 // output tile (pixels)
constexpr int HH = TH + 2, HW = TW + 2;        // halo window
constexpr int CB = 64;                         // channels per tile (128-byte rows)
constexpr int WTAP = (CB / 8) * 12;            // floats per tap in shared memory: 8 taps + 4 floats of padding per 8-channel chunk, so that
                                               // the eight chunks a quarter warp reads (48 bytes apart) hit eight different bank groups
constexpr int kDwThreads = 256;
constexpr int IN_BYTES = HH * HW * CB * 2;     // 23040
constexpr int IN_BUF = (IN_BYTES + 1023) / 1024 * 1024;
constexpr int OUT_BYTES = TH * TW * CB * 2;    // 16384
constexpr int kMaxNin = 4;                     // input ring depth limit.  Round 1: with one box in flight per CTA the kernel
                                               // spent 26 % of its samples waiting for the TMA load (profiles/)
                                               // (~98 KB per CTA -> two CTAs per SM)

struct DwTmaArgs {
  const float* w;     // [G][9][C]
  const float* bias;  // [G][C]
  int n_img, group_imgs, H, W, C, act;
  int cblocks;        // C / 64 (C % 64 == 0) or 1 with cb = C (C in {16, 32})
  int cb;             // channels per tile
  int tiles_x, tiles_y;
  int total_tiles;
  int has_res;
  int nin;            // input ring depth (2..kMaxNin): loads run nin-1 tiles ahead
  int in_buf;         // bytes reserved per input ring slot (1024-byte multiple)
  int out_buf;        // bytes per staging / residual tile buffer (1024-byte multiple)
  int res_slots;      // residual buffers: nin (ringed with the inputs), 1 (loads then run one tile ahead only) or 0
};

__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap* map, int c, int x, int y, int n, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c), "r"(x), "r"(y), "r"(n), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_box(const CUtensorMap* map, uint32_t src, int c, int x, int y, int n) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c), "r"(x), "r"(y), "r"(n)
               : "memory");
}

struct DwTile {
  int cbi, tx, ty, n;
};
// Tile coordinates advanced incrementally (tile += gridDim.x): no divisions in the per-tile path.
struct DwIter {
  DwTile t, s;   // current tile, and the mixed-radix digits of the step
  template <typename P>
  __device__ __forceinline__ void init(const P& p, int tile, int step);
  template <typename P>
  __device__ __forceinline__ void advance(const P& p) {
    t.cbi += s.cbi;
    if (t.cbi >= p.cblocks) { t.cbi -= p.cblocks; t.tx += 1; }
    t.tx += s.tx;
    if (t.tx >= p.tiles_x) { t.tx -= p.tiles_x; t.ty += 1; }
    t.ty += s.ty;
    if (t.ty >= p.tiles_y) { t.ty -= p.tiles_y; t.n += 1; }
    t.n += s.n;
  }
};
__device__ __forceinline__ DwTile dw_decode(const DwTmaArgs& p, int tile) {
  DwTile t;
  t.cbi = tile % p.cblocks; tile /= p.cblocks;
  t.tx = tile % p.tiles_x; tile /= p.tiles_x;
  t.ty = tile % p.tiles_y;
  t.n = tile / p.tiles_y;
  return t;
}

template <typename P>
__device__ __forceinline__ void DwIter::init(const P& p, int tile, int step) {
  t = dw_decode(p, tile);
  s = dw_decode(p, step);
}

// One thread = one 8-channel chunk (c8) of one tile column, R consecutive output rows (R * c8n * TW * parts == 256 * R
// covers the TH x TW tile).  Per kernel column q the 3 weight vectors stay in registers while the thread walks
// down R + 2 input rows, each input vector feeding up to 3 output rows: 3*(R+2) + 9 shared-memory loads for R
// outputs instead of 27 per output.
template <int R>
__device__ __forceinline__ void dw_tile_compute(const DwTmaArgs& p, const uint8_t* in, const uint8_t* res, uint8_t* out,
                                                const float* w_s, const float* b_s, int tid, int row_bytes) {
  const int c8n = p.cb >> 3;
  const int c8 = tid % c8n;
  const int col = (tid / c8n) % TW;
  const int oy0 = (tid / (c8n * TW)) * R;
  F2 acc2[R][4];
  {
    const float4 b0 = *reinterpret_cast<const float4*>(b_s + c8 * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(b_s + c8 * 8 + 4);
#pragma unroll
    for (int o = 0; o < R; ++o) {
      acc2[o][0] = f2_make(b0.x, b0.y); acc2[o][1] = f2_make(b0.z, b0.w);
      acc2[o][2] = f2_make(b1.x, b1.y); acc2[o][3] = f2_make(b1.z, b1.w);
    }
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    F2 w[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float4 w0 = *reinterpret_cast<const float4*>(w_s + (r * 3 + q) * WTAP + c8 * 12);
      const float4 w1 = *reinterpret_cast<const float4*>(w_s + (r * 3 + q) * WTAP + c8 * 12 + 4);
      w[r][0] = f2_make(w0.x, w0.y); w[r][1] = f2_make(w0.z, w0.w);
      w[r][2] = f2_make(w1.x, w1.y); w[r][3] = f2_make(w1.z, w1.w);
    }
#pragma unroll
    for (int ri = 0; ri < R + 2; ++ri) {
      F2 v[4];
      unpack8_f2(*reinterpret_cast<const uint4*>(in + ((oy0 + ri) * HW + col + q) * row_bytes + c8 * 16), v);
#pragma unroll
      for (int o = 0; o < R; ++o) {
        const int r = ri - o;   // kernel row that maps input row ri to output row o (compile-time after unrolling)
        if (r >= 0 && r < 3) {
#pragma unroll
          for (int e = 0; e < 4; ++e) f2_fma(acc2[o][e], v[e], w[r][e]);
        }
      }
    }
  }
  float acc[R][8];
#pragma unroll
  for (int o = 0; o < R; ++o)
#pragma unroll
    for (int e = 0; e < 4; ++e) f2_get(acc2[o][e], acc[o][2 * e], acc[o][2 * e + 1]);
#pragma unroll
  for (int o = 0; o < R; ++o) {
    const int pp = (oy0 + o) * TW + col;
    if (p.act == DCFA_ACT_SILU) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[o][e] = silu_fast(acc[o][e]);
    } else if (p.act == DCFA_ACT_RELU) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[o][e] = fmaxf(acc[o][e], 0.0f);
    }
    if (p.has_res) {
      float rr[8];
      unpack8(*reinterpret_cast<const uint4*>(res + pp * row_bytes + c8 * 16), rr);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[o][e] += rr[e];
    }
    *reinterpret_cast<uint4*>(out + pp * row_bytes + c8 * 16) = pack8(acc[o]);
  }
}

// R = output rows per thread = channel block / 16; the narrower the block, the smaller the register tile and the
// shared-memory footprint, so more CTAs fit an SM (the kernel is latency-bound: per tile it waits for a box, computes
// and stores, with block-wide barriers in between).
template <int R>
__global__ void __launch_bounds__(kDwThreads, R == 4 ? 2 : (R == 2 ? 3 : 4)) dwconv_tma_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                   const __grid_constant__ CUtensorMap map_y,
                                                                   const __grid_constant__ CUtensorMap map_r,
                                                                   const DwTmaArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
  // layout: in[nin] | res[nin] (only with a residual) | out[2] | w (9*64 fp32) | bias (64 fp32) | barriers; buffer
  // sizes follow the channel block (p.in_buf, p.out_buf), so that narrow layers get a deeper ring at two CTAs per SM
  const int NIN = p.nin;
  const uint32_t IN_BUF = (uint32_t)p.in_buf, OUT_BUF = (uint32_t)p.out_buf;
  const uint32_t res_ring = p.res_slots > 1 ? OUT_BUF : 0u;   // residual slot pitch (0: one shared buffer)
  const uint32_t res_total = (uint32_t)p.res_slots * OUT_BUF;
  const uint32_t s_in = base;
  const uint32_t s_res = s_in + NIN * IN_BUF;
  const uint32_t s_out = s_res + res_total;
  float* w_s = reinterpret_cast<float*>(gbase + NIN * IN_BUF + res_total + 2 * OUT_BUF);
  float* b_s = w_s + 9 * WTAP;
  const uint32_t bars = s_out + 2 * OUT_BUF + (9 * WTAP + CB) * 4;
  const uint32_t bar_in = bars;               // nin barriers (input box [+ residual box] landed)

  const int tid = threadIdx.x;
  ptx::pdl_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < NIN; ++i) ptx::mbar_init(bar_in + 8u * i, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  ptx::pdl_wait();

  const int row_bytes = p.cb * 2;             // bytes per pixel row in shared memory (dense, no swizzle)
  const uint32_t in_bytes = (uint32_t)(HH * HW) * row_bytes;
  const uint32_t out_bytes = (uint32_t)(TH * TW) * row_bytes;

  auto issue = [&](const DwTile& t, int slot) {   // one thread: input halo box (+ residual box) of tile t
    const uint32_t bar = bar_in + 8u * slot;
    ptx::mbar_arrive_expect_tx(bar, in_bytes + (p.has_res ? out_bytes : 0u));
    tma_load_box(s_in + (uint32_t)slot * IN_BUF, &map_x, t.cbi * p.cb, t.tx * TW - 1, t.ty * TH - 1, t.n, bar);
    if (p.has_res) tma_load_box(s_res + (uint32_t)slot * res_ring, &map_r, t.cbi * p.cb, t.tx * TW, t.ty * TH, t.n, bar);
  };
  // with a single residual buffer the residual box of tile i+2 must not land before tile i+1 consumed its own:
  // the ring then runs one tile ahead only
  const int ahead = (p.has_res && p.res_slots == 1) ? 1 : NIN - 1;
  DwIter cur, pre;                             // tile being processed; tile whose box is requested next (thread 32)
  cur.init(p, blockIdx.x, gridDim.x);
  pre = cur;
  if (tid == 32) {
    for (int d = 0; d < ahead; ++d) {
      if (pre.t.n < p.n_img) issue(pre.t, d);
      pre.advance(p);
    }
  }

  int cur_key = -1;
  uint32_t it = 0;
  int slot = 0;
  uint32_t ring_ph = 0;                        // parity of the ring pass `slot` is in
  for (; cur.t.n < p.n_img; cur.advance(p), ++it) {
    const DwTile t = cur.t;
    int g = 0;
    for (int nn = t.n; nn >= p.group_imgs; nn -= p.group_imgs) ++g;   // n / group_imgs without a division (1 or 2 groups)
    const int key = g * p.cblocks + t.cbi;
    if (key != cur_key) {                     // this (group, channel block)'s weights and bias
      __syncthreads();
      for (int i = tid; i < 9 * p.cb; i += kDwThreads) {
        const int tap = i / p.cb, c = i - tap * p.cb;
        w_s[tap * WTAP + (c >> 3) * 12 + (c & 7)] = __ldg(p.w + ((int64_t)g * 9 + tap) * p.C + t.cbi * p.cb + c);
      }
      for (int i = tid; i < p.cb; i += kDwThreads) b_s[i] = __ldg(p.bias + (int64_t)g * p.C + t.cbi * p.cb + i);
      cur_key = key;
      __syncthreads();
    }
    ptx::mbar_wait(bar_in + 8u * slot, ring_ph);
    const uint8_t* in = gbase + slot * IN_BUF;
    const uint8_t* res = gbase + NIN * IN_BUF + slot * res_ring;
    uint8_t* out = gbase + NIN * IN_BUF + res_total + (it & 1u) * OUT_BUF;

    // the staging buffer (it & 1) was last read by the TMA store issued two tiles ago
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();

    dw_tile_compute<R>(p, in, res, out, w_s, b_s, tid, row_bytes);
    ptx::fence_proxy_async_smem();   // staged tile -> async proxy (TMA store); no bulk loads are owned by these threads
    __syncthreads();                 // input slot and residual buffer consumed, staging complete
    if (tid == 32) {                 // refill the ring (this thread owns all bulk loads)
      if (pre.t.n < p.n_img) issue(pre.t, slot + ahead >= NIN ? slot + ahead - NIN : slot + ahead);
      pre.advance(p);
    }
    if (tid == 0) {                  // this thread owns all bulk stores
      tma_store_box(&map_y, s_out + (it & 1u) * OUT_BUF, t.cbi * p.cb, t.tx * TW, t.ty * TH, t.n);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (++slot == NIN) { slot = 0; ring_ph ^= 1u; }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn dw_encode_fn() {
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

// rank-4 map (C, W, H, N) over a plain NHWC bf16 view, box (cb, bw, bh, 1), no swizzle
int make_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int N, int ld, int64_t img_stride, int cb, int bw, int bh) {
  EncodeTiledFn enc = dw_encode_fn();
  if (!enc) return fail(DCFA_E_CUDA, "dwconv: cuTensorMapEncodeTiled entry point unavailable");
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)img_stride * 2};
  const cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)bw, (cuuint32_t)bh, 1u};
  const cuuint32_t es[4] = {1u, 1u, 1u, 1u};
  const int64_t run = (C == ld) ? (int64_t)C * 2 * W : (int64_t)C * 2;
  const CUtensorMapL2promotion promo = run >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : (run >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B);
  CUresult cr = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(DCFA_E_CUDA, "dwconv: cuTensorMapEncodeTiled failed with %d", (int)cr);
  return DCFA_OK;
}

// ------------------------------------------------------------------------------------------ fallback path
constexpr int XT = 4;

struct DwArgs {
  View<const __nv_bfloat16> x;
  View<const __nv_bfloat16> res;
  View<__nv_bfloat16> y;
  const float* w;     // [G][9][C]
  const float* bias;  // [G][C]
  int n_img, group_imgs, H, W, C, act;
  int xg;             // ceil(W / XT)
  int64_t total;
};

__global__ void __launch_bounds__(256) dwconv3x3_kernel(const DwArgs p) {
  const int c8n = p.C >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int c8 = (int)(t % c8n); t /= c8n;
    const int xg = (int)(t % p.xg); t /= p.xg;
    const int y = (int)(t % p.H);
    const int n = (int)(t / p.H);
    const int g = n / p.group_imgs;
    const int c = c8 * 8;
    const int x0 = xg * XT;

    const float* wp = p.w + ((int64_t)g * 9) * p.C + c;
    const float* bp = p.bias + (int64_t)g * p.C + c;
    float acc[XT][8];
    {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bp));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bp) + 1);
#pragma unroll
      for (int j = 0; j < XT; ++j) {
        acc[j][0] = b0.x; acc[j][1] = b0.y; acc[j][2] = b0.z; acc[j][3] = b0.w;
        acc[j][4] = b1.x; acc[j][5] = b1.y; acc[j][6] = b1.z; acc[j][7] = b1.w;
      }
    }
    const __nv_bfloat16* xin = p.x.p + p.x.img_off(n) + c;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int iy = y + r - 1;
      if (iy < 0 || iy >= p.H) continue;
      float wr[3][8];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)(r * 3 + q) * p.C));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)(r * 3 + q) * p.C) + 1);
        wr[q][0] = w0.x; wr[q][1] = w0.y; wr[q][2] = w0.z; wr[q][3] = w0.w;
        wr[q][4] = w1.x; wr[q][5] = w1.y; wr[q][6] = w1.z; wr[q][7] = w1.w;
      }
#pragma unroll
      for (int q = 0; q < XT + 2; ++q) {
        const int ix = x0 + q - 1;
        if (ix < 0 || ix >= p.W) continue;
        float v[8];
        unpack8(ldg128(xin + (int64_t)(iy * p.W + ix) * p.x.ld), v);
#pragma unroll
        for (int j = 0; j < XT; ++j) {
          const int tap = q - j;  // input column q feeds output j through kernel column q - j
          if (tap >= 0 && tap < 3) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(v[e], wr[tap][e], acc[j][e]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < XT; ++j) {
      const int ox = x0 + j;
      if (ox >= p.W) continue;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = apply_act(acc[j][e], p.act);
      const int64_t pix = (int64_t)y * p.W + ox;
      if (p.res.p) {
        float rr[8];
        unpack8(ldg128(p.res.p + p.res.img_off(n) + pix * p.res.ld + c), rr);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] += rr[e];
      }
      stg128(p.y.p + p.y.img_off(n) + pix * p.y.ld + c, pack8(o));
    }
  }
}

template <typename T>
bool plain_view(const View<T>& v) {
  return v.gi <= 0 || v.gstride == (int64_t)v.gi * v.img_stride;
}

}  // namespace

int launch_dwconv(const dcfa_op& op, void* const* bufs, cudaStream_t st) {
  DwArgs a;
  a.x = resolve<const __nv_bfloat16>(op.x, bufs);
  a.res = resolve<const __nv_bfloat16>(op.x2, bufs);
  a.y = resolve<__nv_bfloat16>(op.y, bufs);
  a.w = resolve_ptr<const float>(op.w, bufs);
  a.bias = resolve_ptr<const float>(op.bias, bufs);
  a.n_img = op.n_img;
  a.group_imgs = op.group_imgs > 0 ? op.group_imgs : op.n_img;
  a.H = op.Hi; a.W = op.Wi; a.C = op.Cin; a.act = op.act;
  DCFA_REQUIRE(a.x.p && a.y.p && a.w && a.bias, "dwconv: missing tensor");
  DCFA_REQUIRE(a.C > 0 && a.C % 8 == 0, "dwconv: C %d must be a multiple of 8", a.C);
  DCFA_REQUIRE(a.n_img > 0 && a.n_img % a.group_imgs == 0, "dwconv: bad grouping");
  DCFA_REQUIRE(((uintptr_t)a.x.p % 16) == 0 && a.x.ld % 8 == 0 && a.x.img_stride % 8 == 0 && a.x.gstride % 8 == 0 &&
                   ((uintptr_t)a.y.p % 16) == 0 && a.y.ld % 8 == 0 && a.y.img_stride % 8 == 0 && a.y.gstride % 8 == 0,
               "dwconv: views must be 16-byte aligned");
  if (a.res.p)
    DCFA_REQUIRE(((uintptr_t)a.res.p % 16) == 0 && a.res.ld % 8 == 0 && a.res.img_stride % 8 == 0 && a.res.gstride % 8 == 0,
                 "dwconv: residual view must be 16-byte aligned");
  DCFA_REQUIRE(((uintptr_t)a.w % 16) == 0 && ((uintptr_t)a.bias % 16) == 0, "dwconv: params must be 16-byte aligned");

  // ---- TMA path: plain views, channel count a multiple of 64 or exactly 16 / 32
  const bool cb_ok = (a.C % CB == 0) || a.C == 32 || a.C == 16;
  if (cb_ok && plain_view(a.x) && plain_view(a.y) && (!a.res.p || plain_view(a.res))) {
    DwTmaArgs t;
    t.w = a.w; t.bias = a.bias;
    t.n_img = a.n_img; t.group_imgs = a.group_imgs; t.H = a.H; t.W = a.W; t.C = a.C; t.act = a.act;
    t.cb = a.C % CB == 0 ? CB : a.C;
    t.cblocks = a.C / t.cb;
    t.tiles_x = ceil_div(a.W, TW);
    t.tiles_y = ceil_div(a.H, TH);
    const int64_t total = (int64_t)a.n_img * t.tiles_x * t.tiles_y * t.cblocks;
    DCFA_REQUIRE(total < (1ll << 31), "dwconv: too many tiles");
    t.total_tiles = (int)total;
    t.has_res = a.res.p ? 1 : 0;
    alignas(64) CUtensorMap mx, my, mr;
    memset(&mr, 0, sizeof(mr));
    int rc = make_map(&mx, a.x.p, a.C, a.W, a.H, a.n_img, a.x.ld, a.x.img_stride, t.cb, HW, HH);
    if (rc) return rc;
    rc = make_map(&my, a.y.p, a.C, a.W, a.H, a.n_img, a.y.ld, a.y.img_stride, t.cb, TW, TH);
    if (rc) return rc;
    if (a.res.p) {
      rc = make_map(&mr, a.res.p, a.C, a.W, a.H, a.n_img, a.res.ld, a.res.img_stride, t.cb, TW, TH);
      if (rc) return rc;
    }
    // CTAs per SM by channel block (see the kernel's launch bounds); ring depth: as deep as that leaves room for
    const int ctas = t.cb == 64 ? 2 : (t.cb == 32 ? 3 : 4);
    t.in_buf = (HH * HW * t.cb * 2 + 1023) / 1024 * 1024;
    t.out_buf = (TH * TW * t.cb * 2 + 1023) / 1024 * 1024;
    const int budget = (227 * 1024) / ctas - 1536;
    const int fixed = 1024 + 2 * t.out_buf + (9 * WTAP + CB) * 4 + 64;
    const int per_slot = t.in_buf + (t.has_res ? t.out_buf : 0);
    t.nin = std::min(kMaxNin, (budget - fixed) / per_slot);
    t.res_slots = t.has_res ? t.nin : 0;
    if (t.nin < 2) {                 // wide channel block with a residual: one residual buffer, inputs double-buffered
      t.nin = 2;
      t.res_slots = 1;
    }
    const size_t smem = (size_t)fixed + (size_t)t.nin * t.in_buf + (size_t)t.res_slots * t.out_buf;
    static DeviceOnce attr_set;
    if (attr_set.needed()) {
      const int mx_smem = 1024 + kMaxNin * (IN_BUF + OUT_BYTES) + 2 * OUT_BYTES + (9 * WTAP + CB) * 4 + 64;
      cudaError_t e = cudaFuncSetAttribute(dwconv_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx_smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx_smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(dwconv_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx_smem);
      if (e != cudaSuccess) return fail(DCFA_E_CUDA, "dwconv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      attr_set.mark();
    }
    int64_t grid = (int64_t)sm_count() * ctas;
    if (grid > total) grid = total;
    if (t.cb == 64) launch_pdl(dwconv_tma_kernel<4>, dim3((unsigned)grid), dim3(kDwThreads), smem, st, mx, my, mr, t);
    else if (t.cb == 32) launch_pdl(dwconv_tma_kernel<2>, dim3((unsigned)grid), dim3(kDwThreads), smem, st, mx, my, mr, t);
    else launch_pdl(dwconv_tma_kernel<1>, dim3((unsigned)grid), dim3(kDwThreads), smem, st, mx, my, mr, t);
    DCFA_CHECK_LAUNCH("dwconv_tma_kernel");
    return DCFA_OK;
  }

  a.xg = ceil_div(a.W, XT);
  a.total = (int64_t)a.n_img * a.H * a.xg * (a.C >> 3);
  int64_t blocks = (a.total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  launch_k(dwconv3x3_kernel, dim3((unsigned)blocks), dim3(256), 0, st, 0, false, a);
  DCFA_CHECK_LAUNCH("dwconv3x3_kernel");
  return DCFA_OK;
}

}  // namespace dcfa
