// dcfa_nms: DecodeBox.non_max_suppression (utils/utils_bbox.py:87-168) without its per-image Python loop
// and its three host synchronisations, bit-exact in the kept indices.
//
//   nms_prepare  xywh -> xyxy in place (:92-97), class max with "first max wins" (:106), conf >= thr (:111),
//                one 64-bit sort key per anchor:  class(8) | ~orderable(conf)(32) | anchor(24)
//                (non-candidates get the all-ones sentinel).
//   nms_sort     one CTA per image, bitonic sort of the keys (shared memory up to 16384 keys, global above):
//                ascending key order == classes ascending (:130,:136), scores descending, ties by lower
//                anchor index -- exactly torch's stable descending sort that torchvision.ops.nms uses.
//   nms_greedy   one CTA per image.  Sorted candidates are processed in chunks of 64: the chunk's 64x64
//                suppression bits are computed in parallel, resolved serially by one thread, and only the
//                KEPT boxes of the chunk are then tested against all later candidates.  This equals the
//                greedy scan of torchvision (only kept boxes ever suppress) at O(kept * n) instead of O(n^2).
//
// IoU arithmetic follows torchvision.ops.nms, the un-vendored dependency the reference calls at :145:
//   DCFA_IOU_TV_CPU   areas rounded separately, (double)iou > thr          (torchvision/csrc/ops/cpu/nms_kernel.cpp)
//   DCFA_IOU_TV_CUDA  Sa + Sb evaluated as fma(wb, hb, Sa), iou > (float)thr (SASS of torchvision 0.26.0+cu128
//                     nms_kernel_impl<float> for sm_100: FMUL, FFMA, FADD, IEEE divide, FSETP.GT on F2F.F32.F64(thr))
// All box arithmetic uses _rn intrinsics so nvcc cannot contract anything else.
#include <cooperative_groups.h>
#include <cstdio>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dcfa {
namespace {

constexpr uint64_t kSentinel = ~0ull;
constexpr int kNmsThreads = 1024;
constexpr int kChunk = 64;
constexpr int kSmemSortMax = 16384;

__device__ __forceinline__ uint32_t orderable(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct PrepArgs {
  float* pred;
  uint64_t* keys;   // [B][Apad], candidates compacted to the front (any order: the key is a total order)
  int32_t* cand;    // [B] candidate counters, zeroed before the launch
  int B, A, Apad, nc;
  float conf_thres;
};

__global__ void __launch_bounds__(256) nms_prepare_kernel(const PrepArgs p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // Apad is a multiple of 32 (or the whole image fits one warp), so a warp never straddles two images
  const int b = (int)(i / p.Apad);
  const int a = (int)(i - (int64_t)b * p.Apad);
  uint64_t key = kSentinel;
  if (b < p.B && a < p.A) {
    float* row = p.pred + ((int64_t)b * p.A + a) * (4 + p.nc);
    const float cx = row[0], cy = row[1], w = row[2], h = row[3];
    const float hw = __fdiv_rn(w, 2.0f), hh = __fdiv_rn(h, 2.0f);
    row[0] = __fsub_rn(cx, hw);
    row[1] = __fsub_rn(cy, hh);
    row[2] = __fadd_rn(cx, hw);
    row[3] = __fadd_rn(cy, hh);
    float best = row[4];
    int bi = 0;
    for (int c = 1; c < p.nc; ++c) {
      const float v = row[4 + c];
      if (best != best) break;                    // torch.max propagates NaN
      if (v > best || v != v) { best = v; bi = c; }
    }
    if (best >= p.conf_thres) {                   // false for NaN
      const float conf = best + 0.0f;             // -0.0 -> +0.0 (equal under torch.sort)
      key = ((uint64_t)bi << 56) | ((uint64_t)(~orderable(conf)) << 24) | (uint64_t)a;
    }
  }
  // warp-aggregated append
  const unsigned mask = __ballot_sync(0xffffffffu, key != kSentinel);
  if (mask) {
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(p.cand + b, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (key != kSentinel) p.keys[(int64_t)b * p.Apad + base + __popc(mask & ((1u << lane) - 1u))] = key;
  }
}

struct SortArgs {
  uint64_t* keys;        // [B][Apad]
  const int32_t* cand;   // [B]
  int Apad;
};

// Fallback for images with more than 16384 candidates: one CTA per image, bitonic sort in global memory.
__global__ void __launch_bounds__(kNmsThreads) nms_sort_global_kernel(const SortArgs p) {
  uint64_t* k = p.keys + (int64_t)blockIdx.x * p.Apad;
  const int n = p.cand[blockIdx.x];
  if (n <= kSmemSortMax) return;            // nms_sort_kernel handled it
  int N = 2;
  while (N < n) N <<= 1;
  for (int i = n + threadIdx.x; i < N; i += kNmsThreads) k[i] = kSentinel;  // N <= Apad
  __syncthreads();
  const int half = N >> 1;
  for (int kk = 2; kk <= N; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < half; t += kNmsThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // t-th index with bit j clear
        const int ixj = i | j;
        const uint64_t a = k[i], b = k[ixj];
        const bool asc = (i & kk) == 0;
        if ((a > b) == asc) { k[i] = b; k[ixj] = a; }
      }
      __syncthreads();
    }
  }
}

// Cluster-wide bitonic sort of up to CS * 1024 * E keys (CS = CTAs of the image's cluster).  CTA `rank` owns the
// keys [rank * 1024E, (rank + 1) * 1024E); thread t owns E consecutive ones in registers.  Compare-exchange
// distances below E are register moves, below 32E warp shuffles, below 1024E passes over the CTA's (padded,
// conflict-free) shared memory, and the remaining log2(CS) distances read the partner CTA's shared memory
// through DSMEM -- each side keeps its min or max, so nothing is written remotely.
// Stages beyond next_pow2(n) only see sentinels and are skipped.
__device__ __forceinline__ int sort_pad(int i) { return i + (i >> 4); }

template <int E>
__device__ __forceinline__ void cluster_sort(uint64_t* g, int n, uint64_t* s_keys, cg::cluster_group& cluster) {
  constexpr int Nloc = kNmsThreads * E;
  const int rank = (int)cluster.block_rank();
  int Neff = 2;
  while (Neff < n) Neff <<= 1;
  const int t = threadIdx.x, lane = t & 31;
  const int lbase = t * E;                  // index inside the CTA
  const int base = rank * Nloc + lbase;     // index inside the image
  uint64_t v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) v[e] = base + e < n ? g[base + e] : kSentinel;

  for (int kk = 2; kk <= Neff; kk <<= 1) {
    int j = kk >> 1;
    for (; j >= Nloc; j >>= 1) {            // partner key lives in CTA rank ^ (j / Nloc), same local index
#pragma unroll
      for (int e = 0; e < E; ++e) s_keys[sort_pad(lbase + e)] = v[e];
      cluster.sync();
      const int d = j / Nloc;
      const uint64_t* remote = cluster.map_shared_rank(s_keys, rank ^ d);
      const bool lower = (rank & d) == 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const uint64_t o = remote[sort_pad(lbase + e)];
        const bool asc = ((base + e) & kk) == 0;
        const uint64_t lo = v[e] < o ? v[e] : o, hi = v[e] < o ? o : v[e];
        v[e] = (lower == asc) ? lo : hi;
      }
      cluster.sync();                        // the partner has read before anybody overwrites
    }
    if (j >= 32 * E) {
#pragma unroll
      for (int e = 0; e < E; ++e) s_keys[sort_pad(lbase + e)] = v[e];
      __syncthreads();
      for (; j >= 32 * E; j >>= 1) {
        for (int c = t; c < Nloc / 2; c += kNmsThreads) {
          const int i = ((c & ~(j - 1)) << 1) | (c & (j - 1));   // c-th local index with bit j clear
          const int pi = sort_pad(i), pj = sort_pad(i | j);
          const uint64_t a = s_keys[pi], b = s_keys[pj];
          const bool asc = ((rank * Nloc + i) & kk) == 0;
          if ((a > b) == asc) { s_keys[pi] = b; s_keys[pj] = a; }
        }
        __syncthreads();
      }
#pragma unroll
      for (int e = 0; e < E; ++e) v[e] = s_keys[sort_pad(lbase + e)];
    }
    // distances E .. 16E: the partner key sits in lane ^ (j / E)
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      if (d * E <= (kk >> 1)) {
        const bool lower = (lane & d) == 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const uint64_t o = __shfl_xor_sync(0xffffffffu, v[e], d);
          const bool asc = ((base + e) & kk) == 0;
          const uint64_t lo = v[e] < o ? v[e] : o, hi = v[e] < o ? o : v[e];
          v[e] = (lower == asc) ? lo : hi;
        }
      }
    }
    // distances below E: both keys are in this thread's registers
#pragma unroll
    for (int jj = E / 2; jj >= 1; jj >>= 1) {
      if (jj <= (kk >> 1)) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & jj) == 0) {
            const bool asc = ((base + e) & kk) == 0;
            const uint64_t a = v[e], b = v[e | jj];
            if ((a > b) == asc) { v[e] = b; v[e | jj] = a; }
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e)
    if (base + e < n) g[base + e] = v[e];
}

template <int MODE>
__device__ __forceinline__ bool suppresses(const float4& a, const float4& b, float thr_f, double thr_d) {
  // a = earlier (higher score) box, b = later box; both (x1, y1, x2, y2)
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  // Disjoint boxes: inter == 0, so the IoU is +-0 or NaN and never exceeds a threshold >= 0 -- skip the
  // divide (most pairs).  Negative thresholds take the full path.
  const bool nonneg = (MODE == DCFA_IOU_TV_CUDA) ? (thr_f >= 0.0f) : (thr_d >= 0.0);
  if (nonneg && !(right > left && bottom > top)) return false;
  const float w = fmaxf(__fsub_rn(right, left), 0.0f), h = fmaxf(__fsub_rn(bottom, top), 0.0f);
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float wb = __fsub_rn(b.z, b.x), hb = __fsub_rn(b.w, b.y);
  if (MODE == DCFA_IOU_TV_CUDA) {
    const float iou = __fdiv_rn(inter, __fsub_rn(__fmaf_rn(wb, hb, sa), inter));
    return iou > thr_f;
  } else {
    const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, __fmul_rn(wb, hb)), inter));
    return (double)iou > thr_d;
  }
}

struct GreedyArgs {
  const float* pred;     // xyxy already
  uint64_t* keys;        // [B][Apad] compacted candidate keys, sorted in place
  float4* sbox;          // [B][A] scratch: boxes in sorted order (used only when an image has more than `cap` candidates)
  float* out_det;
  int32_t* out_idx;
  int32_t* out_cnt;
  int32_t* out_cand;
  const int32_t* cand;
  int B, A, Apad, nc;
  int cap;               // candidates per image whose boxes + classes fit the shared-memory carve-out
  float thr_f;
  double thr_d;
};

#ifdef DCFA_NMS_TIMING
#define NT_DECL long long nt_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long nt_last = clock64(); int nt_rounds = 0;
#define NT(i) { const long long nt_now = clock64(); nt_t[i] += nt_now - nt_last; nt_last = nt_now; }
#else
#define NT_DECL
#define NT(i)
#endif

// state of a sorted candidate
constexpr uint8_t kAlive = 0, kGone = 1, kKept = 2;

// exclusive rank of this thread's flag among the CTA's 1024 threads (ordered), total in `found`.
// One barrier: every warp redoes the 32-entry scan of the warp counts itself.
__device__ __forceinline__ int block_rank(int flag, int* s_wcnt, int& found) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) s_wcnt[warp] = __popc(bal);
  __syncthreads();
  const int c = s_wcnt[lane];
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  found = __shfl_sync(0xffffffffu, incl, 31);
  const int wbase = __shfl_sync(0xffffffffu, incl - c, warp);
  return wbase + __popc(bal & ((1u << lane) - 1u));
}

// One cluster of CS CTAs per image: sort, then the greedy rounds.
//
// Every CTA holds the whole image's sorted boxes, classes and alive/gone/kept states, and runs steps (1)-(3) of a
// round redundantly (identical inputs, identical results, no communication).  Only step (4) -- the kept boxes of
// the round against every later candidate, the bulk of the arithmetic -- is split: CTA r tests the positions
// next + r*1024 + tid (+ CS*1024 ...) and stores each "gone" into the state array of EVERY CTA through DSMEM;
// one cluster barrier closes the round.
// A CTA that has passed the barrier may start storing round k+1's "gone" marks while a slower CTA still scans the
// window of round k+1 in step (1): those marks only touch positions behind the 64th alive entry of the window
// (or behind the window when it holds fewer than 64), so the chunk, `next` and `found >= 64` come out the same
// in every CTA.
template <int MODE>
__global__ void __launch_bounds__(kNmsThreads) nms_cluster_kernel(const GreedyArgs p) {
  // dynamic shared memory, greedy phase: float4 box[cap] | uint8 state[A16] | uint8 cls[cap]
  //                        sort phase  : uint64 keys[sort_pad(1024 * E)]
  extern __shared__ __align__(16) uint8_t g_smem[];
  __shared__ float4 c_box[kChunk];
  __shared__ int c_cls[kChunk];
  __shared__ unsigned long long c_mask[kChunk];
  __shared__ float4 k_box[kChunk];
  __shared__ int k_cls[kChunk];
  __shared__ int c_pos[kChunk];
  __shared__ int s_wcnt[2][32];
  __shared__ int s_next;
  __shared__ unsigned long long s_keep;

  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int crank = (int)cluster.block_rank();
  const int b = blockIdx.x / CS;
  const int tid = threadIdx.x;
  uint64_t* keys = p.keys + (int64_t)b * p.Apad;
  const float* pred = p.pred + (int64_t)b * p.A * (4 + p.nc);
  const int n = p.cand[b];
  NT_DECL

  // ---- sort (images with more than 16384 candidates were sorted by nms_sort_global_kernel)
  if (n > 1 && n <= kSmemSortMax) {
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(g_smem);
    const int per_cta = (n + CS - 1) / CS;
    if (per_cta <= 1024) cluster_sort<1>(keys, n, s_keys, cluster);
    else if (per_cta <= 2048) cluster_sort<2>(keys, n, s_keys, cluster);
    else if (per_cta <= 4096) cluster_sort<4>(keys, n, s_keys, cluster);
    else if (per_cta <= 8192) cluster_sort<8>(keys, n, s_keys, cluster);
    else cluster_sort<16>(keys, n, s_keys, cluster);
    __threadfence();
    cluster.sync();   // every CTA's slice of the sorted keys is in global memory; the sort buffer is free
  }

  NT(0)
  // ---- greedy rounds
  // Boxes and classes of the sorted candidates live in shared memory when they fit (the round loop is a chain of
  // dependent steps: its latency is the cost); otherwise in the global scratch, read back through L2.
  const bool fits = n <= p.cap;
  float4* s_box = reinterpret_cast<float4*>(g_smem);
  uint8_t* state = g_smem + (size_t)p.cap * 16;
  uint8_t* s_cls = state + (size_t)((p.A + 15) / 16) * 16;
  float4* box = fits ? s_box : p.sbox + (int64_t)b * p.A;

  for (int j = tid; j < n; j += kNmsThreads) {
    const uint64_t key = __ldcg(keys + j);
    const int a = (int)(key & 0xFFFFFFull);
    const float* row = pred + (int64_t)a * (4 + p.nc);
    const float4 bx = make_float4(row[0], row[1], row[2], row[3]);
    if (fits) { s_box[j] = bx; s_cls[j] = (uint8_t)(key >> 56); }
    else if (crank == 0) box[j] = bx;     // one writer for the shared global scratch
    state[j] = kAlive;
  }
  if (!fits) __threadfence();
  cluster.sync();     // also: nobody stores a remote "gone" before every CTA has initialised its states

  NT(1)
  // Each round takes the next (up to) 64 ALIVE candidates in sorted order -- found by an ordered block-wide
  // compaction over a window of 1024 positions -- so the number of rounds scales with the candidates that
  // survive the boxes kept so far, not with the number of candidates.
  int pos0 = 0, par = 0;
  while (pos0 < n) {
    // (1) ordered compaction of the alive positions in [pos0, pos0 + 1024)
    const int j0 = pos0 + tid;
    const int alive = (j0 < n && state[j0] == kAlive) ? 1 : 0;
    int found;
    const int rank = block_rank(alive, s_wcnt[par], found);
    par ^= 1;                                            // the next scan must not overwrite counts still being read
    if (found == 0) { pos0 += kNmsThreads; continue; }   // uniform over the cluster
    if (alive && rank < kChunk) {
      c_pos[rank] = j0;
      c_box[rank] = box[j0];
      c_cls[rank] = fits ? (int)s_cls[j0] : (int)(__ldcg(keys + j0) >> 56);
      if (rank == kChunk - 1) s_next = j0 + 1;   // the next round resumes right after the 64th alive entry
    }
    if (tid == 0 && found < kChunk) s_next = pos0 + kNmsThreads;
    __syncthreads();
    const int cn = min(found, kChunk);
    const int next = s_next;
    NT(2)
    // (2) suppression bits inside the chunk: bit i of c_mask[j] <=> box i (< j) suppresses box j.  The 2016
    //     pairs are spread evenly by folding the triangle: warp w owns rows w and 63 - w (63 pairs, two per lane);
    //     the lanes' bits are OR-reduced with redux.sync, so every row is written exactly once, without atomics.
    {
      const int w = tid >> 5, lane = tid & 31;
      unsigned a_lo = 0u, a_hi = 0u, b_lo = 0u, b_hi = 0u;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int k = lane + 32 * q;
        if (k < 63) {
          const bool first = k < w;
          const int j = first ? w : 63 - w;
          const int i = first ? k : k - w;
          if (j < cn && c_cls[i] == c_cls[j] && suppresses<MODE>(c_box[i], c_box[j], p.thr_f, p.thr_d)) {
            const unsigned bit = 1u << (i & 31);
            if (first) { if (i < 32) a_lo |= bit; else a_hi |= bit; }
            else { if (i < 32) b_lo |= bit; else b_hi |= bit; }
          }
        }
      }
      a_lo = __reduce_or_sync(0xffffffffu, a_lo); a_hi = __reduce_or_sync(0xffffffffu, a_hi);
      b_lo = __reduce_or_sync(0xffffffffu, b_lo); b_hi = __reduce_or_sync(0xffffffffu, b_hi);
      if (lane == 0) {
        c_mask[w] = ((unsigned long long)a_hi << 32) | a_lo;
        c_mask[63 - w] = ((unsigned long long)b_hi << 32) | b_lo;
      }
    }
    __syncthreads();
    NT(3)
    // (3) resolution by warp 0, lane l holding candidates l and l + 32: a candidate is dead once a kept box
    //     suppresses it and kept once all its suppressors are dead; every pass decides at least the first
    //     undecided candidate, in practice a handful of passes decide all 64.
    if (tid < 32) {
      const unsigned long long mA = c_mask[tid], mB = c_mask[tid + 32];
      bool uA = tid < cn, uB = tid + 32 < cn;
      unsigned long long keep = 0ull, dead = 0ull;
      while (__any_sync(0xffffffffu, uA || uB)) {
        const bool dA = uA && (mA & keep) != 0ull, dB = uB && (mB & keep) != 0ull;
        const bool kA = uA && !dA && (mA & ~dead) == 0ull, kB = uB && !dB && (mB & ~dead) == 0ull;
        keep |= (unsigned long long)__ballot_sync(0xffffffffu, kA) | ((unsigned long long)__ballot_sync(0xffffffffu, kB) << 32);
        dead |= (unsigned long long)__ballot_sync(0xffffffffu, dA) | ((unsigned long long)__ballot_sync(0xffffffffu, dB) << 32);
        uA = uA && !dA && !kA;
        uB = uB && !dB && !kB;
      }
      if (tid == 0) s_keep = keep;
    }
    __syncthreads();
    const unsigned long long keep64 = s_keep;
    const int kc = __popcll(keep64);
    if (tid < cn) {
      const bool kept = (keep64 >> tid) & 1ull;
      if (kept) {
        const int r = __popcll(keep64 & ((1ull << tid) - 1ull));
        k_box[r] = c_box[tid];
        k_cls[r] = c_cls[tid];
      }
      state[c_pos[tid]] = kept ? kKept : kGone;   // every entry of the chunk is now decided (same in every CTA)
    }
    __syncthreads();
    NT(4)
    // (4) the boxes kept in this round suppress the later candidates; this CTA's share of the positions
    //     (warp-interleaved, so that every CTA sees the same mix of early and late positions)
    for (int j = next + ((tid >> 5) * CS + crank) * 32 + (tid & 31); j < n; j += CS * kNmsThreads) {
      if (state[j] != kAlive) continue;
      const float4 bj = box[j];
      const int cj = fits ? (int)s_cls[j] : (int)(__ldcg(keys + j) >> 56);
      for (int k = 0; k < kc; ++k) {
        if (k_cls[k] == cj && suppresses<MODE>(k_box[k], bj, p.thr_f, p.thr_d)) {
          for (int r = 0; r < CS; ++r) cluster.map_shared_rank(state, r)[j] = kGone;
          break;
        }
      }
    }
    NT(5)
    cluster.sync();
    NT(6)
    pos0 = next;
#ifdef DCFA_NMS_TIMING
    ++nt_rounds;
#endif
  }
#ifdef DCFA_NMS_TIMING
  if (b == 0 && (tid == 0 || tid == 1023))
    printf("nms timing cta %d tid %d rounds %d: sort %lld init %lld | step1 %lld step2 %lld step3 %lld step4 %lld csync %lld\n", crank, tid,
           nt_rounds, nt_t[0], nt_t[1], nt_t[2], nt_t[3], nt_t[4], nt_t[5], nt_t[6]);
#endif
  if (crank != 0) return;

  // (5) kept candidates write themselves out, in sorted order (ordered compaction of the kKept states)
  int total = 0;
  for (int w0 = 0; w0 < n; w0 += kNmsThreads) {
    const int j = w0 + tid;
    const int kept = (j < n && state[j] == kKept) ? 1 : 0;
    int found;
    const int rank = block_rank(kept, s_wcnt[par], found);
    par ^= 1;
    if (kept) {
      const uint64_t key = __ldcg(keys + j);
      const int a = (int)(key & 0xFFFFFFull);
      const int cls = (int)(key >> 56);
      const int pos = total + rank;
      float* o = p.out_det + ((int64_t)b * p.A + pos) * 6;
      const float4 bx = box[j];
      o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
      o[4] = pred[(int64_t)a * (4 + p.nc) + 4 + cls];
      o[5] = (float)cls;
      p.out_idx[(int64_t)b * p.A + pos] = a;
    }
    total += found;
  }
  if (tid == 0) {
    p.out_cnt[b] = total;
    if (p.out_cand) p.out_cand[b] = n;
  }
}

// padded keys per image: a power of two >= 32 (so that a warp of the prepare kernel never straddles two images)
int next_pow2(int v) {
  int r = 32;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace
}  // namespace dcfa

extern "C" int64_t dcfa_nms_workspace_bytes(int B, int A) {
  if (B <= 0 || A <= 0) return 0;
  const int64_t apad = dcfa::next_pow2(A);
  const int64_t key_bytes = ((int64_t)B * apad * 8 + 255) / 256 * 256;
  return key_bytes + (int64_t)B * A * 16 + 256 + (((int64_t)B * 4 + 255) / 256) * 256;
}

extern "C" int dcfa_nms(float* pred, int B, int A, int nc, float conf_thres, double nms_thres, int iou_mode,
                        float* out_det, int32_t* out_idx, int32_t* out_cnt, int32_t* out_cand, void* workspace,
                        int64_t workspace_bytes, void* stream) {
  using namespace dcfa;
  cudaStream_t st = (cudaStream_t)stream;
  DCFA_REQUIRE(pred && out_det && out_idx && out_cnt && workspace, "nms: null pointer");
  DCFA_REQUIRE(B > 0 && A > 0 && nc > 0, "nms: bad sizes B=%d A=%d nc=%d", B, A, nc);
  DCFA_REQUIRE(nc <= 256, "nms: nc %d > 256 unsupported", nc);
  DCFA_REQUIRE(A <= 65536, "nms: A %d > 65536 unsupported", A);
  DCFA_REQUIRE(iou_mode == DCFA_IOU_TV_CPU || iou_mode == DCFA_IOU_TV_CUDA, "nms: bad iou_mode %d", iou_mode);
  DCFA_REQUIRE(workspace_bytes >= dcfa_nms_workspace_bytes(B, A), "nms: workspace too small");
  DCFA_REQUIRE(((uintptr_t)workspace % 16) == 0, "nms: workspace must be 16-byte aligned");
  const int Apad = next_pow2(A);
  uint64_t* keys = reinterpret_cast<uint64_t*>(workspace);
  const int64_t key_bytes = ((int64_t)B * Apad * 8 + 255) / 256 * 256;
  float4* sbox = reinterpret_cast<float4*>(reinterpret_cast<char*>(workspace) + key_bytes);

  int32_t* cand = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(sbox) + (((int64_t)B * A * 16 + 255) / 256) * 256);
  cudaError_t me = cudaMemsetAsync(cand, 0, (size_t)B * 4, st);
  if (me != cudaSuccess) return fail(DCFA_E_CUDA, "nms: cudaMemsetAsync: %s", cudaGetErrorString(me));
  PrepArgs pa{pred, keys, cand, B, A, Apad, nc, conf_thres};
  const int64_t tot = (int64_t)B * Apad;
  nms_prepare_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pa);
  DCFA_CHECK_LAUNCH("nms_prepare_kernel");

  if (Apad > kSmemSortMax) {
    SortArgs sa{keys, cand, Apad};
    nms_sort_global_kernel<<<(unsigned)B, kNmsThreads, 0, st>>>(sa);
    DCFA_CHECK_LAUNCH("nms_sort_global_kernel");
  }

  // cluster size: the largest power of two <= 8 that still gives every image its own SMs
  int CS = 8;
  while (CS > 1 && (int64_t)B * CS > sm_count()) CS >>= 1;
  // shared-memory carve-out: greedy phase state[A16] + 17 bytes per resident candidate; sort phase up to
  // 16384 / CS (rounded up to a power of two >= 1024) padded keys
  const int A16 = (A + 15) / 16 * 16;
  const int kNmsSmemMax = 200 * 1024;
  int cap = (kNmsSmemMax - A16) / 17;
  cap = cap >= A16 ? A16 : cap / 16 * 16;
  size_t smem = (size_t)A16 + (size_t)cap * 17;
  {
    int per_cta = (std::min(Apad, kSmemSortMax) + CS - 1) / CS;
    int nloc = 1024;
    while (nloc < per_cta) nloc <<= 1;
    smem = std::max(smem, (size_t)(nloc + nloc / 16) * 8);
  }
  GreedyArgs ga{pred, keys, sbox, out_det, out_idx, out_cnt, out_cand, cand, B, A, Apad, nc, cap, (float)nms_thres, nms_thres};
  static DeviceOnce attr_done;
  if (attr_done.needed()) {
    cudaError_t e1 = cudaFuncSetAttribute(nms_cluster_kernel<DCFA_IOU_TV_CPU>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNmsSmemMax);
    cudaError_t e2 = cudaFuncSetAttribute(nms_cluster_kernel<DCFA_IOU_TV_CUDA>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNmsSmemMax);
    if (e1 != cudaSuccess || e2 != cudaSuccess)
      return fail(DCFA_E_CUDA, "nms: cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    attr_done.mark();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CS));
  cfg.blockDim = dim3(kNmsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = iou_mode == DCFA_IOU_TV_CUDA ? cudaLaunchKernelEx(&cfg, nms_cluster_kernel<DCFA_IOU_TV_CUDA>, ga)
                                                : cudaLaunchKernelEx(&cfg, nms_cluster_kernel<DCFA_IOU_TV_CPU>, ga);
  if (le != cudaSuccess) return fail(DCFA_E_CUDA, "nms: cluster launch (B=%d, cluster %d): %s", B, CS, cudaGetErrorString(le));
  count_launch();
  return DCFA_OK;
}
