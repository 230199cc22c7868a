// Shared device/host helpers for the dcfa_b200 kernel library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <type_traits>
#include <utility>
#include <stdint.h>
#include <string.h>

#include "../../include/dcfa_b200.h"

namespace dcfa {

// ---------------------------------------------------------------------------------------------
// error plumbing (api.cu owns the storage)
// ---------------------------------------------------------------------------------------------
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();        // SM count of the CURRENT device (cached per device ordinal)
int current_device();  // cudaGetDevice, -1 on failure

// One-time, PER-DEVICE setup guard (cudaFuncSetAttribute is a per-device property: a process-wide `static bool` would leave
// the second GPU of a multi-device process without its shared-memory opt-in).  Use as a function-local static.
struct DeviceOnce {
  unsigned char done[64] = {};
  bool needed() const { const int d = current_device(); return d < 0 || d >= 64 || !done[d]; }
  void mark() { const int d = current_device(); if (d >= 0 && d < 64) done[d] = 1; }
};

#define DCFA_CHECK_LAUNCH(name)                                                         \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return dcfa::fail(DCFA_E_CUDA, "%s launch failed: %s", name, \
                                              cudaGetErrorString(e__));                 \
    dcfa::count_launch();                                                               \
  } while (0)

#define DCFA_REQUIRE(cond, ...)                                        \
  do {                                                                 \
    if (!(cond)) return dcfa::fail(DCFA_E_INVALID, __VA_ARGS__);       \
  } while (0)

// A resolved activation view (device pointer + strides in elements).
template <typename T>
struct View {
  T* p;
  int64_t img_stride;
  int64_t gstride;
  int ld;
  int gi;
  __host__ __device__ __forceinline__ int64_t img_off(int n) const {
    if (gi <= 0) return (int64_t)n * img_stride;
    int g = n / gi;
    return (int64_t)(n - g * gi) * img_stride + (int64_t)g * gstride;
  }
};

template <typename T>
static inline View<T> resolve(const dcfa_view& v, void* const* bufs) {
  View<T> r;
  r.p = (v.buf < 0) ? nullptr : reinterpret_cast<T*>(static_cast<char*>(bufs[v.buf]) + v.off);
  r.img_stride = v.img_stride;
  r.gstride = v.gstride;
  r.ld = v.ld;
  r.gi = v.gi;
  return r;
}

template <typename T>
static inline T* resolve_ptr(const dcfa_view& v, void* const* bufs) {
  return (v.buf < 0) ? nullptr : reinterpret_cast<T*>(static_cast<char*>(bufs[v.buf]) + v.off);
}

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// SiLU with ONE special-function op: x * sigmoid(x) = h + h * tanh(h), h = x / 2.  The conv epilogues are bound by
// the MUFU pipe (16 results / cycle / SM): exp + reciprocal cost two ops per output, tanh.approx.f32 one.
// |error| <= |x| * 2^-12 (tanh.approx.f32: relative error 2^-11), below the bf16 rounding of the stored result
// for x > -4 and below 1.3e-3 absolute down to x = -5.5, where SiLU itself is < 0.023 in magnitude.
__device__ __forceinline__ float silu_fast(float x) {
#ifdef DCFA_EXACT_SILU
  return x * sigmoid_fast(x);
#else
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
#endif
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == DCFA_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == DCFA_ACT_SILU) return silu_fast(v);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}

// 8 bf16 <-> 8 floats through one 128-bit word
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

// Packed fp32 pairs: sm_100 executes fma.rn.f32x2 (two IEEE fp32 FMAs, one instruction) -- the depthwise kernel is
// bound by instruction issue, and 9 of its ~18 instructions per output are FMAs.
struct F2 {
  unsigned long long v;
};
__device__ __forceinline__ F2 f2_make(float lo, float hi) {
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_get(const F2& a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ void f2_fma(F2& acc, const F2& a, const F2& b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.v) : "l"(a.v), "l"(b.v));
}
// 8 bf16 -> 4 fp32 pairs (element 2i in the low half of word i: shift; element 2i+1: mask)
__device__ __forceinline__ void unpack8_f2(const uint4& q, F2* v) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = f2_make(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
}

__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg128(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// kernel launches: immediate, or RECORDED into a plan (dcfa_plan_create) and replayed by dcfa_plan_run
// ---------------------------------------------------------------------------------------------
// One prepared launch: function, geometry, attributes and a private copy of every argument (tensor maps included), so
// that replaying it costs one cudaLaunchKernelExC and nothing else -- no tensor-map encoding, no getenv, no attribute calls.
struct LaunchRecord {
  const void* func = nullptr;
  dim3 grid, block;
  size_t smem = 0;
  int cluster = 0;      // cluster dimension x (0: no cluster attribute)
  int pdl = 1;          // programmatic stream serialization allowed
  int nargs = 0;
  uint32_t off[8] = {};
  alignas(64) unsigned char blob[1216];
};
struct Recorder {
  LaunchRecord* recs;
  int cap, n;
  bool overflow;
};
Recorder*& recorder();   // thread-local; non-null while dcfa_plan_* prepares a dispatch unit (launches are recorded, not issued)
cudaError_t launch_record(const LaunchRecord& r, cudaStream_t st);
bool pdl_enabled();

template <typename T>
inline void pack_arg(LaunchRecord& r, uint32_t& pos, const T& v) {
  pos = (pos + 63u) & ~63u;
  static_assert(std::is_trivially_copyable<T>::value, "kernel arguments must be trivially copyable");
  if (pos + sizeof(T) <= sizeof(r.blob) && r.nargs < 8) {
    memcpy(r.blob + pos, &v, sizeof(T));
    r.off[r.nargs++] = pos;
  } else {
    r.nargs = 99;   // flagged below
  }
  pos += (uint32_t)sizeof(T);
}

// EVERY kernel of the forward path is launched through this helper.  With pdl the kernel carries the programmatic-
// dependent-launch attribute (see ptx::pdl_wait) and must execute griddepcontrol.wait before its first global access that
// depends on earlier work; DCFA_PDL=0 in the environment falls back to plain stream-ordered launches (debugging).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, bool pdl,
                            Args&&... args) {
  static_assert(sizeof...(KArgs) == sizeof...(Args), "argument count mismatch");
  LaunchRecord r;
  r.func = reinterpret_cast<const void*>(kernel);
  r.grid = grid; r.block = block; r.smem = smem; r.cluster = cluster; r.pdl = pdl ? 1 : 0;
  uint32_t pos = 0;
  (pack_arg<typename std::remove_cv<typename std::remove_reference<KArgs>::type>::type>(r, pos, args), ...);
  if (r.nargs > 8) return cudaErrorInvalidValue;
  if (Recorder* rec = recorder()) {
    if (rec->n < rec->cap) rec->recs[rec->n++] = r;
    else rec->overflow = true;
    return cudaSuccess;
  }
  return launch_record(r, st);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return launch_k(kernel, grid, block, smem, st, 0, true, std::forward<Args>(args)...);
}

// launchers implemented in the individual .cu files (host side; bufs already resolved by the caller)
int launch_stem(const dcfa_op& op, void* const* bufs, cudaStream_t st);
int launch_conv(const dcfa_op& op, void* const* bufs, cudaStream_t st);
int launch_conv_tma(const dcfa_op& op, void* const* bufs, cudaStream_t st);
int launch_conv_strip(const dcfa_op& op, void* const* bufs, cudaStream_t st, bool* taken);   // 3x3 stride-1 halo-strip path
int launch_dwconv(const dcfa_op& op, void* const* bufs, cudaStream_t st);
int launch_chain(const dcfa_op& pw1, const dcfa_op& dw, const dcfa_op& pw2, void* const* bufs, cudaStream_t st);
int launch_ghost(const dcfa_op& pw, const dcfa_op& dw, void* const* bufs, cudaStream_t st);
int launch_cbam_pool(const dcfa_op& op, void* const* bufs, cudaStream_t st);
int launch_cbam_mlp(const dcfa_op& op, void* const* bufs, cudaStream_t st);
int launch_cbam_stats(const dcfa_op& op, void* const* bufs, cudaStream_t st);
int launch_cbam_apply(const dcfa_op& op, void* const* bufs, cudaStream_t st);
int launch_cbam_fused(const dcfa_op& pool, const dcfa_op& mlp, const dcfa_op& stats, const dcfa_op& apply, void* const* bufs,
                      cudaStream_t st);
int launch_maxpool5(const dcfa_op& op, void* const* bufs, cudaStream_t st);
int launch_sppf_fused(const dcfa_op* ops, void* const* bufs, cudaStream_t st);   // 19 records, see cbam.cu
int launch_upsample(const dcfa_op& op, void* const* bufs, cudaStream_t st);
int launch_dfl(const dcfa_op& op, void* const* bufs, cudaStream_t st);

}  // namespace dcfa
