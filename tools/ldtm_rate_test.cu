// Experiment (not part of the product): throughput of tcgen05.ld (TMEM -> registers) per SM on B200, for the shapes an
// epilogue can use (32x32b.x16 / .x32 / .x64) and 4 / 8 / 16 reading warps, with and without tcgen05.mma traffic into
// another accumulator stage.  Motivation: the conv epilogue spends ~6 000 cycles per 128 x 256 fp32 accumulator tile
// whatever the number of epilogue warps; is the TMEM read port the limit?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I dcfa-yolo_b200/csrc -o tools/_bin/ldtm_rate tools/ldtm_rate_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace dcfa;

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t& sink) {
  if constexpr (X == 16) {
    uint32_t r[16];
    ptx::tmem_ld_x16(taddr, r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) sink ^= r[i];
  } else if constexpr (X == 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= r[i];
  }
}

// two x16 loads in flight before one wait (what the conv epilogue does)
__device__ __forceinline__ void ld2x16(uint32_t t0, uint32_t t1, uint32_t& sink) {
  uint32_t a[16], b[16];
  ptx::tmem_ld_x16(t0, a);
  ptx::tmem_ld_x16(t1, b);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) sink ^= a[i] ^ b[i];
}

// mode 0: x16 + wait; 1: two x16 + wait; 2: x32 + wait.  mma: warp (nwarps) issues M128 x N256 x K16 MMAs into columns 256..511
__global__ void __launch_bounds__(1024) k_rate(int mode, int reps, int mma, long long* cycles, uint32_t* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gb = raw + (base - ptx::smem_u32(raw));
  const uint32_t s_a = base, s_b = base + 16384, bar = base + 16384 + 32768, slot = bar + 16;
  const int tid = threadIdx.x, warp = tid >> 5, nread = (blockDim.x >> 5) - 1;
  for (int i = tid; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gb)[i] = 0;
  if (warp == 0) {
    if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
    __syncwarp();
    ptx::tmem_alloc(slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(gb + 16384 + 32768 + 16);
  uint32_t sink = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nread) {
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int grp = warp >> 2, ngrp = nread >> 2;   // groups share the 256 columns chunk-interleaved
    __syncwarp();
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (mode == 0) {
        for (int j = grp; j < 16; j += ngrp) ld<16>(lane_base + j * 16, sink);
      } else if (mode == 1) {
        for (int j = grp; j < 16; j += 2 * ngrp) ld2x16(lane_base + j * 16, lane_base + ((j + ngrp) & 15) * 16, sink);
      } else {
        for (int j = grp; j < 8; j += ngrp) ld<32>(lane_base + j * 32, sink);
      }
    }
    t1 = clock64();
  } else if (mma && (tid & 31) == 0) {
    // SWIZZLE_128B K-major operands of zeros: only the issue / accumulator traffic matters
    const uint32_t idesc = ptx::make_idesc_bf16_f32(128, 256);
    const uint64_t hi = ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    ptx::fence_proxy_async_smem();
    for (int r = 0; r < reps * 6; ++r) {
      for (int k = 0; k < 4; ++k)
        ptx::umma_bf16(tmem + 256, hi | (uint64_t)(((s_a & 0x3FFFFu) >> 4) + 2 * k), hi | (uint64_t)(((s_b & 0x3FFFFu) >> 4) + 2 * k), idesc, 1u);
    }
    ptx::umma_commit(bar);
    ptx::mbar_wait(bar, 0);
  }
  if (warp < nread && (tid & 31) == 0) { cycles[blockIdx.x * 32 + warp] = t1 - t0; }
  out[blockIdx.x * blockDim.x + tid] = sink;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d_cyc; uint32_t* d_out;
  cudaMalloc(&d_cyc, 148 * 32 * 8); cudaMalloc(&d_out, 148 * 1024 * 4);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 200;
  const char* names[3] = {"x16 + wait", "2 x16 + wait", "x32 + wait"};
  for (int mma = 0; mma < 2; ++mma)
    for (int mode = 0; mode < 3; ++mode)
      for (int nread = 4; nread <= 16; nread *= 2) {
        cudaMemset(d_cyc, 0, 148 * 32 * 8);
        k_rate<<<148, (nread + 1) * 32, 60 * 1024>>>(mode, reps, mma, d_cyc, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[32];
        cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < nread; ++i) mx = h[i] > mx ? h[i] : mx;
        const double bytes = (double)reps * 128 * 256 * 4;   // every rep reads the whole 128 x 256 fp32 tile once
        printf("mma %d  %-14s %2d warps: %8lld cycles for %d tiles -> %6.1f cycles per 128x256 tile, %6.1f B/cycle/SM\n", mma,
               names[mode], nread, mx, reps, (double)mx / reps, bytes / (double)mx);
      }
  return 0;
}
