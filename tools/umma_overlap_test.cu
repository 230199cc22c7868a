// Experiment (not part of the product): can a NO-SWIZZLE K-major tcgen05 shared-memory descriptor describe
// OVERLAPPING rows, i.e. row n of the operand = 16 consecutive bf16 starting 8 elements after row n-1?
//   canonical no-swizzle K-major layout: core matrix = 8 rows x 16 bytes stored contiguously (row pitch 16 B),
//   LBO = byte distance between core matrices adjacent in K, SBO = between 8-row groups.
//   With LBO = 16 and SBO = 128 the address of (row n, K chunk j) is base + 16 n + 16 j: row n holds elements
//   flat[8n .. 8n+15] -- the im2col of a [y][x][4 channel] bf16 image patch (pixel pairs per row, four pixels of K)
//   without ever materialising it.  This is what the stem kernel needs (dcfa-yolo_b200/csrc/stem.cu).
// The test computes D[m][n] = sum_k A[m][k] * flat[8 n + off + k] for three accumulating MMAs (off = 0, 288, 576
// elements... one per kernel row) and compares with the host, for the four (LBO,SBO) interpretations of A and B.
// It also times a loop of six N=128 K=16 MMAs (the stem's per-tile tensor work) with clock64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I dcfa-yolo_b200/csrc -o tools/_bin/umma_overlap tools/umma_overlap_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace dcfa;

constexpr int NFLAT = 2048;          // bf16 elements of the flat "patch" (4 KB)
constexpr int ROWOFF = 144;          // elements between kernel rows (36 pixels x 4 channels)
constexpr int M = 128, N = 128, K = 16;

__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}

// a: [3][128][16] bf16 row-major (three weight tiles); flat: [NFLAT]; d: [128][N] fp32
// avar: 0 -> A stored [rowgroup][kchunk][8][8] (LBO 128, SBO 256); 1 -> descriptor fields swapped
// bvar: 0 -> B (LBO 16, SBO 128); 1 -> swapped
__global__ void __launch_bounds__(128) k_test(const __nv_bfloat16* a, const __nv_bfloat16* flat, float* d, int avar, int bvar,
                                              int reps, long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gb = raw + (base - ptx::smem_u32(raw));
  const uint32_t s_a = base, s_b = base + 3 * 4096, bar = s_b + NFLAT * 2 + 1024, slot = bar + 8;
  const int tid = threadIdx.x;
  // A tiles, canonical no-swizzle: element (m, k) of tile t at t*4096 + (m/8)*256 + (k/8)*128 + (m%8)*16 + (k%8)*2
  for (int i = tid; i < 3 * M * K; i += 128) {
    const int t = i / (M * K), m = (i / K) % M, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(gb + t * 4096 + (m / 8) * 256 + (k / 8) * 128 + (m % 8) * 16 + (k % 8) * 2) = a[i];
  }
  for (int i = tid; i < NFLAT; i += 128) *reinterpret_cast<__nv_bfloat16*>(gb + 3 * 4096 + i * 2) = flat[i];
  if (tid < 32) {
    if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
    __syncwarp();
    ptx::tmem_alloc(slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(gb + 3 * 4096 + NFLAT * 2 + 1024 + 8);
  const uint32_t idesc = ptx::make_idesc_bf16_f32(M, N);
  if (tid == 0) {
    ptx::fence_proxy_async_smem();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int ky = 0; ky < 3; ++ky) {
        const uint64_t ad = avar == 0 ? desc_nosw(s_a + ky * 4096, 128, 256) : desc_nosw(s_a + ky * 4096, 256, 128);
        const uint64_t bd = bvar == 0 ? desc_nosw(s_b + ky * ROWOFF * 2, 16, 128) : desc_nosw(s_b + ky * ROWOFF * 2, 128, 16);
        ptx::umma_bf16(tmem, ad, bd, idesc, ky ? 1u : 0u);
        if (reps > 1) ptx::umma_bf16(tmem + 128, ad, bd, idesc, ky ? 1u : 0u);   // second accumulator (the other pixel parity)
      }
    }
    ptx::umma_commit(bar);
    ptx::mbar_wait(bar, 0);
    if (cycles) *cycles = clock64() - t0;
  }
  __syncthreads();
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    uint32_t acc[16];
    ptx::tmem_ld_x16(tmem + (uint32_t)c + ((uint32_t)((tid >> 5) * 32) << 16), acc);
    ptx::tmem_ld_wait();
    for (int n = 0; n < 16; ++n) d[tid * N + c + n] = __uint_as_float(acc[n]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, 256);
}

// three-input max (PTX ISA 8.6+, sm_100): does it assemble and give the right answer?
__global__ void k_max3(const float* x, float* y) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(x[0]), "f"(x[1]), "f"(x[2]));
  y[0] = r;
}

int main() {
  __nv_bfloat16* ha = (__nv_bfloat16*)malloc(3 * M * K * 2);
  __nv_bfloat16* hf = (__nv_bfloat16*)malloc(NFLAT * 2);
  float* fa = (float*)malloc(3 * M * K * 4);
  float* ff = (float*)malloc(NFLAT * 4);
  srand(3);
  for (int i = 0; i < 3 * M * K; ++i) { float v = (float)(rand() % 17 - 8) / 8.0f; ha[i] = __float2bfloat16(v); fa[i] = __bfloat162float(ha[i]); }
  for (int i = 0; i < NFLAT; ++i) { float v = (float)(rand() % 13 - 6) / 4.0f; hf[i] = __float2bfloat16(v); ff[i] = __bfloat162float(hf[i]); }
  __nv_bfloat16 *da, *df; float* dd; long long* dc;
  cudaMalloc(&da, 3 * M * K * 2); cudaMalloc(&df, NFLAT * 2); cudaMalloc(&dd, M * N * 4); cudaMalloc(&dc, 8);
  cudaMemcpy(da, ha, 3 * M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(df, hf, NFLAT * 2, cudaMemcpyHostToDevice);
  const int smem = 1024 + 3 * 4096 + NFLAT * 2 + 1024 + 64;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float* hd = (float*)malloc(M * N * 4);
  int rc = 1;
  for (int avar = 0; avar < 2; ++avar)
    for (int bvar = 0; bvar < 2; ++bvar) {
      k_test<<<1, 128, smem>>>(da, df, dd, avar, bvar, 1, nullptr);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("avar %d bvar %d: CUDA error %s\n", avar, bvar, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hd, dd, M * N * 4, cudaMemcpyDeviceToHost);
      int bad = 0; double maxerr = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int ky = 0; ky < 3; ++ky)
            for (int k = 0; k < K; ++k) ref += (double)fa[(ky * M + m) * K + k] * ff[8 * n + ky * ROWOFF + k];
          double err = fabs(ref - hd[m * N + n]);
          if (err > 1e-3) ++bad;
          if (err > maxerr) maxerr = err;
        }
      printf("A %s, B %s : %s (mismatches %d / %d, max err %.3g)\n", avar ? "(LBO 256,SBO 128)" : "(LBO 128,SBO 256)",
             bvar ? "(LBO 128,SBO 16)" : "(LBO 16,SBO 128) overlapping rows", bad ? "WRONG" : "ok", bad, M * N, maxerr);
      if (!bad && avar == 0 && bvar == 0) rc = 0;
    }
  for (int reps : {100, 1000}) {
    k_test<<<1, 128, smem>>>(da, df, dd, 0, 0, reps, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("timing: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("timing: %d x 6 MMAs (M128 N128 K16, no-swizzle operands): %.1f cycles per group of 6\n", reps, (double)c / reps);
  }
  float hx[3] = {1.5f, -2.0f, 7.25f}, hy = 0; float *dx, *dy;
  cudaMalloc(&dx, 12); cudaMalloc(&dy, 4); cudaMemcpy(dx, hx, 12, cudaMemcpyHostToDevice);
  k_max3<<<1, 1>>>(dx, dy); cudaDeviceSynchronize(); cudaMemcpy(&hy, dy, 4, cudaMemcpyDeviceToHost);
  printf("max3: %g (expect 7.25)\n", hy);
  return rc;
}
