// Experiment (not part of the product): issue cost of tcgen05.mma M128 x N x K16 (bf16, no-swizzle K-major operands)
// for N = 16 .. 256, with the A operand in the "channel-plane" layout the fused 1x1 -> depthwise kernel wants:
// element (pixel m, channel k) at  (k / 8) * PLANE + m * 16 + (k % 8) * 2   (LBO = PLANE, SBO = 128), so that a
// spatial tap is just a start-address offset of 16 bytes per pixel.  Also checks the result of a depthwise 3x3 written as
// nine accumulating MMAs against diagonal B tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I dcfa-yolo_b200/csrc -o tools/_bin/umma_nsweep tools/umma_nsweep_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace dcfa;

constexpr int NPIX = 320;            // pixels in the plane buffer (M = 128 rows + tap shifts up to 2 * 18 + 2)
constexpr int PLANE = NPIX * 16;     // bytes per 8-channel plane
constexpr int HWID = 18;             // halo row width: tap (dy, dx) = shift dy * 18 + dx pixels

__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}

// x: [NPIX][16] bf16 (16 channels), w: [9][16] bf16 taps; d: [128][16] fp32 = depthwise 3x3 at pixel m (top-left anchored)
__global__ void __launch_bounds__(128) k_dw(const __nv_bfloat16* x, const __nv_bfloat16* w, float* d, int n_cols, int reps,
                                            long long* cycles) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gb = raw + (base - ptx::smem_u32(raw));
  // A planes: 2 x PLANE; B tiles: 9 x [n_cols rows x 16 k] canonical (LBO 128, SBO 256); barrier; slot
  const uint32_t s_a = base, s_b = base + 2 * PLANE, b_tile = (uint32_t)n_cols * 32u;
  const uint32_t bar = s_b + 9 * 256 * 32, slot = bar + 8;
  const int tid = threadIdx.x;
  for (int i = tid; i < NPIX * 16; i += 128) {
    const int m = i / 16, k = i % 16;
    *reinterpret_cast<__nv_bfloat16*>(gb + (k / 8) * PLANE + m * 16 + (k % 8) * 2) = x[i];
  }
  for (int i = tid; i < 9 * n_cols * 16; i += 128) {   // B[t][n][k] = (n == k) ? w[t][n] : 0   (n >= 16: zero)
    const int t = i / (n_cols * 16), n = (i / 16) % n_cols, k = i % 16;
    const __nv_bfloat16 v = (n == k) ? w[t * 16 + n] : __float2bfloat16(0.0f);
    *reinterpret_cast<__nv_bfloat16*>(gb + 2 * PLANE + t * b_tile + (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) = v;
  }
  if (tid < 32) {
    if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
    __syncwarp();
    ptx::tmem_alloc(slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(gb + 2 * PLANE + 9 * 256 * 32 + 8);
  const uint32_t idesc = ptx::make_idesc_bf16_f32(128, n_cols);
  if (tid == 0) {
    ptx::fence_proxy_async_smem();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int t = 0; t < 9; ++t) {
        const int shift = (t / 3) * HWID + (t % 3);
        ptx::umma_bf16(tmem + (uint32_t)((r & 1) * 256), desc_nosw(s_a + shift * 16, PLANE, 128), desc_nosw(s_b + t * b_tile, 128, 256), idesc, t ? 1u : 0u);
      }
    ptx::umma_commit(bar);
    ptx::mbar_wait(bar, 0);
    if (cycles) *cycles = clock64() - t0;
  }
  __syncthreads();
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  {
    uint32_t acc[16];
    ptx::tmem_ld_x16(tmem + ((uint32_t)((tid >> 5) * 32) << 16), acc);
    ptx::tmem_ld_wait();
    for (int n = 0; n < 16; ++n) d[tid * 16 + n] = __uint_as_float(acc[n]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  __nv_bfloat16* hx = (__nv_bfloat16*)malloc(NPIX * 16 * 2);
  __nv_bfloat16* hw = (__nv_bfloat16*)malloc(9 * 16 * 2);
  float* fx = (float*)malloc(NPIX * 16 * 4);
  float fw[9 * 16];
  srand(5);
  for (int i = 0; i < NPIX * 16; ++i) { float v = (float)(rand() % 17 - 8) / 8.0f; hx[i] = __float2bfloat16(v); fx[i] = v; }
  for (int i = 0; i < 9 * 16; ++i) { float v = (float)(rand() % 13 - 6) / 4.0f; hw[i] = __float2bfloat16(v); fw[i] = v; }
  __nv_bfloat16 *dx, *dw; float* dd; long long* dc;
  cudaMalloc(&dx, NPIX * 16 * 2); cudaMalloc(&dw, 9 * 16 * 2); cudaMalloc(&dd, 128 * 16 * 4); cudaMalloc(&dc, 8);
  cudaMemcpy(dx, hx, NPIX * 16 * 2, cudaMemcpyHostToDevice); cudaMemcpy(dw, hw, 9 * 16 * 2, cudaMemcpyHostToDevice);
  const int smem = 1024 + 2 * PLANE + 9 * 256 * 32 + 64;
  cudaFuncSetAttribute(k_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float hd[128 * 16];
  int rc = 0;
  k_dw<<<1, 128, smem>>>(dx, dw, dd, 16, 1, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(hd, dd, sizeof(hd), cudaMemcpyDeviceToHost);
  int bad = 0; double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int c = 0; c < 16; ++c) {
      double ref = 0;
      for (int t = 0; t < 9; ++t) ref += (double)fw[t * 16 + c] * fx[(m + (t / 3) * HWID + (t % 3)) * 16 + c];
      const double err = fabs(ref - hd[m * 16 + c]);
      if (err > 1e-3) ++bad;
      if (err > maxerr) maxerr = err;
    }
  printf("depthwise 3x3 as 9 diagonal MMAs (A planes LBO %d SBO 128, shifts of 16 B per pixel): %s (mismatches %d, max err %.3g)\n", PLANE,
         bad ? "WRONG" : "ok", bad, maxerr);
  if (bad) rc = 1;
  for (int n : {16, 32, 64, 128, 256}) {
    k_dw<<<1, 128, smem>>>(dx, dw, dd, n, 400, dc);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("timing N=%d: CUDA error %s\n", n, cudaGetErrorString(e)); return 1; }
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("timing: M128 N%-3d K16: %.1f cycles per MMA\n", n, (double)c / (400 * 9));
  }
  return rc;
}
