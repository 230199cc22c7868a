// Experiment (not part of the product): can a tcgen05 shared-memory descriptor start r ROWS into a SWIZZLE_128B
// K-major tile (start address + r * 128 bytes), i.e. can the nine taps of a 3x3 conv be served from one halo tile?
// A is a 320-row x 64-col bf16 tile laid out exactly as TMA would write it (chunk ^= row & 7, 1024-byte aligned base);
// D[m][n] = sum_k A[r + m][k] * B[n][k] is computed for several r with the descriptor's base_offset field set to
// mode 0: 0, mode 1: r & 7, and compared with the host result.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I dcfa-yolo_b200/csrc -o /tmp/umma_rowoffset tools/umma_rowoffset_test.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"

using namespace dcfa;

constexpr int AROWS = 320, K = 64, N = 16;

__global__ void __launch_bounds__(128) k_test(const __nv_bfloat16* a, const __nv_bfloat16* b, float* d, int r, int mode) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gb = raw + (base - ptx::smem_u32(raw));
  const uint32_t s_a = base, s_b = base + AROWS * 128, bar = s_b + 2048, slot = bar + 8;
  const int tid = threadIdx.x;
  for (int i = tid; i < AROWS * 8; i += 128) {   // 16-byte chunks, swizzled by absolute row
    const int row = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(gb + row * 128 + ((c ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(a + row * K + c * 8);
  }
  for (int i = tid; i < N * 8; i += 128) {
    const int row = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(gb + AROWS * 128 + row * 128 + ((c ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(b + row * K + c * 8);
  }
  if (tid < 32) {
    if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
    __syncwarp();
    ptx::tmem_alloc(slot, 32);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<uint32_t*>(gb + AROWS * 128 + 2048 + 8);
  if (tid == 0) {
    ptx::fence_proxy_async_smem();
    const uint64_t hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint32_t a_addr = s_a + (uint32_t)r * 128u;
    const uint64_t boff = mode == 1 ? ((uint64_t)((a_addr >> 7) & 7u) << 49) : 0ull;
    const uint32_t idesc = ptx::make_idesc_bf16_f32(128, N);
    for (int k = 0; k < 4; ++k)
      ptx::umma_bf16(tmem, hi | boff | (uint64_t)(((a_addr + k * 32) & 0x3FFFFu) >> 4), hi | (uint64_t)(((s_b + k * 32) & 0x3FFFFu) >> 4), idesc,
                     k ? 1u : 0u);
    ptx::umma_commit(bar);
  }
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  uint32_t acc[16];
  ptx::tmem_ld_x16(tmem + ((uint32_t)((tid >> 5) * 32) << 16), acc);
  ptx::tmem_ld_wait();
  for (int n = 0; n < N; ++n) d[tid * N + n] = __uint_as_float(acc[n]);
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, 32);
}

int main() {
  __nv_bfloat16 *ha = (__nv_bfloat16*)malloc(AROWS * K * 2), *hb = (__nv_bfloat16*)malloc(N * K * 2);
  float* fa = (float*)malloc(AROWS * K * 4); float* fb = (float*)malloc(N * K * 4);
  srand(1);
  for (int i = 0; i < AROWS * K; ++i) { float v = (float)(rand() % 17 - 8) / 8.0f; ha[i] = __float2bfloat16(v); fa[i] = __bfloat162float(ha[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (float)(rand() % 13 - 6) / 4.0f; hb[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *da, *db; float* dd;
  cudaMalloc(&da, AROWS * K * 2); cudaMalloc(&db, N * K * 2); cudaMalloc(&dd, 128 * N * 4);
  cudaMemcpy(da, ha, AROWS * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, N * K * 2, cudaMemcpyHostToDevice);
  const int smem = 1024 + AROWS * 128 + 2048 + 64;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float hd[128 * N];
  const int rs[] = {0, 8, 1, 2, 3, 7, 9, 17, 18, 34, 36, 37};
  for (int mode = 0; mode < 2; ++mode)
    for (int r : rs) {
      k_test<<<1, 128, smem>>>(da, db, dd, r, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d r %d: CUDA error %s\n", mode, r, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hd, dd, sizeof(hd), cudaMemcpyDeviceToHost);
      int bad = 0; double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)fa[(r + m) * K + k] * fb[n * K + k];
          double err = fabs(ref - hd[m * N + n]);
          if (err > 1e-3) ++bad;
          if (err > maxerr) maxerr = err;
        }
      printf("base_offset mode %d  r = %2d : %s (mismatches %d / %d, max err %.3g)\n", mode, r, bad ? "WRONG" : "ok", bad, 128 * N, maxerr);
    }
  return 0;
}
