// Experiment (not part of the product): per-SM throughput of the special-function and conversion instructions the conv
// epilogue leans on -- tanh.approx.f32 (MUFU.TANH), ex2.approx.f32, rcp.approx.f32, cvt.rn.bf16x2.f32 (F2FP), fma.rn.f32x2
// (FFMA2) -- measured with 32 warps per SM, 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_bin/mufu_rate tools/mufu_rate_test.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, int iters, long long* cycles) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + 1) + 0.1f * i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(v[i])); v[i] = __uint_as_float(r << 16); }
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  float* d; long long* c; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&c, 8);
  const char* names[5] = {"tanh.approx.f32", "ex2.approx.f32", "rcp.approx.f32", "cvt.rn.bf16x2.f32", "fma.rn.f32"};
  const int iters = 2000;
  for (int op = 0; op < 5; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      if (op == 0) k<0><<<148, 1024>>>(d, iters, c);
      if (op == 1) k<1><<<148, 1024>>>(d, iters, c);
      if (op == 2) k<2><<<148, 1024>>>(d, iters, c);
      if (op == 3) k<3><<<148, 1024>>>(d, iters, c);
      if (op == 4) k<4><<<148, 1024>>>(d, iters, c);
      cudaDeviceSynchronize();
    }
    long long cyc; cudaMemcpy(&cyc, c, 8, cudaMemcpyDeviceToHost);
    const double ops = 1024.0 * 8 * iters;
    printf("%-20s %.2f results / cycle / SM\n", names[op], ops / (double)cyc);
  }
  return 0;
}
