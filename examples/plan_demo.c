/*
 * A caller WITHOUT Python: runs YoloBody.forward (reference nets/yolo_mul.py:397-462) through the C ABI alone.
 *
 *   plan_demo <plan file> <rgb.bin> <depth.bin> <out.bin>
 *
 * The plan file is written once by the host-side compiler (`dcfa_b200.plan.Plan(...).save(path)`); the inputs are raw
 * tensors in the layout dcfa_plan_get_info reports (fp32 [B,3,H,W], or uint8 [B,H,W,3] / [B,H,W]); out.bin receives
 * dbox [B,4,A] followed by cls [B,nc,A] (fp32).  Built by dcfa-yolo_b200/csrc/Makefile (target demo) with plain gcc:
 * only include/dcfa_b200.h, libdcfa_b200.so and the CUDA runtime are needed.
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "dcfa_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define DK(x) do { int r_ = (x); if (r_ != DCFA_OK) { fprintf(stderr, "%s: %d %s\n", #x, r_, dcfa_last_error()); return 3; } } while (0)

static void* read_file(const char* path, size_t want) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  void* p = malloc(want);
  size_t got = fread(p, 1, want, f);
  fclose(f);
  if (got != want) { free(p); return NULL; }
  return p;
}

int main(int argc, char** argv) {
  if (argc != 5) { fprintf(stderr, "usage: %s plan rgb.bin depth.bin out.bin\n", argv[0]); return 1; }
  DK(dcfa_device_check(0));
  dcfa_plan* plan = NULL;
  DK(dcfa_plan_load(argv[1], &plan));
  dcfa_plan_info in;
  DK(dcfa_plan_get_info(plan, &in));
  const size_t px = (size_t)in.batch * in.height * in.width;
  const size_t rgb_bytes = in.input_u8 ? px * 3 : px * 3 * sizeof(float);
  const size_t dep_bytes = in.input_u8 ? (in.depth_plane ? px : px * 3) : px * 3 * sizeof(float);
  void* h_rgb = read_file(argv[2], rgb_bytes);
  void* h_dep = read_file(argv[3], dep_bytes);
  if (!h_rgb || !h_dep) { fprintf(stderr, "input files do not match the plan (%zu / %zu bytes expected)\n", rgb_bytes, dep_bytes); return 1; }
  void *d_rgb, *d_dep;
  float *x[3], *dbox, *cls;
  CK(cudaMalloc(&d_rgb, rgb_bytes));
  CK(cudaMalloc(&d_dep, dep_bytes));
  for (int l = 0; l < 3; ++l) CK(cudaMalloc((void**)&x[l], sizeof(float) * in.batch * in.no * in.level_hw[l][0] * in.level_hw[l][1]));
  const size_t nb = (size_t)in.batch * 4 * in.anchors, nc = (size_t)in.batch * in.num_classes * in.anchors;
  CK(cudaMalloc((void**)&dbox, nb * sizeof(float)));
  CK(cudaMalloc((void**)&cls, nc * sizeof(float)));
  CK(cudaMemcpy(d_rgb, h_rgb, rgb_bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_dep, h_dep, dep_bytes, cudaMemcpyHostToDevice));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  for (int it = 0; it < 3; ++it) DK(dcfa_plan_forward(plan, d_rgb, d_dep, x[0], x[1], x[2], dbox, cls, st));   /* replays: no re-preparation */
  CK(cudaStreamSynchronize(st));
  float* out = (float*)malloc((nb + nc) * sizeof(float));
  CK(cudaMemcpy(out, dbox, nb * sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out + nb, cls, nc * sizeof(float), cudaMemcpyDeviceToHost));
  FILE* f = fopen(argv[4], "wb");
  if (!f || fwrite(out, sizeof(float), nb + nc, f) != nb + nc) { fprintf(stderr, "cannot write %s\n", argv[4]); return 1; }
  fclose(f);
  printf("plan_demo: B=%d %dx%d nc=%d A=%d, %d kernel launches per forward, %lld launches total\n", in.batch, in.height, in.width,
         in.num_classes, in.anchors, dcfa_plan_num_launches(plan), (long long)dcfa_launch_count());
  dcfa_plan_destroy(plan);
  return 0;
}
