// C-ABI front door: error plumbing, device check, and the flat op-list executor (dcfa_run_ops).
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include <stdlib.h>

#include "common.cuh"

namespace dcfa {

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
int g_sms[64] = {};
}  // namespace

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DCFA_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int current_device() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int sm_count() {
  const int dev = current_device();
  if (dev < 0 || dev >= 64) return 148;
  if (g_sms[dev] == 0) {
    int n = 0;
    g_sms[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return g_sms[dev];
}

}  // namespace dcfa

extern "C" {

int dcfa_abi_version(void) { return DCFA_ABI_VERSION; }
int dcfa_sizeof_view(void) { return (int)sizeof(dcfa_view); }
int dcfa_sizeof_op(void) { return (int)sizeof(dcfa_op); }
const char* dcfa_last_error(void) { return dcfa::g_err; }
int64_t dcfa_launch_count(void) { return dcfa::g_launches.load(std::memory_order_relaxed); }

int dcfa_device_check(int dev) {
  int major = 0, minor = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) return dcfa::fail(DCFA_E_CUDA, "device_check: %s", cudaGetErrorString(e));
  if (major != 10) return dcfa::fail(DCFA_E_ARCH, "device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
  return DCFA_OK;
}

int dcfa_run_ops(const dcfa_op* ops, int n_ops, void* const* bufs, int nbufs, void* stream) {
  using namespace dcfa;
  DCFA_REQUIRE(ops && n_ops >= 0 && bufs && nbufs > 0, "run_ops: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < n_ops; ++i) {
    const dcfa_op& op = ops[i];
    const dcfa_view* vs[9] = {&op.x, &op.x2, &op.y, &op.w, &op.scale, &op.bias, &op.a0, &op.a1, &op.a2};
    for (int k = 0; k < 9; ++k)
      if (vs[k]->buf >= nbufs || (vs[k]->buf >= 0 && bufs[vs[k]->buf] == nullptr))
        return fail(DCFA_E_INVALID, "run_ops: op %d (kind %d) view %d references buffer %d (nbufs %d)", i, op.kind, k,
                    vs[k]->buf, nbufs);
    int rc;
    // peephole: the four records of one CBAM run as a single cluster launch when the shape allows
    if (op.kind == DCFA_OP_CBAM_POOL && i + 3 < n_ops && ops[i + 1].kind == DCFA_OP_CBAM_MLP &&
        ops[i + 2].kind == DCFA_OP_CBAM_STATS && ops[i + 3].kind == DCFA_OP_CBAM_APPLY) {
      bool ok = true;
      for (int j = 1; j <= 3 && ok; ++j) {
        const dcfa_op& o = ops[i + j];
        const dcfa_view* ws[9] = {&o.x, &o.x2, &o.y, &o.w, &o.scale, &o.bias, &o.a0, &o.a1, &o.a2};
        for (int k = 0; k < 9; ++k)
          if (ws[k]->buf >= nbufs || (ws[k]->buf >= 0 && bufs[ws[k]->buf] == nullptr)) ok = false;
      }
      if (ok) {
        rc = launch_cbam_fused(op, ops[i + 1], ops[i + 2], ops[i + 3], bufs, st);
        if (rc < 0) {
          char tmp[400];
          snprintf(tmp, sizeof(tmp), "%s", dcfa_last_error());
          return fail(rc, "op %d (fused cbam): %s", i, tmp);
        }
        if (rc == 1) { i += 3; continue; }
      }
    }
    // peephole: 1x1 conv -> depthwise -> 1x1 conv chains the plan marked as private run as one fused kernel
    if (op.kind == DCFA_OP_CONV && (op.flags & DCFA_CONV_FLAG_CHAIN_HEAD) && i + 2 < n_ops && ops[i + 1].kind == DCFA_OP_DWCONV &&
        ops[i + 2].kind == DCFA_OP_CONV) {
      bool ok = true;
      for (int j = 1; j <= 2 && ok; ++j) {
        const dcfa_op& o = ops[i + j];
        const dcfa_view* ws[9] = {&o.x, &o.x2, &o.y, &o.w, &o.scale, &o.bias, &o.a0, &o.a1, &o.a2};
        for (int k = 0; k < 9; ++k)
          if (ws[k]->buf >= nbufs || (ws[k]->buf >= 0 && bufs[ws[k]->buf] == nullptr)) ok = false;
      }
      if (ok) {
        rc = launch_chain(op, ops[i + 1], ops[i + 2], bufs, st);
        if (rc < 0) {
          char tmp[400];
          snprintf(tmp, sizeof(tmp), "%s", dcfa_last_error());
          return fail(rc, "op %d (fused chain): %s", i, tmp);
        }
        if (rc == 1) { i += 2; continue; }
      }
    }
    switch (op.kind) {
      case DCFA_OP_STEM: rc = launch_stem(op, bufs, st); break;
      case DCFA_OP_CONV: rc = launch_conv(op, bufs, st); break;
      case DCFA_OP_DWCONV: rc = launch_dwconv(op, bufs, st); break;
      case DCFA_OP_CBAM_POOL: rc = launch_cbam_pool(op, bufs, st); break;
      case DCFA_OP_CBAM_MLP: rc = launch_cbam_mlp(op, bufs, st); break;
      case DCFA_OP_CBAM_STATS: rc = launch_cbam_stats(op, bufs, st); break;
      case DCFA_OP_CBAM_APPLY: rc = launch_cbam_apply(op, bufs, st); break;
      case DCFA_OP_MAXPOOL5: rc = launch_maxpool5(op, bufs, st); break;
      case DCFA_OP_UPSAMPLE: rc = launch_upsample(op, bufs, st); break;
      case DCFA_OP_DFL: rc = launch_dfl(op, bufs, st); break;
      default: return fail(DCFA_E_INVALID, "run_ops: op %d has unknown kind %d", i, op.kind);
    }
    if (rc != DCFA_OK) {
      char tmp[400];
      snprintf(tmp, sizeof(tmp), "%s", dcfa_last_error());
      return fail(rc, "op %d: %s", i, tmp);
    }
  }
  return DCFA_OK;
}

}  // extern "C"
