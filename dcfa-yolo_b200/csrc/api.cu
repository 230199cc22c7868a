// C-ABI front door: error plumbing, device check, and the flat op-list executor (dcfa_run_ops).
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <new>
#include <vector>

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

Recorder*& recorder() {
  static thread_local Recorder* r = nullptr;
  return r;
}

// launches that are only being recorded are counted when the plan replays them
void count_launch(int n) {
  if (!recorder()) g_launches.fetch_add(n, std::memory_order_relaxed);
}

cudaError_t launch_record(const LaunchRecord& r, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = r.grid;
  cfg.blockDim = r.block;
  cfg.dynamicSmemBytes = r.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned na = 0;
  if (r.cluster > 0) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)r.cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (r.pdl && pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  void* argv[8];
  for (int i = 0; i < r.nargs; ++i) argv[i] = const_cast<unsigned char*>(r.blob) + r.off[i];
  return cudaLaunchKernelExC(&cfg, r.func, argv);
}

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

}  // extern "C"

namespace dcfa {
namespace {

bool views_ok(const dcfa_op& o, void* const* bufs, int nbufs) {
  const dcfa_view* vs[9] = {&o.x, &o.x2, &o.y, &o.w, &o.scale, &o.bias, &o.a0, &o.a1, &o.a2};
  for (int k = 0; k < 9; ++k)
    if (vs[k]->buf >= nbufs || (vs[k]->buf >= 0 && bufs[vs[k]->buf] == nullptr)) return false;
  return true;
}

// Number of consecutive records the dispatcher treats as one unit at index i (the fusion candidates), 1 otherwise.
bool is_cbam(const dcfa_op* ops, int i, int n_ops) {
  return i + 3 < n_ops && ops[i].kind == DCFA_OP_CBAM_POOL && ops[i + 1].kind == DCFA_OP_CBAM_MLP &&
         ops[i + 2].kind == DCFA_OP_CBAM_STATS && ops[i + 3].kind == DCFA_OP_CBAM_APPLY;
}

constexpr int kSppfSpan = 19;   // CBAM, (MAXPOOL5, CBAM) x 3: SPPF_CBAM's attention / pooling sequence

int unit_span(const dcfa_op* ops, int i, int n_ops) {
  const dcfa_op& op = ops[i];
  if (is_cbam(ops, i, n_ops) && i + kSppfSpan <= n_ops) {
    bool sppf = true;
    for (int s = 1; s < 4 && sppf; ++s)
      sppf = ops[i + 5 * s - 1].kind == DCFA_OP_MAXPOOL5 && is_cbam(ops, i + 5 * s, n_ops);
    if (sppf) return kSppfSpan;
  }
  if (op.kind == DCFA_OP_CBAM_POOL && i + 3 < n_ops && ops[i + 1].kind == DCFA_OP_CBAM_MLP &&
      ops[i + 2].kind == DCFA_OP_CBAM_STATS && ops[i + 3].kind == DCFA_OP_CBAM_APPLY)
    return 4;
  if (op.kind == DCFA_OP_CONV && (op.flags & DCFA_CONV_FLAG_CHAIN_HEAD) && i + 2 < n_ops && ops[i + 1].kind == DCFA_OP_DWCONV &&
      ops[i + 2].kind == DCFA_OP_CONV)
    return 3;
  if (op.kind == DCFA_OP_CONV && (op.flags & DCFA_CONV_FLAG_GHOST_HEAD) && i + 1 < n_ops && ops[i + 1].kind == DCFA_OP_DWCONV)
    return 2;
  return 1;
}

int run_single(const dcfa_op& op, int i, void* const* bufs, cudaStream_t st) {
  int rc;
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
  return DCFA_OK;
}

// Executes (or, under a Recorder, records) the dispatch unit that starts at op i: one op, or the fused form of a CBAM
// quadruple / a private 1x1 -> depthwise -> 1x1 chain when the shapes allow it (else its records one by one).
int dispatch_unit(const dcfa_op* ops, int i, int span, void* const* bufs, int nbufs, cudaStream_t st) {
  for (int j = 0; j < span; ++j)
    if (!views_ok(ops[i + j], bufs, nbufs))
      return fail(DCFA_E_INVALID, "run_ops: op %d (kind %d) references a missing buffer (nbufs %d)", i + j, ops[i + j].kind, nbufs);
  if (span == kSppfSpan) {
    const int rc = launch_sppf_fused(ops + i, bufs, st);
    if (rc < 0) {
      char tmp[400];
      snprintf(tmp, sizeof(tmp), "%s", dcfa_last_error());
      return fail(rc, "op %d (fused SPPF_CBAM): %s", i, tmp);
    }
    if (rc == 1) return DCFA_OK;
    for (int s = 0; s < 4; ++s) {   // its units one by one
      if (s > 0) {
        const int r1 = dispatch_unit(ops, i + 5 * s - 1, 1, bufs, nbufs, st);
        if (r1 != DCFA_OK) return r1;
      }
      const int r4 = dispatch_unit(ops, i + 5 * s, 4, bufs, nbufs, st);
      if (r4 != DCFA_OK) return r4;
    }
    return DCFA_OK;
  }
  if (span == 4) {
    const int rc = launch_cbam_fused(ops[i], ops[i + 1], ops[i + 2], ops[i + 3], bufs, st);
    if (rc < 0) {
      char tmp[400];
      snprintf(tmp, sizeof(tmp), "%s", dcfa_last_error());
      return fail(rc, "op %d (fused cbam): %s", i, tmp);
    }
    if (rc == 1) return DCFA_OK;
  } else if (span == 3) {
    const int rc = launch_chain(ops[i], ops[i + 1], ops[i + 2], bufs, st);
    if (rc < 0) {
      char tmp[400];
      snprintf(tmp, sizeof(tmp), "%s", dcfa_last_error());
      return fail(rc, "op %d (fused chain): %s", i, tmp);
    }
    if (rc == 1) return DCFA_OK;
  } else if (span == 2) {
    const int rc = launch_ghost(ops[i], ops[i + 1], bufs, st);
    if (rc < 0) {
      char tmp[400];
      snprintf(tmp, sizeof(tmp), "%s", dcfa_last_error());
      return fail(rc, "op %d (fused ghost module): %s", i, tmp);
    }
    if (rc == 1) return DCFA_OK;
  }
  for (int j = 0; j < span; ++j) {
    const int rc = run_single(ops[i + j], i + j, bufs, st);
    if (rc != DCFA_OK) return rc;
  }
  return DCFA_OK;
}

}  // namespace
}  // namespace dcfa

// ---------------------------------------------------------------------------------------------------------------------
// plan object: the op list prepared ONCE (kernel variants chosen, tensor maps encoded, launch geometry fixed), replayed
// with one cudaLaunchKernelExC per kernel
// ---------------------------------------------------------------------------------------------------------------------
struct dcfa_plan {
  std::vector<dcfa_op> ops;
  std::vector<void*> bound;      // per buffer index: pointer fixed at creation, or nullptr = supplied at every run
  std::vector<void*> last;       // pointers the current records were prepared with
  struct Unit {
    int first, span;
    bool dynamic;                // references a buffer that is supplied per run
    bool prepared;
    std::vector<dcfa::LaunchRecord> recs;
  };
  std::vector<Unit> units;
  int device = -1;
  // dcfa_plan_load only: library-owned device memory and the model facts stored in the plan file
  void* own_blob = nullptr;
  void* own_arena = nullptr;
  dcfa_plan_info info = {};
};

namespace dcfa {
namespace {

bool unit_uses_dynamic(const dcfa_plan& p, const dcfa_plan::Unit& u) {
  for (int j = 0; j < u.span; ++j) {
    const dcfa_op& o = p.ops[u.first + j];
    const dcfa_view* vs[9] = {&o.x, &o.x2, &o.y, &o.w, &o.scale, &o.bias, &o.a0, &o.a1, &o.a2};
    for (int k = 0; k < 9; ++k)
      if (vs[k]->buf >= 0 && vs[k]->buf < (int)p.bound.size() && p.bound[vs[k]->buf] == nullptr) return true;
  }
  return false;
}

bool unit_pointers_changed(const dcfa_plan& p, const dcfa_plan::Unit& u, void* const* bufs) {
  for (int j = 0; j < u.span; ++j) {
    const dcfa_op& o = p.ops[u.first + j];
    const dcfa_view* vs[9] = {&o.x, &o.x2, &o.y, &o.w, &o.scale, &o.bias, &o.a0, &o.a1, &o.a2};
    for (int k = 0; k < 9; ++k)
      if (vs[k]->buf >= 0 && bufs[vs[k]->buf] != p.last[vs[k]->buf]) return true;
  }
  return false;
}

int prepare_unit(dcfa_plan& p, dcfa_plan::Unit& u, void* const* bufs) {
  LaunchRecord tmp[24];   // the largest unit (SPPF_CBAM, 19 records) run as separate kernels
  Recorder rec{tmp, 24, 0, false};
  recorder() = &rec;
  const int rc = dispatch_unit(p.ops.data(), u.first, u.span, bufs, (int)p.bound.size(), nullptr);
  recorder() = nullptr;
  if (rc != DCFA_OK) return rc;
  if (rec.overflow || rec.n == 0) return fail(DCFA_E_INVALID, "plan: op %d produced %d launches", u.first, rec.n);
  u.recs.assign(tmp, tmp + rec.n);
  u.prepared = true;
  return DCFA_OK;
}

}  // namespace
}  // namespace dcfa

extern "C" {

int dcfa_run_ops(const dcfa_op* ops, int n_ops, void* const* bufs, int nbufs, void* stream) {
  using namespace dcfa;
  DCFA_REQUIRE(ops && n_ops >= 0 && bufs && nbufs > 0, "run_ops: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < n_ops;) {
    const int span = unit_span(ops, i, n_ops);
    const int rc = dispatch_unit(ops, i, span, bufs, nbufs, st);
    if (rc != DCFA_OK) return rc;
    i += span;
  }
  return DCFA_OK;
}

int dcfa_plan_create(const dcfa_op* ops, int n_ops, void* const* bufs, int nbufs, dcfa_plan** out) {
  using namespace dcfa;
  DCFA_REQUIRE(ops && n_ops > 0 && bufs && nbufs > 0 && out, "plan_create: bad arguments");
  dcfa_plan* p = new (std::nothrow) dcfa_plan();
  DCFA_REQUIRE(p != nullptr, "plan_create: out of memory");
  p->ops.assign(ops, ops + n_ops);
  p->bound.assign(bufs, bufs + nbufs);
  p->last.assign(nbufs, nullptr);
  p->device = current_device();
  for (int i = 0; i < n_ops;) {
    dcfa_plan::Unit u;
    u.first = i;
    u.span = unit_span(ops, i, n_ops);
    u.prepared = false;
    u.dynamic = false;
    p->units.push_back(u);
    i += u.span;
  }
  for (auto& u : p->units) {
    u.dynamic = unit_uses_dynamic(*p, u);
    if (!u.dynamic) {   // everything this unit touches is bound: prepare it now
      const int rc = prepare_unit(*p, u, p->bound.data());
      if (rc != DCFA_OK) { delete p; return rc; }
    }
  }
  for (int b = 0; b < nbufs; ++b) p->last[b] = p->bound[b];
  *out = p;
  return DCFA_OK;
}

int dcfa_plan_run(dcfa_plan* p, void* const* bufs, int nbufs, void* stream) {
  using namespace dcfa;
  DCFA_REQUIRE(p && bufs && nbufs == (int)p->bound.size(), "plan_run: bad arguments");
  DCFA_REQUIRE(current_device() == p->device, "plan_run: the plan was created on device %d, the current device is %d", p->device,
               current_device());
  for (int b = 0; b < nbufs; ++b)
    DCFA_REQUIRE(p->bound[b] == nullptr || bufs[b] == p->bound[b] || bufs[b] == nullptr, "plan_run: buffer %d is bound to another pointer", b);
  std::vector<void*> cur(p->bound);
  for (int b = 0; b < nbufs; ++b)
    if (cur[b] == nullptr) cur[b] = bufs[b];
  cudaStream_t st = (cudaStream_t)stream;
  for (auto& u : p->units) {
    if (!u.prepared || (u.dynamic && unit_pointers_changed(*p, u, cur.data()))) {
      const int rc = prepare_unit(*p, u, cur.data());
      if (rc != DCFA_OK) return rc;
    }
  }
  p->last = cur;
  for (auto& u : p->units)
    for (const auto& r : u.recs) {
      cudaError_t e = launch_record(r, st);
      if (e != cudaSuccess) return fail(DCFA_E_CUDA, "plan_run: op %d launch failed: %s", u.first, cudaGetErrorString(e));
      count_launch();
    }
  return DCFA_OK;
}

int dcfa_plan_num_launches(const dcfa_plan* p) {
  if (!p) return 0;
  int n = 0;
  for (const auto& u : p->units) n += (int)u.recs.size();
  return n;
}

void dcfa_plan_destroy(dcfa_plan* p) {
  if (!p) return;
  if (p->own_blob) cudaFree(p->own_blob);
  if (p->own_arena) cudaFree(p->own_arena);
  delete p;
}

// ---- plan files (written by dcfa_b200.plan.Plan.save): lets a caller with no Python run the forward pass
int dcfa_plan_load(const char* path, dcfa_plan** out) {
  using namespace dcfa;
  DCFA_REQUIRE(path && out, "plan_load: bad arguments");
  FILE* f = fopen(path, "rb");
  DCFA_REQUIRE(f != nullptr, "plan_load: cannot open %s", path);
  dcfa_plan_file_header h;
  std::vector<dcfa_op> ops;
  std::vector<char> blob;
  bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "DCFAPLN1", 8) == 0;
  if (ok) ok = h.abi_version == DCFA_ABI_VERSION && h.sizeof_op == (int32_t)sizeof(dcfa_op) && h.n_ops > 0 && h.blob_bytes >= 0 &&
               h.arena_bytes > 0 && h.nbufs == DCFA_NUM_BUFS;
  if (ok) {
    ops.resize(h.n_ops);
    blob.resize((size_t)h.blob_bytes);
    ok = fread(ops.data(), sizeof(dcfa_op), h.n_ops, f) == (size_t)h.n_ops &&
         (h.blob_bytes == 0 || fread(blob.data(), 1, blob.size(), f) == blob.size());
  }
  fclose(f);
  DCFA_REQUIRE(ok, "plan_load: %s is not a plan file of this library version", path);
  void *dblob = nullptr, *darena = nullptr;
  cudaError_t e = cudaMalloc(&dblob, blob.size() + 256);
  if (e == cudaSuccess) e = cudaMalloc(&darena, (size_t)h.arena_bytes + 256);
  if (e == cudaSuccess) e = cudaMemcpy(dblob, blob.data(), blob.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(dblob); cudaFree(darena);
    return fail(DCFA_E_CUDA, "plan_load: %s", cudaGetErrorString(e));
  }
  void* bufs[DCFA_NUM_BUFS] = {};
  bufs[DCFA_BUF_BLOB] = dblob;
  bufs[DCFA_BUF_ARENA] = darena;
  dcfa_plan* p = nullptr;
  const int rc = dcfa_plan_create(ops.data(), h.n_ops, bufs, DCFA_NUM_BUFS, &p);
  if (rc != DCFA_OK) { cudaFree(dblob); cudaFree(darena); return rc; }
  p->own_blob = dblob;
  p->own_arena = darena;
  p->info = h.info;
  *out = p;
  return DCFA_OK;
}

int dcfa_plan_get_info(const dcfa_plan* p, dcfa_plan_info* info) {
  using namespace dcfa;
  DCFA_REQUIRE(p && info, "plan_get_info: bad arguments");
  *info = p->info;
  return DCFA_OK;
}

int dcfa_plan_forward(dcfa_plan* p, const void* rgb, const void* depth, float* x0, float* x1, float* x2, float* dbox, float* cls,
                      void* stream) {
  void* bufs[DCFA_NUM_BUFS] = {};
  bufs[DCFA_BUF_RGB] = const_cast<void*>(rgb);
  bufs[DCFA_BUF_NIR] = const_cast<void*>(depth);
  bufs[DCFA_BUF_X0] = x0; bufs[DCFA_BUF_X1] = x1; bufs[DCFA_BUF_X2] = x2;
  bufs[DCFA_BUF_DBOX] = dbox; bufs[DCFA_BUF_CLS] = cls;
  return dcfa_plan_run(p, bufs, DCFA_NUM_BUFS, stream);
}

}  // extern "C"
