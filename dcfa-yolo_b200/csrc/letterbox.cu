// dcfa_letterbox_u8: resize_image (utils/utils.py:24-37) on the device -- PIL's `image.resize((nw, nh), Image.BICUBIC)` pasted
// on a grey (128) canvas -- bit-exact for 8-bit images.
//
// Pillow resamples 8-bit images in FIXED POINT (libImaging/Resample.c): per output coordinate a window [xmin, xmin + xmax) of
// source pixels and double-precision bicubic weights (a = -0.5, support 2 * max(scale, 1): antialiased when shrinking),
// normalised to sum 1 and rounded to integers with 22 fractional bits; the horizontal pass runs first over the source rows the
// vertical pass needs and rounds to uint8, then the vertical pass does the same.  A pass is skipped when that axis keeps its
// size.  The three kernels below restate exactly that: the weights are computed on the device with explicitly rounded
// double operations (no FMA contraction), the passes are integer arithmetic.
// dcfa_pack_detections: DecodeBox.yolo_correct_boxes (utils/utils_bbox.py:60-85, :170-173) on the device, fused with the
// packing of (count, first K rows) per image that precedes the single device->host copy of a batch.
#include "common.cuh"

namespace dcfa {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c: PRECISION_BITS

__device__ __forceinline__ double bicubic(double x) {   // Resample.c: bicubic_filter, a = -0.5
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dadd_rn(__dmul_rn(a + 2.0, x), -(a + 3.0)), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(x, -5.0), x), 8.0), x), -4.0), a);
  return 0.0;
}

// Resample.c: precompute_coeffs + normalize_coeffs_8bpc for one axis; thread = output coordinate
__global__ void coeff_kernel(int in_size, int out_size, int ksize, int32_t* __restrict__ bounds, int32_t* __restrict__ kk) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out_size) return;
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);   // in0 = 0
  const double ss = __ddiv_rn(1.0, filterscale);
  int xmin = (int)__dadd_rn(__dadd_rn(center, -support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x)
    ww = __dadd_rn(ww, bicubic(__dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss)));
  int32_t* k = kk + (int64_t)xx * ksize;
  for (int x = 0; x < ksize; ++x) {
    double w = 0.0;
    if (x < xmax) {
      w = bicubic(__dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss));
      if (ww != 0.0) w = __ddiv_rn(w, ww);
    }
    const double f = __dmul_rn(w, (double)(1 << kPrecisionBits));
    k[x] = w < 0.0 ? (int32_t)__dadd_rn(-0.5, f) : (int32_t)__dadd_rn(0.5, f);
  }
  bounds[2 * xx] = xmin;
  bounds[2 * xx + 1] = xmax;
}

__device__ __forceinline__ uint8_t clip8(int32_t v) {   // Resample.c: clip8 (lookup of v >> PRECISION_BITS, clamped)
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

struct PassArgs {
  const uint8_t* src;   // [src_h][src_w][C]
  uint8_t* dst;         // [dst_h][dst_w][C] (the canvas for the last pass)
  const int32_t* bounds;
  const int32_t* kk;
  int ksize, C;
  int src_w, src_h;     // source extent
  int out_w, out_h;     // extent of the region this pass writes
  int dst_w;            // destination row pitch in pixels
  int top, left;        // position of the region inside the destination
  int row0;             // horizontal pass: first source row (ybox_first); vertical pass: subtracted from bounds[2*y]
};

// horizontal pass: out(y, x) = clip8(sum_k src(row0 + y, xmin + k) * kk[x][k]); thread = (x, y), all channels
__global__ void resample_h_kernel(const PassArgs p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= p.out_w) return;
  const int xmin = p.bounds[2 * x], xmax = p.bounds[2 * x + 1];
  const int32_t* k = p.kk + (int64_t)x * p.ksize;
  const uint8_t* s = p.src + ((int64_t)(p.row0 + y) * p.src_w + xmin) * p.C;
  uint8_t* d = p.dst + ((int64_t)(p.top + y) * p.dst_w + p.left + x) * p.C;
  for (int c = 0; c < p.C; ++c) {
    int32_t acc = 1 << (kPrecisionBits - 1);
    for (int i = 0; i < xmax; ++i) acc += (int32_t)s[i * p.C + c] * k[i];
    d[c] = clip8(acc);
  }
}

// vertical pass: out(y, x) = clip8(sum_k src(ymin + k - row0, x) * kk[y][k])
__global__ void resample_v_kernel(const PassArgs p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= p.out_w) return;
  const int ymin = p.bounds[2 * y] - p.row0, ymax = p.bounds[2 * y + 1];
  const int32_t* k = p.kk + (int64_t)y * p.ksize;
  const uint8_t* s = p.src + ((int64_t)ymin * p.src_w + x) * p.C;
  uint8_t* d = p.dst + ((int64_t)(p.top + y) * p.dst_w + p.left + x) * p.C;
  const int64_t pitch = (int64_t)p.src_w * p.C;
  for (int c = 0; c < p.C; ++c) {
    int32_t acc = 1 << (kPrecisionBits - 1);
    for (int i = 0; i < ymax; ++i) acc += (int32_t)s[i * pitch + c] * k[i];
    d[c] = clip8(acc);
  }
}

// canvas fill (grey 128 outside the pasted region) or plain copy (no pass needed)
__global__ void canvas_kernel(uint8_t* dst, int H, int W, int C, int top, int left, int nh, int nw, const uint8_t* copy_src) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const bool inside = y >= top && y < top + nh && x >= left && x < left + nw;
  uint8_t* d = dst + ((int64_t)y * W + x) * C;
  if (!inside) {
    for (int c = 0; c < C; ++c) d[c] = 128;
  } else if (copy_src) {
    const uint8_t* s = copy_src + ((int64_t)(y - top) * nw + (x - left)) * C;
    for (int c = 0; c < C; ++c) d[c] = s[c];
  }
}

int ksize_for(int in_size, int out_size) {   // Resample.c: ksize = (int)ceil(support) * 2 + 1
  double scale = (double)in_size / out_size;
  if (scale < 1.0) scale = 1.0;
  const double support = 2.0 * scale;
  int c = (int)support;
  if ((double)c < support) ++c;
  return c * 2 + 1;
}

// yolo_correct_boxes in the reference's dtype flow (utils/utils_bbox.py:60-85 called from :170-173): the detections are
// float32, the shapes int64 -> every intermediate is float64 except `box_hw *= scale` (in place: rounded to float32)
// and the final assignment into the float32 rows.
__global__ void pack_detections_kernel(const float* __restrict__ det, const int32_t* __restrict__ cnt, int B, int A, int K,
                                       const int32_t* __restrict__ image_hw, int in_h, int in_w, int letterbox,
                                       float* __restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float* o = out + (int64_t)b * (1 + 6 * K);
  const int n = cnt[b];
  if (i == 0) o[0] = (float)n;
  if (i >= K) return;
  float* r = o + 1 + 6 * i;
  if (i >= n) {
    for (int c = 0; c < 6; ++c) r[c] = 0.0f;
    return;
  }
  const float* d = det + ((int64_t)b * A + i) * 6;
  if (!image_hw) {
    for (int c = 0; c < 6; ++c) r[c] = d[c];
    return;
  }
  const float x1 = d[0], y1 = d[1], x2 = d[2], y2 = d[3];
  // box_xy = (d[0:2] + d[2:4]) / 2, box_wh = d[2:4] - d[0:2]   (float32, utils/utils_bbox.py:172)
  const float cxy[2] = {__fdiv_rn(__fadd_rn(y1, y2), 2.0f), __fdiv_rn(__fadd_rn(x1, x2), 2.0f)};   // (y, x) order (:63)
  float hw[2] = {__fsub_rn(y2, y1), __fsub_rn(x2, x1)};
  const double in_shape[2] = {(double)in_h, (double)in_w};
  const double img[2] = {(double)image_hw[2 * b], (double)image_hw[2 * b + 1]};
  double yx[2] = {(double)cxy[0], (double)cxy[1]};
  if (letterbox) {
    const double r0 = __ddiv_rn(in_shape[0], img[0]), r1 = __ddiv_rn(in_shape[1], img[1]);
    const double m = r0 < r1 ? r0 : r1;
    for (int k = 0; k < 2; ++k) {
      const double new_shape = rint(__dmul_rn(img[k], m));                                   // np.round: half to even
      const double offset = __ddiv_rn(__ddiv_rn(__dadd_rn(in_shape[k], -new_shape), 2.0), in_shape[k]);
      const double scale = __ddiv_rn(in_shape[k], new_shape);
      yx[k] = __dmul_rn(__dadd_rn(yx[k], -offset), scale);
      hw[k] = (float)__dmul_rn((double)hw[k], scale);                                        // in-place float32 *= float64
    }
  }
  float box[4];
  for (int k = 0; k < 2; ++k) {
    if (letterbox) {   // box_yx is float64 here: mins / maxes are float64
      const double half = __ddiv_rn((double)hw[k], 2.0);
      box[k] = (float)__dmul_rn(__dadd_rn(yx[k], -half), img[k]);
      box[2 + k] = (float)__dmul_rn(__dadd_rn(yx[k], half), img[k]);
    } else {           // everything stayed float32 up to the final `boxes *= image_shape` (computed in float64, stored as float32)
      const float half = __fdiv_rn(hw[k], 2.0f);
      box[k] = (float)__dmul_rn((double)__fsub_rn(cxy[k], half), img[k]);
      box[2 + k] = (float)__dmul_rn((double)__fadd_rn(cxy[k], half), img[k]);
    }
  }
  r[0] = box[0]; r[1] = box[1]; r[2] = box[2]; r[3] = box[3]; r[4] = d[4]; r[5] = d[5];
}

}  // namespace
}  // namespace dcfa

extern "C" {

int64_t dcfa_letterbox_workspace_bytes(int src_h, int src_w, int channels, int out_h, int out_w) {
  if (src_h <= 0 || src_w <= 0 || out_h <= 0 || out_w <= 0 || channels <= 0) return 0;
  // worst case: resample to the full canvas; tables for both axes + the horizontally resampled temporary
  const int64_t kh = dcfa::ksize_for(src_w, out_w), kv = dcfa::ksize_for(src_h, out_h);
  const int64_t tables = ((int64_t)out_w * (2 + kh) + (int64_t)out_h * (2 + kv)) * 4;
  return ((tables + 255) / 256 + 1) * 256 + (int64_t)src_h * out_w * channels + 256;
}

int dcfa_letterbox_u8(const uint8_t* src, int src_h, int src_w, int channels, uint8_t* dst, int out_h, int out_w,
                      int letterbox, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace dcfa;
  DCFA_REQUIRE(src && dst && workspace, "letterbox: missing tensor");
  DCFA_REQUIRE(src_h > 0 && src_w > 0 && out_h > 0 && out_w > 0 && (channels == 1 || channels == 3), "letterbox: bad shape");
  DCFA_REQUIRE(workspace_bytes >= dcfa_letterbox_workspace_bytes(src_h, src_w, channels, out_h, out_w), "letterbox: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  // resize_image (utils/utils.py:24-37): scale = min(w/iw, h/ih); nw = int(iw*scale); nh = int(ih*scale); paste centred
  int nw = out_w, nh = out_h;
  if (letterbox) {
    const double sw = (double)out_w / src_w, sh = (double)out_h / src_h;
    const double scale = sw < sh ? sw : sh;
    nw = (int)(src_w * scale);
    nh = (int)(src_h * scale);
    DCFA_REQUIRE(nw > 0 && nh > 0, "letterbox: image %dx%d collapses to %dx%d", src_w, src_h, nw, nh);
  }
  const int left = (out_w - nw) / 2, top = (out_h - nh) / 2;
  const bool need_h = nw != src_w, need_v = nh != src_h;
  const int kh = ksize_for(src_w, nw), kv = ksize_for(src_h, nh);
  char* ws = static_cast<char*>(workspace);
  int32_t* bounds_h = reinterpret_cast<int32_t*>(ws);
  int32_t* kk_h = bounds_h + 2 * (int64_t)nw;
  int32_t* bounds_v = kk_h + (int64_t)nw * kh;
  int32_t* kk_v = bounds_v + 2 * (int64_t)nh;
  const int64_t tables = ((int64_t)nw * (2 + kh) + (int64_t)nh * (2 + kv)) * 4;
  uint8_t* tmp = reinterpret_cast<uint8_t*>(ws + ((tables + 255) / 256 + 1) * 256);

  const dim3 blk(128);
  // grey canvas (and the pasted region itself when no pass is needed: Image.resize returns a copy)
  canvas_kernel<<<dim3(ceil_div(out_w, 128), out_h), blk, 0, st>>>(dst, out_h, out_w, channels, top, left, nh, nw,
                                                                  (!need_h && !need_v) ? src : nullptr);
  DCFA_CHECK_LAUNCH("canvas_kernel");
  if (need_h) {
    coeff_kernel<<<ceil_div(nw, 128), blk, 0, st>>>(src_w, nw, kh, bounds_h, kk_h);
    DCFA_CHECK_LAUNCH("coeff_kernel");
  }
  if (need_v) {
    coeff_kernel<<<ceil_div(nh, 128), blk, 0, st>>>(src_h, nh, kv, bounds_v, kk_v);
    DCFA_CHECK_LAUNCH("coeff_kernel");
  }
  // Resample.c ImagingResampleInner: the horizontal pass covers the source rows [ybox_first, ybox_last) the vertical pass
  // reads; those bounds depend only on (src_h, nh) and are recomputed here on the host in the same double arithmetic
  int ybox_first = 0, ybox_last = src_h;
  if (need_v) {
    const double scale = (double)src_h / nh, fs = scale < 1.0 ? 1.0 : scale, support = 2.0 * fs;
    auto lo = [&](int yy) { int v = (int)(((yy + 0.5) * scale) - support + 0.5); return v < 0 ? 0 : v; };
    auto hi = [&](int yy) { int v = (int)(((yy + 0.5) * scale) + support + 0.5); return v > src_h ? src_h : v; };
    ybox_first = lo(0);
    ybox_last = hi(nh - 1);
  }
  PassArgs p;
  p.C = channels;
  if (need_h) {
    p.src = src; p.src_w = src_w; p.src_h = src_h;
    p.bounds = bounds_h; p.kk = kk_h; p.ksize = kh;
    p.out_w = nw; p.out_h = need_v ? ybox_last - ybox_first : nh;
    p.row0 = need_v ? ybox_first : 0;
    if (need_v) { p.dst = tmp; p.dst_w = nw; p.top = 0; p.left = 0; }
    else { p.dst = dst; p.dst_w = out_w; p.top = top; p.left = left; }
    resample_h_kernel<<<dim3(ceil_div(nw, 128), p.out_h), blk, 0, st>>>(p);
    DCFA_CHECK_LAUNCH("resample_h_kernel");
  }
  if (need_v) {
    p.src = need_h ? tmp : src; p.src_w = nw; p.src_h = need_h ? ybox_last - ybox_first : src_h;
    p.bounds = bounds_v; p.kk = kk_v; p.ksize = kv;
    p.out_w = nw; p.out_h = nh;
    p.row0 = need_h ? ybox_first : 0;
    p.dst = dst; p.dst_w = out_w; p.top = top; p.left = left;
    resample_v_kernel<<<dim3(ceil_div(nw, 128), nh), blk, 0, st>>>(p);
    DCFA_CHECK_LAUNCH("resample_v_kernel");
  }
  return DCFA_OK;
}

int dcfa_pack_detections(const float* det, const int32_t* cnt, int B, int A, int K, const int32_t* image_hw, int in_h, int in_w,
                         int letterbox, float* out, void* stream) {
  using namespace dcfa;
  DCFA_REQUIRE(det && cnt && out && B > 0 && A > 0 && K > 0 && K <= A, "pack_detections: bad arguments");
  pack_detections_kernel<<<dim3(ceil_div(K, 128), B), 128, 0, (cudaStream_t)stream>>>(det, cnt, B, A, K, image_hw, in_h, in_w,
                                                                                    letterbox, out);
  DCFA_CHECK_LAUNCH("pack_detections_kernel");
  return DCFA_OK;
}

}  // extern "C"
