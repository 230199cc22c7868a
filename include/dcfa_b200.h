/*
 * dcfa_b200 -- C ABI of the B200 (sm_100a) kernel library behind the DCFA-YOLO
 * dual-channel inference hot path.
 *
 * The reference (heitieya/DCFA-YOLO) has no FFI: its boundary is the Python
 * class API of three modules.  Each entry point / op kind below names the
 * reference code it replaces (file:line under the reference root):
 *
 *   DCFA_OP_STEM        nets/yolo_mul.py:104-115   Conv_maxpool (conv3x3+BN+ReLU+maxpool3/2)
 *   DCFA_OP_CONV        nets/yolo_mul.py:190-204   Conv (conv+BN+SiLU); nets/repghost.py:291-305;
 *                       nets/yolo_mul.py:138-151   ShuffleNetV2 1x1 convs (+BN+ReLU);
 *                       nets/repghost.py:80-84     RepGhostModule.primary_conv;
 *                       nets/yolo_mul.py:388-391   head cv2/cv3 (incl. final biased 1x1, fp32 NCHW out)
 *   DCFA_OP_DWCONV      nets/yolo_mul.py:144-146   ShuffleNetV2 depthwise 3x3 (+bias)+BN;
 *                       nets/repghost.py:98-115    RepGhost cheap_operation + fusion_bn (+SiLU) (+residual :279)
 *   DCFA_OP_CBAM_POOL   nets/yolo_mul.py:59-60,70-71   avg/max pool over HW (partials)
 *   DCFA_OP_CBAM_MLP    nets/yolo_mul.py:63-73     fc1-relu-fc2 on avg and max, add, sigmoid
 *   DCFA_OP_CBAM_STATS  nets/yolo_mul.py:100,86-88 x*gate, mean/max over channels
 *   DCFA_OP_CBAM_APPLY  nets/yolo_mul.py:89-90,101 conv7x7 on the 2-plane map, sigmoid, scale
 *   DCFA_OP_MAXPOOL5    nets/yolo_mul.py:17,26-30  MaxPool2d(5,1,2) inside SPPF_CBAM
 *   DCFA_OP_UPSAMPLE    nets/yolo_mul.py:421,426,433  feat3_rgb+feat3_nir, bilinear(align_corners=True)
 *   DCFA_OP_DFL         nets/yolo_mul.py:312-322,459-461  view/split + DFL softmax expectation
 *   dcfa_decode_box     utils/utils_bbox.py:30-40,49-58   dist2bbox(xywh) * strides, sigmoid, normalise
 *   dcfa_nms            utils/utils_bbox.py:87-168 + torchvision.ops.nms (un-vendored dependency)
 *   dcfa_pack_detections  utils/utils_bbox.py:60-85,170-173  yolo_correct_boxes (un-letterbox) + result packing
 *   dcfa_letterbox_u8   utils/utils.py:24-37  resize_image: PIL BICUBIC resize + grey letterbox (the caller-side pre-processing)
 *
 * Conventions
 *   - Plain pointers and sizes only; no torch types.  Every pointer is a DEVICE
 *     pointer unless stated otherwise.  The caller owns all memory.
 *   - Every function returns 0 on success, a negative DCFA_E_* code otherwise;
 *     dcfa_last_error() returns a thread-local message.  Nothing throws or
 *     aborts across the ABI.
 *   - All work is enqueued asynchronously on `stream` (a cudaStream_t passed as
 *     void*).  No host synchronisation, no allocation: every call is CUDA-graph
 *     capturable.
 *   - Activations are bf16 NHWC inside the path; fp32 at the reference-facing
 *     surface (inputs NCHW fp32, head maps NCHW fp32, dbox/cls/decoded fp32).
 */
#ifndef DCFA_B200_H_
#define DCFA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCFA_ABI_VERSION 1

enum {
  DCFA_OK = 0,
  DCFA_E_INVALID = -1,   /* bad argument / unsupported shape */
  DCFA_E_CUDA = -2,      /* a CUDA runtime call failed */
  DCFA_E_ARCH = -3       /* device is not sm_100 */
};

/* op kinds executed by dcfa_run_ops */
enum {
  DCFA_OP_STEM = 1,
  DCFA_OP_CONV = 2,
  DCFA_OP_DWCONV = 3,
  DCFA_OP_CBAM_POOL = 4,
  DCFA_OP_CBAM_MLP = 5,
  DCFA_OP_CBAM_STATS = 6,
  DCFA_OP_CBAM_APPLY = 7,
  DCFA_OP_MAXPOOL5 = 8,
  DCFA_OP_UPSAMPLE = 9,
  DCFA_OP_DFL = 10
};

/* activation applied in conv / depthwise epilogues */
enum { DCFA_ACT_NONE = 0, DCFA_ACT_RELU = 1, DCFA_ACT_SILU = 2 };

/* DCFA_OP_STEM flags: the inputs are uint8 NHWC images [group_imgs,Hi,Wi,3] (the output of cvtColor/resize_image,
 * utils/utils.py:9-37, before preprocess_input's /255, which the caller folds into `scale`) instead of fp32 NCHW */
enum {
  DCFA_STEM_FLAG_U8 = 0x100,
  /* with DCFA_STEM_FLAG_U8 and two groups: x2 is ONE uint8 plane per image [group_imgs,Hi,Wi] -- the depth image before
   * cvtColor replicates it to three channels (utils/utils.py:14-19); the kernel replicates it, so the result is
   * identical to passing the replicated image and the upload is a third of it */
  DCFA_STEM_FLAG_X2_PLANE = 0x200
};

/* DCFA_OP_CONV flags (above the low 8 bits, which hold the k-block width of the TMA packing): this 1x1 conv, the DWCONV
 * after it and the 1x1 CONV after that form a chain whose two intermediate tensors nobody else reads and whose output
 * does not alias its input -- dcfa_run_ops may run the three records as one fused kernel */
enum {
  DCFA_CONV_FLAG_CHAIN_HEAD = 0x400,
  /* 3x3 stride-2 conv, Cin = 32, dense even-width input: w holds SIX k-blocks of 64 per n-tile, one per (kernel row,
   * pixel pair): [0 | w(dy,0)] applied to input pair ox-1 and [w(dy,1) | w(dy,2)] applied to pair ox */
  DCFA_CONV_FLAG_PAIR = 0x800,
  /* this 1x1 conv and the DWCONV after it are a RepGhostModule (nets/repghost.py:70-123) whose intermediate tensor nobody
   * else reads and whose output does not alias its input -- dcfa_run_ops may run the two records as one fused kernel */
  DCFA_CONV_FLAG_GHOST_HEAD = 0x1000,
  /* the last conv of a head level (nets/yolo_mul.py:388-391 merged: Cout = 64 box + nc class channels, fp32 NCHW map
   * x[i]): the epilogue ALSO applies DFL (:312-322) to the box channels and gathers the level into the (B, 4, A) /
   * (B, nc, A) tensors of :459-461 -- a1 = dbox fp32 [n_img,4,A], a2 = cls fp32 [n_img,nc,A], A = total anchors,
   * hidden = first anchor of this level, nc = classes.  Replaces a separate DCFA_OP_DFL pass over the maps. */
  DCFA_CONV_FLAG_DFL = 0x2000
};

/* DCFA_OP_CONV output modes */
enum { DCFA_OUT_BF16_NHWC = 0, DCFA_OUT_F32_NCHW = 1 };

/*
 * A strided view of an activation tensor (or a flat parameter array).
 *   address(n, y, x, c) = bufs[buf] + off
 *                         + elem_size * ( (n % gi) * img_stride + (n / gi) * gstride
 *                                         + (y * W + x) * ld + c )
 * `gi` (images per group) and `gstride` let one launch cover both modalities:
 * either as a [2B,...] tensor (gstride = gi*img_stride) or as two channel slots
 * of one concat buffer (gstride = slot width).  gi == 0 means "no grouping".
 * For parameter arrays only buf/off are used.
 */
typedef struct dcfa_view {
  int32_t buf;        /* index into the bufs[] array given to dcfa_run_ops; -1 = absent */
  int32_t ld;         /* elements between consecutive pixels (NHWC) */
  int64_t off;        /* byte offset inside the buffer */
  int64_t img_stride; /* elements between consecutive images of one group */
  int64_t gstride;    /* elements between groups */
  int32_t gi;         /* images per group (0: single group) */
  int32_t pad_;
} dcfa_view;

/*
 * One op of the flat execution plan.  Field use per kind:
 *
 * STEM   x,x2 = fp32 NCHW inputs of group 0 / group 1 [group_imgs,3,Hi,Wi]; w = bf16 [G][nblk][3][128*16]: per group,
 *        64-channel block (nblk = max(1, BN/64)) and kernel row ky one K-major tile in the canonical NO-SWIZZLE layout
 *        (element (m,k) at (m/8)*128 + (k/8)*64 + (m%8)*8 + k%8).  Row m = 32*q + 16*e + cl: e = 0 rows carry the weights
 *        for EVEN conv columns (K = kx*4 + ci: four consecutive pixels x 4 channel slots, unused slots zero), e = 1 rows
 *        those for ODD conv columns (K = (kx+1)*4 + ci); row slot j = 16*q + cl holds channel 64*blk + j % min(BN, 64),
 *        BN = Cout rounded up to 32/64/128; rows of channels with a negative BN scale are negated; w_gstride = nblk*3*128*16;
 *        scale (>= 0), bias = fp32 [G][BN]; y = bf16 NHWC [n_img,Ho,Wo,Cout], Ho = (Hi-1)/2+1, Wo likewise; Cout % 8 == 0.
 *        With flags & DCFA_STEM_FLAG_U8: x,x2 = uint8 NHWC [group_imgs,Hi,Wi,3] (x2 a single plane [group_imgs,Hi,Wi]
 *        with DCFA_STEM_FLAG_X2_PLANE), scale already divided by 255.
 * CONV   x = bf16 NHWC input view (Cin channels starting at the view's offset); w = bf16 packed
 *        [G][n_tiles][k_blocks][BN*64] (128B-swizzled K-major tile images, K = (ky*ks+kx)*Cin+ci, zero
 *        padded to k_blocks*64); scale,bias = fp32 [G][n_tiles*BN]; x2 = optional bf16 residual added
 *        after the activation; y = output (out_mode).  f0 = post-activation scale (1.0 = none).
 *        parts > 0 (bf16 NHWC, TMA path only): SPLIT output -- channels [0, parts) go to y, channels [parts, Cout) to
 *        the view a0 (whose channel 0 is channel `parts`); parts % 16 == 0.
 * DWCONV x -> y depthwise 3x3 s1 p1, w = fp32 [G][9][Cin] (BN folded), bias = fp32 [G][Cin], act,
 *        x2 = optional residual (added after activation).
 * CBAM_POOL   x -> a0 = fp32 [n_img][parts][Cin] sums, a1 = same shape maxima; parts pixel chunks.
 * CBAM_MLP    a0,a1 partials -> a2 = fp32 gate [n_img][Cin]; w = fp32 fc1 [G][hidden][Cin],
 *             scale = fp32 fc2 [G][Cin][hidden]; Hi*Wi is the pooling divisor.
 * CBAM_STATS  x, a2 gate -> a0 = fp32 [n_img][Hi][Wi][2] (mean_c, max_c of x*gate).
 * CBAM_APPLY  x, a2 gate, a0 stats, w = fp32 [G][2][7][7] -> y = x*gate*sigmoid(conv7x7(stats)).
 * MAXPOOL5    x -> y, 5x5 s1 p2, -inf padding.
 * UPSAMPLE    x (+ x2 if present, summed first) [n_img,Hi,Wi,Cin] -> y [n_img,Ho,Wo,Cin] bilinear,
 *             align_corners=True.
 * DFL    x = fp32 NCHW head maps of the 3 levels at a0.off/a1.off/a2.off inside buffers a0/a1/a2.buf
 *        ([n_img, 64+nc, H_l, W_l]; level 0 is Hi x Wi, each further level halves with ceil);
 *        y = fp32 dbox [n_img,4,A]; x2 = fp32 cls logits out [n_img,nc,A].
 */
typedef struct dcfa_op {
  int32_t kind;
  int32_t act;
  int32_t out_mode;
  int32_t flags;
  dcfa_view x, x2, y, w, scale, bias, a0, a1, a2;
  int32_t n_img, group_imgs;
  int32_t Hi, Wi, Cin;
  int32_t Ho, Wo, Cout;
  int32_t ksize, stride;
  int32_t BN, n_tiles, k_blocks, K_real;
  int32_t hidden, parts, nc, A;
  int32_t out_ctot;   /* CONV F32_NCHW: total channels of the destination tensor */
  int32_t out_coff;   /* CONV F32_NCHW: first destination channel */
  int64_t w_gstride;  /* elements between the weight sets of two groups */
  int64_t sb_gstride; /* elements between the scale/bias (or fc2) sets of two groups */
  float f0, f1, f2, f3;
} dcfa_op;

/* ABI self-description (the Python loader checks these against its ctypes mirror). */
int dcfa_abi_version(void);
int dcfa_sizeof_view(void);
int dcfa_sizeof_op(void);

/* Thread-local description of the last failure. */
const char* dcfa_last_error(void);

/* 0 when device `dev` can run the library (compute capability 10.x), else DCFA_E_ARCH/CUDA. */
int dcfa_device_check(int dev);

/* Number of kernels launched by this process through the library so far. */
int64_t dcfa_launch_count(void);

/*
 * Execute `n_ops` ops in order on `stream`.  `ops` is HOST memory (read during
 * the call only).  `bufs[nbufs]` are the device base pointers the views index.
 * Consecutive records may run as ONE kernel when their shapes allow: CBAM_POOL, CBAM_MLP, CBAM_STATS, CBAM_APPLY of one
 * tensor (a thread-block cluster per image; the partial-sum / gate / stats scratch views are then not written), and
 * CONV(1x1) -> DWCONV -> CONV(1x1) chains whose head carries DCFA_CONV_FLAG_CHAIN_HEAD (the two intermediate tensors
 * are then not written), and CONV(1x1, DCFA_CONV_FLAG_GHOST_HEAD) -> DWCONV pairs.  Results are the same either way.
 */
int dcfa_run_ops(const dcfa_op* ops, int n_ops, void* const* bufs, int nbufs, void* stream);

/*
 * The same op list as a PLAN object: prepared once (kernel variants chosen, TMA tensor maps encoded, launch geometry and
 * arguments fixed), then replayed with one launch call per kernel -- what a long-lived caller should use.
 *   bufs[i] != NULL at creation: bound for the plan's lifetime (the parameter blob and the activation arena);
 *   bufs[i] == NULL: supplied at every dcfa_plan_run (inputs / outputs); the few ops that touch such a buffer are
 *   re-prepared only when its pointer differs from the previous run.
 * A plan belongs to the device that was current at creation and is not thread-safe; runs on one stream at a time share
 * its arena.  dcfa_plan_run enqueues asynchronously and is CUDA-graph capturable.
 */
typedef struct dcfa_plan dcfa_plan;
int dcfa_plan_create(const dcfa_op* ops, int n_ops, void* const* bufs, int nbufs, dcfa_plan** out);
int dcfa_plan_run(dcfa_plan* plan, void* const* bufs, int nbufs, void* stream);
int dcfa_plan_num_launches(const dcfa_plan* plan);   /* kernels per run (0 for parts not prepared yet) */
void dcfa_plan_destroy(dcfa_plan* plan);

/*
 * Buffer indices of the plans the host-side compiler emits for YoloBody.forward (nets/yolo_mul.py:397-462):
 * what the views of its ops reference.
 */
enum {
  DCFA_BUF_BLOB = 0,   /* packed parameters (folded BN, pre-tiled bf16 weights) */
  DCFA_BUF_ARENA = 1,  /* activations */
  DCFA_BUF_RGB = 2,    /* input: fp32 [B,3,H,W], or uint8 [B,H,W,3] (plans compiled for uint8 input) */
  DCFA_BUF_NIR = 3,    /* input: second modality; uint8 [B,H,W] for depth-plane plans */
  DCFA_BUF_X0 = 4,     /* outputs: head maps fp32 [B,64+nc,H/8,W/8], [.., H/16, W/16], [.., H/32, W/32] */
  DCFA_BUF_X1 = 5,
  DCFA_BUF_X2 = 6,
  DCFA_BUF_DBOX = 7,   /* output: fp32 [B,4,A] */
  DCFA_BUF_CLS = 8,    /* output: fp32 [B,nc,A] */
  DCFA_NUM_BUFS = 9
};

/* Facts about a compiled forward plan, stored in plan files. */
typedef struct dcfa_plan_info {
  int32_t batch, height, width, num_classes;
  int32_t anchors;          /* A = sum of the three level sizes */
  int32_t no;               /* 64 + num_classes */
  int32_t level_hw[3][2];   /* (H_l, W_l) of the three head levels */
  int32_t input_u8;         /* inputs are uint8 NHWC images */
  int32_t depth_plane;      /* the second input is a single uint8 plane */
  int32_t reserved[3];
} dcfa_plan_info;

/* Plan file = this header, n_ops dcfa_op records, blob_bytes of parameters (written by dcfa_b200.plan.Plan.save). */
typedef struct dcfa_plan_file_header {
  char magic[8];            /* "DCFAPLN1" */
  int32_t abi_version, sizeof_op, n_ops, nbufs;
  int64_t blob_bytes, arena_bytes;
  dcfa_plan_info info;
} dcfa_plan_file_header;

/*
 * For callers without the Python plan compiler: load a plan file, upload its parameters and allocate its arena (the
 * library owns both until dcfa_plan_destroy), then run YoloBody.forward on device pointers.
 *   rgb/depth: inputs as the plan was compiled for (see dcfa_plan_get_info); x0..x2, dbox, cls: outputs as above.
 */
int dcfa_plan_load(const char* path, dcfa_plan** out);
int dcfa_plan_get_info(const dcfa_plan* plan, dcfa_plan_info* info);
int dcfa_plan_forward(dcfa_plan* plan, const void* rgb, const void* depth, float* x0, float* x1, float* x2, float* dbox,
                      float* cls, void* stream);

/*
 * DecodeBox.decode_box (utils/utils_bbox.py:49-58).
 *   dbox [B,4,A] fp32 (l,t,r,b distances), cls [B,nc,A] fp32 logits (batch stride cls_bstride
 *   elements, row stride A), anchors [2,A] fp32 with element (k,a) at anchors[k*anc_s0 + a*anc_s1],
 *   strides [A] fp32 -> out [B,A,4+nc] fp32 contiguous: (cx,cy,w,h)/(W,H,W,H), sigmoid(cls).
 */
int dcfa_decode_box(const float* dbox, const float* cls, int64_t cls_bstride,
                    const float* anchors, int64_t anc_s0, int64_t anc_s1, const float* strides,
                    int B, int A, int nc, float img_w, float img_h, float* out, void* stream);

/* IoU arithmetic of the un-vendored torchvision.ops.nms the reference calls (utils/utils_bbox.py:145). */
enum {
  DCFA_IOU_TV_CPU = 0,  /* areas rounded separately, fp32 IoU compared with the double threshold   */
  DCFA_IOU_TV_CUDA = 1  /* Sa + fma(wj,hj) fused as torchvision's CUDA kernel, float threshold     */
};

/* Bytes of scratch dcfa_nms needs for (B, A). */
int64_t dcfa_nms_workspace_bytes(int B, int A);

/*
 * DecodeBox.non_max_suppression up to (not including) the host-side un-letterbox
 * (utils/utils_bbox.py:92-168).
 *   pred [B,A,4+nc] fp32: columns 0..3 are rewritten IN PLACE from (cx,cy,w,h) to (x1,y1,x2,y2), as
 *   the reference does (:92-97).  Per image: class max (first max wins), keep rows with
 *   conf >= conf_thres, per class (ascending) greedy NMS in stable descending-score order.
 *   out_det  [B,A,6] fp32: kept rows (x1,y1,x2,y2,conf,cls) in the reference's output order.
 *   out_idx  [B,A] int32: anchor index of every kept row (same order).
 *   out_cnt  [B]   int32: kept rows per image;  out_cand [B] int32 (may be NULL): candidates after
 *   the confidence filter.
 *   Limits: nc <= 256, A < 2^24.
 */
int dcfa_nms(float* pred, int B, int A, int nc, float conf_thres, double nms_thres, int iou_mode,
             float* out_det, int32_t* out_idx, int32_t* out_cnt, int32_t* out_cand,
             void* workspace, int64_t workspace_bytes, void* stream);

/*
 * resize_image (utils/utils.py:24-37) on the device: Pillow's 8-bit `image.resize((nw, nh), Image.BICUBIC)` -- two separable
 * fixed-point passes with an 8-bit intermediate, antialiased when shrinking -- pasted centred on a grey (128) canvas when
 * `letterbox` is non-zero (nw = int(iw*s), nh = int(ih*s), s = min(W/iw, H/ih)), or stretched to the full canvas otherwise.
 * Bit-exact against Pillow (an un-vendored dependency of the reference; pinned empirically, tests/test_letterbox_*.py).
 *   src [src_h, src_w, channels] uint8 (channels = 3, or 1 for the depth plane before cvtColor replicates it);
 *   dst [out_h, out_w, channels] uint8 (one image slot of the batch tensor the stem kernel reads);
 *   workspace: dcfa_letterbox_workspace_bytes(...) bytes of device scratch (weight tables + the horizontal pass's output).
 */
int64_t dcfa_letterbox_workspace_bytes(int src_h, int src_w, int channels, int out_h, int out_w);
int dcfa_letterbox_u8(const uint8_t* src, int src_h, int src_w, int channels, uint8_t* dst, int out_h, int out_w,
                      int letterbox, void* workspace, int64_t workspace_bytes, void* stream);

/*
 * The tail of DecodeBox.non_max_suppression (utils/utils_bbox.py:170-173 + yolo_correct_boxes :60-85) on the device, fused
 * with the packing that precedes the single device->host copy of a batch:
 *   det [B,A,6], cnt [B]: dcfa_nms outputs.  out [B, 1 + 6*K] fp32: per image (count, first K rows).
 *   image_hw [B,2] int32 (device) = original (height, width) of every image: rows become (y1, x1, y2, x2, conf, cls) in
 *   original-image pixels, in the reference's dtype flow (float64 intermediates, float32 `box_hw *= scale`), bit-exact
 *   against the numpy code.  image_hw == NULL: rows are copied unchanged (x1, y1, x2, y2, conf, cls).
 */
int dcfa_pack_detections(const float* det, const int32_t* cnt, int B, int A, int K, const int32_t* image_hw, int in_h, int in_w,
                         int letterbox, float* out, void* stream);

/*
 * The criterion of the reference's validation loop, forward only: Loss.__call__ (nets/yolo_training.py:371-430) with the
 * task-aligned assigner (:75-225), CIoU (:227-262) and BboxLoss (:272-303), applied to what YoloBody.forward returns
 * (utils/utils_fit_mul.py:78-92 calls it under no_grad on the eval-mode outputs).  SURVEY 8(f) N4.
 *   x0..x2       the three head maps [B, 64 + nc, H_l, W_l] fp32 NCHW contiguous (forward's `x` list)
 *   level_hw     HOST int32[6] = (H_0, W_0, H_1, W_1, H_2, W_2);  level_stride: HOST float[3] (model.stride)
 *   gt           DEVICE [B, G, 5] fp32 rows (class, x1, y1, x2, y2) in input pixels, zero padded to G rows per image --
 *                the output of Loss.preprocess (:342-360); G = 0 (gt may be NULL): no targets in the batch
 *   out          DEVICE float[8]: box * 7.5, cls * 0.5, dfl * 1.5, their sum (= the value the reference returns),
 *                target_scores_sum, number of foreground anchors, 0, 0
 *   workspace    dcfa_loss_workspace_bytes(B, A, nc, G) bytes of device scratch, A = sum of H_l * W_l
 * Four launches on `stream`, no host synchronisation, deterministic (fixed-order sums).  Among anchors whose alignment
 * metric is EQUAL at the top-10 boundary the lowest anchor index is taken (torch.topk leaves that order unspecified).
 * Limits: reg_max = 16, A * 4 bytes and G * 128 bytes of shared memory (A <= 56 320, G <= 1 760), labels in [0, nc).
 */
int64_t dcfa_loss_workspace_bytes(int B, int A, int nc, int G);
int dcfa_yolo_loss(const float* x0, const float* x1, const float* x2, int B, int nc, const int32_t* level_hw,
                   const float* level_stride, const float* gt, int G, float* out, void* workspace, int64_t workspace_bytes,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCFA_B200_H_ */
