"""Runtime of the B200 hot path: owns the packed parameter blob and the activation arena on one device and
enqueues a compiled Plan (dcfa_run_ops) on the caller's CUDA stream.  PyTorch is used for device memory and
streams only; every kernel launched here comes from lib/libdcfa_b200.so."""
import ctypes as C

import torch

from . import _lib, abi
from . import plan as P


class Engine:
    """One (batch, H, W) instance of the forward pass on one device."""

    def __init__(self, state_dict, phi, num_classes, batch, height, width, device, input_u8=False, depth_plane=False):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("dcfa_b200 has no CPU path: the engine needs a CUDA (sm_100a) device")
        _lib.check(_lib.lib.dcfa_device_check(self.device.index if self.device.index is not None else torch.cuda.current_device()))
        self.input_u8 = bool(input_u8)
        self.depth_plane = bool(depth_plane)
        self.plan = P.Plan(state_dict, phi, num_classes, batch, height, width, input_u8=self.input_u8,
                           depth_plane=self.depth_plane)
        self.B, self.H, self.W, self.nc, self.no, self.A = batch, height, width, self.plan.nc, self.plan.no, self.plan.A
        self.level_shapes = list(self.plan.level_shapes)
        self.blob = self.plan.blob_tensor.to(self.device)
        self.arena = torch.empty(self.plan.arena_bytes + 256, dtype=torch.uint8, device=self.device)
        self.n_ops = len(self.plan.ops)
        self._bufs = (C.c_void_p * P.NUM_BUFS)()
        self._bufs[P.BUF_BLOB] = self.blob.data_ptr()
        self._bufs[P.BUF_ARENA] = self.arena.data_ptr()
        # the C-side plan object: blob and arena bound, inputs / outputs supplied per run; every launch prepared once
        self._plan = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.dcfa_plan_create(self.plan.op_array, self.n_ops, self._bufs, P.NUM_BUFS, C.byref(self._plan)))
        self.anchors, self.strides = self._make_anchors()

    def __del__(self):
        plan = getattr(self, "_plan", None)
        if plan:
            _lib.lib.dcfa_plan_destroy(plan)
            self._plan = None

    def _make_anchors(self):
        """make_anchors (reference utils/utils_bbox.py:16-28), always fp32; returned as (2,A) and (1,A)."""
        pts, st = [], []
        for (h, w), s in zip(self.level_shapes, (8.0, 16.0, 32.0)):
            sx = torch.arange(w, dtype=torch.float32, device=self.device) + 0.5
            sy = torch.arange(h, dtype=torch.float32, device=self.device) + 0.5
            sy, sx = torch.meshgrid(sy, sx, indexing="ij")
            pts.append(torch.stack((sx, sy), -1).view(-1, 2))
            st.append(torch.full((h * w, 1), s, dtype=torch.float32, device=self.device))
        return torch.cat(pts).t().contiguous(), torch.cat(st).t().contiguous()

    def new_outputs(self):
        dev, b = self.device, self.B
        x = [torch.empty(b, self.no, h, w, dtype=torch.float32, device=dev) for (h, w) in self.level_shapes]
        dbox = torch.empty(b, 4, self.A, dtype=torch.float32, device=dev)
        cls = torch.empty(b, self.nc, self.A, dtype=torch.float32, device=dev)
        return dbox, cls, x

    def run(self, rgb, nir, outputs=None, stream=None):
        """rgb, nir: contiguous CUDA tensors, float32 [B,3,H,W] (or uint8 [B,H,W,3] for an input_u8 engine; nir is
        uint8 [B,H,W] for a depth_plane engine).  Enqueues the whole forward; returns (dbox, cls, x).

        One engine = one activation arena: calls on the same engine must be issued on ONE stream at a time (or be
        ordered by events); concurrent streams need one engine (YoloBody) each."""
        shp = (self.B, self.H, self.W, 3) if self.input_u8 else (self.B, 3, self.H, self.W)
        dt = torch.uint8 if self.input_u8 else torch.float32
        for t, want in ((rgb, shp), (nir, (self.B, self.H, self.W) if self.depth_plane else shp)):
            if tuple(t.shape) != want or t.dtype != dt or not t.is_contiguous() or t.device != self.device:
                raise ValueError("engine.run: expected contiguous %s %s on %s, got %s %s on %s" % (
                    dt, want, self.device, tuple(t.shape), t.dtype, t.device))
        with torch.cuda.device(self.device):   # the library launches on the CURRENT device: make it this engine's
            dbox, cls, x = outputs if outputs is not None else self.new_outputs()
            b = (C.c_void_p * P.NUM_BUFS)(*self._bufs)   # per-call pointer table (calls may come from several host threads)
            b[P.BUF_RGB], b[P.BUF_NIR] = rgb.data_ptr(), nir.data_ptr()
            b[P.BUF_X0], b[P.BUF_X1], b[P.BUF_X2] = x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr()
            b[P.BUF_DBOX], b[P.BUF_CLS] = dbox.data_ptr(), cls.data_ptr()
            st = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib.dcfa_plan_run(self._plan, b, P.NUM_BUFS, C.c_void_p(st)))
        self.last_bufs = b   # the pointer table of the latest call (profiling tools replay single ops with it)
        return dbox, cls, x


def decode_box(dbox, cls, anchors, strides, input_shape):
    """dcfa_decode_box on CUDA tensors -> [B, A, 4+nc] float32 (contiguous)."""
    if not dbox.is_cuda:
        raise RuntimeError("dcfa_b200 has no CPU path: decode_box needs CUDA tensors")
    b, _, a = dbox.shape
    nc = cls.shape[1]
    dbox = dbox.float().contiguous()
    if cls.dtype != torch.float32 or cls.stride(2) != 1 or cls.stride(1) != a:
        cls = cls.float().contiguous()
    anchors = anchors.float().to(dbox.device)
    strides = strides.float().to(dbox.device).reshape(-1).contiguous()
    with torch.cuda.device(dbox.device):
        out = torch.empty(b, a, 4 + nc, dtype=torch.float32, device=dbox.device)
        st = torch.cuda.current_stream(dbox.device).cuda_stream
        _lib.check(_lib.lib.dcfa_decode_box(dbox.data_ptr(), cls.data_ptr(), cls.stride(0), anchors.data_ptr(),
                                            anchors.stride(0), anchors.stride(1), strides.data_ptr(), b, a, nc,
                                            float(input_shape[1]), float(input_shape[0]), out.data_ptr(), C.c_void_p(st)))
    return out


class NmsWorkspace:
    """Device scratch + outputs for dcfa_nms at a fixed (B, A)."""

    def __init__(self, b, a, device):
        self.b, self.a = b, a
        self.bytes = int(_lib.lib.dcfa_nms_workspace_bytes(b, a))
        self.ws = torch.empty(self.bytes, dtype=torch.uint8, device=device)
        self.det = torch.empty(b, a, 6, dtype=torch.float32, device=device)
        self.idx = torch.empty(b, a, dtype=torch.int32, device=device)
        self.cnt = torch.zeros(b, dtype=torch.int32, device=device)
        self.cand = torch.zeros(b, dtype=torch.int32, device=device)


def nms(pred, conf_thres, nms_thres, iou_mode=abi.IOU_TV_CPU, workspace=None):
    """dcfa_nms on a contiguous float32 CUDA tensor [B,A,4+nc] (boxes rewritten in place to xyxy).
    Returns the NmsWorkspace holding det/idx/cnt/cand (device tensors, stream-ordered)."""
    if not pred.is_cuda:
        raise RuntimeError("dcfa_b200 has no CPU path: nms needs a CUDA tensor")
    assert pred.dtype == torch.float32 and pred.is_contiguous() and pred.dim() == 3
    b, a, row = pred.shape
    with torch.cuda.device(pred.device):
        ws = workspace if workspace is not None else NmsWorkspace(b, a, pred.device)
        assert ws.b == b and ws.a == a
        st = torch.cuda.current_stream(pred.device).cuda_stream
        _lib.check(_lib.lib.dcfa_nms(pred.data_ptr(), b, a, row - 4, float(conf_thres), float(nms_thres), int(iou_mode),
                                     ws.det.data_ptr(), ws.idx.data_ptr(), ws.cnt.data_ptr(), ws.cand.data_ptr(),
                                     ws.ws.data_ptr(), ws.bytes, C.c_void_p(st)))
    return ws


def letterbox_u8(src, dst, letterbox_image, workspace=None):
    """resize_image (reference utils/utils.py:24-37) on the device: src uint8 CUDA [h,w,3] or [h,w] -> dst uint8 CUDA
    [H,W,3] / [H,W] (typically one image slot of the batch tensor), bit-exact with PIL's BICUBIC + grey letterbox."""
    if not (src.is_cuda and dst.is_cuda and src.dtype == torch.uint8 and dst.dtype == torch.uint8):
        raise RuntimeError("dcfa_b200 has no CPU path: letterbox_u8 needs uint8 CUDA tensors")
    c = 1 if src.dim() == 2 else int(src.shape[2])
    if (dst.dim() == 2) != (src.dim() == 2) or (dst.dim() == 3 and dst.shape[2] != c) or not (src.is_contiguous() and dst.is_contiguous()):
        raise ValueError("letterbox_u8: src %s / dst %s mismatch" % (tuple(src.shape), tuple(dst.shape)))
    sh, sw, oh, ow = int(src.shape[0]), int(src.shape[1]), int(dst.shape[0]), int(dst.shape[1])
    need = int(_lib.lib.dcfa_letterbox_workspace_bytes(sh, sw, c, oh, ow))
    with torch.cuda.device(src.device):
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=src.device)
        st = torch.cuda.current_stream(src.device).cuda_stream
        _lib.check(_lib.lib.dcfa_letterbox_u8(src.data_ptr(), sh, sw, c, dst.data_ptr(), oh, ow, 1 if letterbox_image else 0,
                                              workspace.data_ptr(), workspace.numel(), C.c_void_p(st)))
    return workspace


def pack_detections(ws, k, out, image_hw=None, input_shape=None, letterbox_image=True):
    """(count, first k rows) per image of an NmsWorkspace into out [B, 1 + 6k] fp32 (device).  With image_hw (int32 CUDA
    [B,2], original (h, w) per image) the rows are un-letterboxed on the device exactly as DecodeBox.yolo_correct_boxes
    does on the host (reference utils/utils_bbox.py:60-85)."""
    b, a = ws.b, ws.a
    with torch.cuda.device(ws.det.device):
        st = torch.cuda.current_stream(ws.det.device).cuda_stream
        ih, iw = (int(input_shape[0]), int(input_shape[1])) if input_shape is not None else (0, 0)
        _lib.check(_lib.lib.dcfa_pack_detections(ws.det.data_ptr(), ws.cnt.data_ptr(), b, a, int(k),
                                                 image_hw.data_ptr() if image_hw is not None else None, ih, iw,
                                                 1 if letterbox_image else 0, out.data_ptr(), C.c_void_p(st)))
    return out
