"""Drop-in for the reference's utils/utils_bbox.py: same names and signatures
(make_anchors :16-28, dist2bbox :30-40, DecodeBox.decode_box :49-58, .yolo_correct_boxes :60-85,
.non_max_suppression :87-174), with decode and NMS executed by lib/libdcfa_b200.so on the GPU.
"""
import numpy as np
import torch

from dcfa_b200 import abi
from dcfa_b200 import engine as _engine


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """Anchor centres (A,2) and per-anchor stride (A,1) for a list of feature maps.  Always fp32: under bf16
    features the reference's dtype-following version (:20) cannot represent 159.5."""
    assert feats is not None
    device = feats[0].device
    pts, st = [], []
    for f, s in zip(feats, strides):
        h, w = f.shape[-2:]
        gy, gx = torch.meshgrid(torch.arange(h, device=device, dtype=torch.float32) + grid_cell_offset,
                                torch.arange(w, device=device, dtype=torch.float32) + grid_cell_offset, indexing='ij')
        pts.append(torch.stack((gx, gy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=torch.float32, device=device))
    return torch.cat(pts), torch.cat(st)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """(l,t,r,b) distances -> boxes; used by the training loss of the reference (kept for import parity)."""
    lt, rb = torch.split(distance, 2, dim)
    x1y1, x2y2 = anchor_points - lt, anchor_points + rb
    if xywh:
        return torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), dim)
    return torch.cat((x1y1, x2y2), dim)


class DecodeBox():
    #: IoU arithmetic of torchvision.ops.nms to reproduce: 'cpu' (the reference's CPU-runnable path, pinned by
    #: the golden vectors) or 'cuda' (torchvision's CUDA kernel: fused area sum, float threshold).  The YOLO facade sets
    #: 'cuda' when it runs on a GPU, as the reference (cuda=True) then calls torchvision's CUDA kernel.
    iou_mode = 'cpu'
    #: detections copied back per image in the first device->host transfer (more are fetched on demand)
    first_fetch = 512
    #: NMS workspaces per (batch, anchors, device): nms_device hands them out round-robin, so up to `ring - 1` batches
    #: can be in flight (nms_device / start_fetch issued, fetch_detections not yet called) on ONE DecodeBox without a
    #: later batch overwriting the rows an earlier fetch will read.
    ring = 3

    def __init__(self, num_classes, input_shape):
        super(DecodeBox, self).__init__()
        self.num_classes = num_classes
        self.bbox_attrs = 4 + num_classes
        self.input_shape = input_shape
        self._nms_ws = {}

    def decode_box(self, inputs):
        dbox, cls, origin_cls, anchors, strides = inputs
        return _engine.decode_box(dbox, cls, anchors, strides, self.input_shape)

    def yolo_correct_boxes(self, box_xy, box_wh, input_shape, image_shape, letterbox_image):
        """numpy, host side, post-NMS (tiny): undo the letterbox; rows (y1, x1, y2, x2) in original-image pixels.
        image_shape is (2,) or one (h, w) row per box.  The dtype flow is the reference's (:60-85): its in-place
        `box_hw *= scale` and `boxes *= image_shape` keep the arrays' dtypes, so float32 inputs are rounded to
        float32 at exactly those two points and the results are bit-identical."""
        yx, hw = box_xy[..., ::-1], box_wh[..., ::-1]
        input_shape, image_shape = np.array(input_shape), np.array(image_shape)
        if letterbox_image:
            new_shape = np.round(image_shape * np.min(input_shape / image_shape, axis=-1, keepdims=True))
            offset = (input_shape - new_shape) / 2. / input_shape
            scale = input_shape / new_shape
            yx = (yx - offset) * scale
            hw = (hw * scale).astype(hw.dtype)
        lo, hi = yx - hw / 2., yx + hw / 2.
        boxes = np.concatenate([lo[..., 0:1], lo[..., 1:2], hi[..., 0:1], hi[..., 1:2]], axis=-1)
        return (boxes * np.concatenate([image_shape, image_shape], axis=-1)).astype(boxes.dtype)

    def nms_device(self, prediction, conf_thres=0.5, nms_thres=0.4):
        """GPU part only: returns the NmsWorkspace (det/idx/cnt on the device, stream-ordered, no host sync)."""
        if not (torch.is_tensor(prediction) and prediction.is_cuda):
            raise RuntimeError("dcfa_b200 has no CPU path: non_max_suppression needs a CUDA tensor")
        pred = prediction
        if pred.dtype != torch.float32 or not pred.is_contiguous():
            pred = pred.float().contiguous()
        if pred.shape[2] != 4 + self.num_classes:
            raise ValueError("prediction has %d columns, expected 4 + num_classes = %d" % (pred.shape[2], 4 + self.num_classes))
        key = (pred.shape[0], pred.shape[1], str(pred.device))
        slot = self._nms_ws.setdefault(key, [[], 0])   # [workspaces, next index]
        if len(slot[0]) < max(int(self.ring), 1):
            slot[0].append(_engine.NmsWorkspace(pred.shape[0], pred.shape[1], pred.device))
            ws = slot[0][-1]
        else:
            ws = slot[0][slot[1] % len(slot[0])]
            slot[1] += 1
            if getattr(ws, 'head_pending', False):
                raise RuntimeError("DecodeBox: %d batches are in flight on this instance (nms_device without fetch_detections); "
                                   "raise DecodeBox.ring or fetch earlier" % len(slot[0]))
        mode = abi.IOU_TV_CUDA if self.iou_mode == 'cuda' else abi.IOU_TV_CPU
        _engine.nms(pred, conf_thres, nms_thres, mode, ws)
        if pred is not prediction:   # keep the reference's in-place xywh -> xyxy side effect (:97)
            prediction[:, :, :4] = pred[:, :, :4].to(prediction.dtype)
        return ws

    def start_fetch(self, ws, input_shape=None, image_shape=None, letterbox_image=True):
        """Enqueue, right behind the NMS kernels on the current stream, ONE kernel that packs (count, first rows) per image
        -- and, when image_shape is given, un-letterboxes the boxes on the device exactly as yolo_correct_boxes does on the
        host (reference :60-85, :170-173; float64 intermediates, bit-identical) -- then the copy into pinned host memory, and
        record an event.  Lets a caller launch the next batch before it collects this one (fetch_detections then only waits
        for the event).  image_shape: the original (h, w) of the images, or a (B, 2) array with one shape per image."""
        b, a = ws.b, ws.a
        k = min(a, self.first_fetch)
        dev = ws.det.device
        if getattr(ws, 'head_k', None) != k:
            ws.head_dev = torch.empty(b, 1 + k * 6, dtype=torch.float32, device=dev)
            ws.head_host = torch.empty(b, 1 + k * 6, dtype=torch.float32).pin_memory()
            ws.hw_host = torch.empty(b, 2, dtype=torch.int32).pin_memory()
            ws.hw_dev = torch.empty(b, 2, dtype=torch.int32, device=dev)
            ws.head_ready = torch.cuda.Event()
            ws.head_k = k
        hw = None
        if image_shape is not None:
            shapes = np.asarray(image_shape)
            if shapes.ndim == 2 and shapes.shape[0] != b:
                raise ValueError("image_shape must be (2,) or (%d, 2), got %s" % (b, tuple(shapes.shape)))
            ws.hw_host.numpy()[...] = shapes.astype(np.int32)           # broadcasts a single (h, w)
            ws.hw_dev.copy_(ws.hw_host, non_blocking=True)
            hw = ws.hw_dev
        _engine.pack_detections(ws, k, ws.head_dev, hw, input_shape, letterbox_image)
        ws.head_host.copy_(ws.head_dev, non_blocking=True)
        ws.head_ready.record(torch.cuda.current_stream(dev))
        ws.head_pending = True
        ws.head_corrected = hw is not None

    def fetch_detections(self, ws, input_shape, image_shape, letterbox_image):
        """Host half of non_max_suppression: waits for the one device->host copy of (count, first rows) per image.  If
        start_fetch already un-letterboxed on the device the rows are final; else the reference's numpy un-letterbox
        (:170-173) runs here.  image_shape is the original (h, w) of the image, or a (B, 2) array with one shape per image
        of the batch.  Returns list of None | float32 (n_i, 6) rows (y1,x1,y2,x2,conf,cls)."""
        b, a = ws.b, ws.a
        k = min(a, self.first_fetch)
        if not getattr(ws, 'head_pending', False):
            self.start_fetch(ws, input_shape, image_shape, letterbox_image)
        ws.head_ready.synchronize()
        ws.head_pending = False
        head = ws.head_host.numpy()
        counts = head[:, 0].astype(np.int64)
        output = [None for _ in range(b)]
        shapes = np.asarray(image_shape)
        per_image = shapes.ndim == 2          # (B, 2): every image has its own original shape (batched facade)
        if per_image and shapes.shape[0] != b:
            raise ValueError("image_shape must be (2,) or (%d, 2), got %s" % (b, tuple(shapes.shape)))
        if counts.max(initial=0) <= k:
            rows = head[:, 1:].reshape(b, k, 6)
            det = rows[np.arange(k)[None, :] < counts[:, None]]            # (sum n_i, 6), image order, a copy
            if det.shape[0]:
                if not ws.head_corrected:
                    # un-letterbox all boxes in one numpy pass (elementwise arithmetic: identical to per-image calls)
                    box_xy, box_wh = (det[:, 0:2] + det[:, 2:4]) / 2, det[:, 2:4] - det[:, 0:2]
                    row_shapes = np.repeat(shapes, counts, axis=0) if per_image else shapes
                    det[:, :4] = self.yolo_correct_boxes(box_xy, box_wh, input_shape, row_shapes, letterbox_image)
                ends = np.cumsum(counts)
                for i in range(b):
                    if counts[i]:
                        output[i] = det[ends[i] - counts[i]:ends[i]]
            return output
        for i in range(b):   # some image kept more than first_fetch rows: fetch those from the device, host un-letterbox
            n = int(counts[i])
            if n == 0:
                continue
            if n <= k and ws.head_corrected:
                output[i] = head[i, 1:1 + n * 6].reshape(n, 6).copy()
                continue
            det = head[i, 1:1 + n * 6].reshape(n, 6).copy() if n <= k else ws.det[i, :n].cpu().numpy()
            box_xy, box_wh = (det[:, 0:2] + det[:, 2:4]) / 2, det[:, 2:4] - det[:, 0:2]
            det[:, :4] = self.yolo_correct_boxes(box_xy, box_wh, input_shape, shapes[i] if per_image else shapes,
                                                 letterbox_image)
            output[i] = det
        return output

    def non_max_suppression(self, prediction, num_classes, input_shape, image_shape, letterbox_image, conf_thres=0.5,
                            nms_thres=0.4):
        if num_classes != self.num_classes:
            raise ValueError("num_classes %r differs from the DecodeBox's %r" % (num_classes, self.num_classes))
        ws = self.nms_device(prediction, conf_thres, nms_thres)
        self.start_fetch(ws, input_shape, image_shape, letterbox_image)   # un-letterbox on the device, one D2H copy
        return self.fetch_detections(ws, input_shape, image_shape, letterbox_image)
