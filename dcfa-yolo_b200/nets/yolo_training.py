"""The pieces of the reference's nets/yolo_training.py that sit on or next to the inference path:

* weights_init (reference nets/yolo_training.py:480-498, called from YoloBody.__init__ at nets/yolo_mul.py:393-394);
* Loss (reference :323-430), FORWARD ONLY: the criterion the reference's validation loop applies under no_grad to the
  eval-mode outputs of YoloBody.forward (utils/utils_fit_mul.py:78-92) -- decode, task-aligned assigner, BCE / CIoU /
  DFL terms -- as four CUDA launches behind `dcfa_yolo_loss` (csrc/loss.cu).  The returned tensor
  carries no autograd graph: backward, EMA and LR schedules are training (N3) and out of scope."""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

_INITS = {
    'normal': lambda w, gain: nn.init.normal_(w, 0.0, gain),
    'xavier': lambda w, gain: nn.init.xavier_normal_(w, gain=gain),
    'kaiming': lambda w, gain: nn.init.kaiming_normal_(w, a=0, mode='fan_in'),
    'orthogonal': lambda w, gain: nn.init.orthogonal_(w, gain=gain),
}


def weights_init(net, init_type='normal', init_gain=0.02):
    """Every module whose class name contains 'Conv' and that owns a `weight` gets `init_type`; every
    BatchNorm2d gets weight ~ N(1, 0.02), bias = 0.  Same selection rule and distributions as the reference."""
    if init_type not in _INITS:
        raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
    fill = _INITS[init_type]
    for m in net.modules():
        name = type(m).__name__
        if 'Conv' in name and getattr(m, 'weight', None) is not None:
            fill(m.weight.data, init_gain)
        elif 'BatchNorm2d' in name:
            nn.init.normal_(m.weight.data, 1.0, 0.02)
            nn.init.constant_(m.bias.data, 0.0)
    print('initialize network with %s type' % init_type)


class Loss:
    """Drop-in for the reference criterion's forward value: `Loss(model)(outputs, batch)` returns the 0-dim tensor
    box * 7.5 + cls * 0.5 + dfl * 1.5 on the device of the head maps (reference :371-430).  `outputs` is what
    YoloBody.forward returns (or just its list of three head maps); `batch` holds one row per ground-truth box,
    (image index, class, cx, cy, w, h) with the box normalised to [0, 1], on the host or on the device.
    `last` keeps the device vector (box, cls, dfl, sum, target_scores_sum, foreground anchors, 0, 0) of the latest call."""

    def __init__(self, model):
        from dcfa_b200 import _lib   # raises if the CUDA library is missing: there is no CPU path
        self._lib = _lib
        self.stride = [float(s) for s in model.stride]
        self.nc = int(model.num_classes)
        self.no = int(model.no)
        self.reg_max = int(model.reg_max)
        if self.reg_max != 16 or self.no != self.nc + 64 or len(self.stride) != 3:
            raise ValueError("Loss: the device criterion is built for reg_max = 16 and three levels (got reg_max %d, no %d, "
                             "%d strides)" % (self.reg_max, self.no, len(self.stride)))
        self.last = None
        self._ws = {}
        self._ring = None
        self._slot = 0

    def preprocess(self, targets, batch_size, scale_tensor):
        """Reference :342-360 on the host (index bookkeeping over a few rows, numpy float32 -- the same IEEE operations):
        (n, 6) -> (B, G, 5) rows (class, x1, y1, x2, y2) in input pixels, zero padded to the largest per-image count,
        order of appearance kept.  Returns a numpy array (a view of `out` when given)."""
        t = targets.detach().to("cpu", torch.float32).numpy().reshape(-1, 6) if isinstance(targets, torch.Tensor) else \
            np.asarray(targets, np.float32).reshape(-1, 6)
        if t.shape[0] == 0:
            return np.zeros((batch_size, 0, 5), np.float32)
        img = t[:, 0]
        _, counts = np.unique(img, return_counts=True)
        G = int(counts.max())
        out = np.zeros((batch_size, G, 5), np.float32)
        order = np.argsort(img, kind="stable")
        key = img[order]
        first = np.searchsorted(key, key, side="left")            # position of the first row of the same image
        rank = np.arange(len(key)) - first
        ok = (key >= 0) & (key < batch_size) & (key == np.floor(key))   # `i == j` for j in range(batch_size)
        out[key[ok].astype(np.int64), rank[ok]] = t[order[ok], 1:]
        xywh = out[..., 1:5] * np.asarray(scale_tensor, np.float32)
        half_w, half_h = xywh[..., 2] / np.float32(2), xywh[..., 3] / np.float32(2)
        out[..., 1:5] = np.stack((xywh[..., 0] - half_w, xywh[..., 1] - half_h, xywh[..., 0] + half_w, xywh[..., 1] + half_h), -1)
        return out

    def _stage(self, gt, dev):
        """Pinned staging ring for the padded targets: slot i is reused only after the copy that last read it has finished."""
        n = gt.size
        if self._ring is None or self._ring[0][0].numel() < n or self._ring[0][1].device != dev:
            cap = max(2 * n, 4096)
            self._ring = [(torch.empty(cap, dtype=torch.float32).pin_memory(), torch.empty(cap, dtype=torch.float32, device=dev),
                           torch.cuda.Event()) for _ in range(4)]
            self._slot = 0
        host, devbuf, ev = self._ring[self._slot]
        self._slot = (self._slot + 1) % len(self._ring)
        ev.synchronize()
        host[:n].numpy()[:] = gt.reshape(-1)
        devbuf[:n].copy_(host[:n], non_blocking=True)
        ev.record(torch.cuda.current_stream(dev))
        return devbuf

    def __call__(self, preds, batch):
        feats = preds[2] if isinstance(preds, tuple) else preds
        if len(feats) != 3:
            raise ValueError("Loss: expected three head maps, got %d" % len(feats))
        dev = feats[0].device
        if dev.type != "cuda":
            raise RuntimeError("Loss: the head maps must be CUDA tensors (there is no CPU path)")
        B = int(feats[0].shape[0])
        maps = []
        for f in feats:
            if f.dtype != torch.float32 or f.dim() != 4 or f.shape[0] != B or f.shape[1] != self.no:
                raise ValueError("Loss: head maps must be fp32 [B, %d, H, W]; got %s %s" % (self.no, f.dtype, tuple(f.shape)))
            maps.append(f.detach().contiguous())
        hw = [int(v) for f in maps for v in f.shape[2:]]
        A = sum(hw[2 * i] * hw[2 * i + 1] for i in range(3))
        # imgsz = feats[0].shape[2:] * stride[0] (:390); targets scaled by (w, h, w, h)
        imgsz = np.array(hw[:2], np.float32) * np.float32(self.stride[0])
        gt = self.preprocess(batch, B, imgsz[[1, 0, 1, 0]])
        G = int(gt.shape[1])
        if G and (gt[..., 0].min() < 0 or gt[..., 0].max() >= self.nc):
            raise ValueError("Loss: class labels must lie in [0, %d)" % self.nc)
        with torch.cuda.device(dev):
            gt_dev = self._stage(gt, dev) if G else None
            out = self.launch(maps, gt_dev, G)
        self.last = out
        return out[3]

    def launch(self, maps, gt_dev, G, out=None):
        """The four launches of dcfa_yolo_loss on the current stream (no host work beyond argument packing; CUDA-graph
        capturable when `out` and the workspace already exist).  maps: three contiguous fp32 head maps; gt_dev: device
        buffer holding (B, G, 5) padded targets.  Returns the device vector described in the class docstring."""
        dev = maps[0].device
        B = int(maps[0].shape[0])
        hw = [int(v) for f in maps for v in f.shape[2:]]
        A = sum(hw[2 * i] * hw[2 * i + 1] for i in range(3))
        lib = self._lib.lib
        need = int(lib.dcfa_loss_workspace_bytes(B, A, self.nc, G))
        key = (dev.index, need)
        ws = self._ws.get(key)
        if ws is None:
            self._ws.clear()
            ws = self._ws[key] = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        if out is None:
            out = torch.empty(8, dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._lib.check(lib.dcfa_yolo_loss(
            maps[0].data_ptr(), maps[1].data_ptr(), maps[2].data_ptr(), B, self.nc, (C.c_int32 * 6)(*hw),
            (C.c_float * 3)(*self.stride), gt_dev.data_ptr() if G else None, G, out.data_ptr(), ws.data_ptr(),
            ws.numel(), C.c_void_p(st)))
        return out
