"""TEST INFRASTRUCTURE -- ctypes front end of oracle/nms_oracle.c (C restatement of the reference's
DecodeBox.non_max_suppression, utils/utils_bbox.py:87-174, incl. torchvision.ops.nms) plus a pure-numpy
restatement of the host-side un-letterbox (utils/utils_bbox.py:60-85, :170-173)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdcfa_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.dcfa_oracle_nms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.dcfa_oracle_nms.restype = None
    return _lib


def nms_raw(pred, conf_thres, nms_thres, iou_mode):
    """pred: float32 [B,A,4+nc] (cx,cy,w,h,cls...), modified in place to xyxy like the reference.
    Returns (det [B,A,6], idx [B,A], cnt [B], cand [B])."""
    assert pred.dtype == np.float32 and pred.flags["C_CONTIGUOUS"] and pred.ndim == 3
    b, a, row = pred.shape
    det = np.zeros((b, a, 6), np.float32)
    idx = np.zeros((b, a), np.int32)
    cnt = np.zeros(b, np.int32)
    cand = np.zeros(b, np.int32)
    _load().dcfa_oracle_nms(pred.ctypes.data, b, a, row - 4, float(conf_thres), float(nms_thres), int(iou_mode),
                            det.ctypes.data, idx.ctypes.data, cnt.ctypes.data, cand.ctypes.data)
    return det, idx, cnt, cand


def yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image):
    """utils/utils_bbox.py:60-85: undo the letterbox, return rows (y1, x1, y2, x2) in original pixels."""
    box_yx = box_xy[..., ::-1]
    box_hw = box_wh[..., ::-1]
    input_shape = np.array(input_shape)
    image_shape = np.array(image_shape)
    if letterbox_image:
        new_shape = np.round(image_shape * np.min(input_shape / image_shape))
        offset = (input_shape - new_shape) / 2. / input_shape
        scale = input_shape / new_shape
        box_yx = (box_yx - offset) * scale
        box_hw = box_hw * scale
    mins = box_yx - box_hw / 2.
    maxs = box_yx + box_hw / 2.
    boxes = np.concatenate([mins[..., 0:1], mins[..., 1:2], maxs[..., 0:1], maxs[..., 1:2]], axis=-1)
    return boxes * np.concatenate([image_shape, image_shape], axis=-1)


def non_max_suppression(pred, input_shape, image_shape, letterbox_image, conf_thres=0.5, nms_thres=0.4, iou_mode=0):
    """Full restatement of DecodeBox.non_max_suppression: list of None | float32 [n_i,6] (y1,x1,y2,x2,conf,cls)."""
    det, idx, cnt, _ = nms_raw(pred, conf_thres, nms_thres, iou_mode)
    out = []
    for b in range(pred.shape[0]):
        if cnt[b] == 0:
            out.append(None)
            continue
        d = det[b, :cnt[b]].copy()
        xy, wh = (d[:, 0:2] + d[:, 2:4]) / 2, d[:, 2:4] - d[:, 0:2]
        d[:, :4] = yolo_correct_boxes(xy, wh, input_shape, image_shape, letterbox_image)
        out.append(d)
    return out
