"""dcfa_nms parity: kept indices bit-exact against (a) torchvision.ops.nms on the same device class
(CUDA op, iou_mode 1), (b) torchvision's CPU op (iou_mode 0) and (c) the C oracle in both modes."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_dcfa_nms(pred_gpu, conf, thr, mode):
    from dcfa_b200 import _lib
    b, a, row = pred_gpu.shape
    nc = row - 4
    det = torch.zeros(b, a, 6, device=pred_gpu.device)
    idx = torch.zeros(b, a, dtype=torch.int32, device=pred_gpu.device)
    cnt = torch.zeros(b, dtype=torch.int32, device=pred_gpu.device)
    cand = torch.zeros(b, dtype=torch.int32, device=pred_gpu.device)
    ws_bytes = _lib.lib.dcfa_nms_workspace_bytes(b, a)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=pred_gpu.device)
    _lib.check(_lib.lib.dcfa_nms(pred_gpu.data_ptr(), b, a, nc, conf, thr, mode, det.data_ptr(), idx.data_ptr(),
                                 cnt.data_ptr(), cand.data_ptr(), ws.data_ptr(), ws_bytes,
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return det.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy(), cand.cpu().numpy()


def reference_nms_torchvision(pred, conf, thr, device):
    """The reference's loop (utils/utils_bbox.py:92-168) calling the real torchvision op on `device`;
    returns per image (kept anchor indices, det rows)."""
    from torchvision.ops import nms
    pred = pred.clone().to(device)
    corner = pred.new(pred.shape)
    corner[:, :, 0] = pred[:, :, 0] - pred[:, :, 2] / 2
    corner[:, :, 1] = pred[:, :, 1] - pred[:, :, 3] / 2
    corner[:, :, 2] = pred[:, :, 0] + pred[:, :, 2] / 2
    corner[:, :, 3] = pred[:, :, 1] + pred[:, :, 3] / 2
    pred[:, :, :4] = corner[:, :, :4]
    res = []
    for image_pred in pred:
        class_conf, class_pred = torch.max(image_pred[:, 4:], 1, keepdim=True)
        mask = class_conf[:, 0] >= conf
        anchors = torch.nonzero(mask)[:, 0]
        det = torch.cat((image_pred[mask][:, :4], class_conf[mask].float(), class_pred[mask].float()), 1)
        keep_idx, keep_det = [], []
        for c in det[:, -1].cpu().unique():
            sel = det[:, -1] == c.to(device)
            dc, ac = det[sel], anchors[sel]
            keep = nms(dc[:, :4], dc[:, 4], thr)
            keep_idx.append(ac[keep])
            keep_det.append(dc[keep])
        if keep_idx:
            res.append((torch.cat(keep_idx).cpu().numpy(), torch.cat(keep_det).cpu().numpy()))
        else:
            res.append((np.zeros(0, np.int64), np.zeros((0, 6), np.float32)))
    return res


def make_pred(b, a, nc, seed, kind):
    g = torch.Generator().manual_seed(seed)
    if kind == "random":
        cxcy = torch.rand(b, a, 2, generator=g)
        wh = torch.rand(b, a, 2, generator=g) * 0.3 + 0.01
        cls = torch.rand(b, a, nc, generator=g)
    elif kind == "clustered":  # heavy overlap, many ties in score, quantised coordinates -> IoU exactly at thresholds
        cxcy = torch.randint(0, 12, (b, a, 2), generator=g).float() / 16 + 0.1
        wh = torch.randint(1, 5, (b, a, 2), generator=g).float() / 8
        cls = torch.randint(0, 8, (b, a, nc), generator=g).float() / 8
    elif kind == "degenerate":  # zero-area boxes and identical boxes
        cxcy = torch.randint(0, 4, (b, a, 2), generator=g).float() / 4
        wh = torch.randint(0, 2, (b, a, 2), generator=g).float() / 4
        cls = torch.full((b, a, nc), 0.75)
    return torch.cat((cxcy, wh, cls), 2).contiguous()


CASES = [
    (2, 8400, 1, "random", 0.5, 0.3),
    (2, 8400, 1, "random", 0.001, 0.5),
    (3, 2100, 4, "random", 0.3, 0.45),
    (2, 3000, 3, "clustered", 0.25, 0.5),
    (2, 3000, 1, "clustered", 0.125, 1.0 / 3.0),
    (2, 500, 2, "degenerate", 0.5, 0.4),
    (1, 1, 1, "random", 0.0, 0.5),
    (2, 37, 1, "random", 2.0, 0.5),       # nothing passes the confidence filter
    (1, 20000, 1, "random", 0.9, 0.5),    # 16 keys per thread in the register/shuffle sort
    (1, 33600, 1, "random", 0.99, 0.5),   # 1280x1280 anchors, few candidates: shared-memory boxes with cap < A
    (1, 33600, 1, "random", 0.05, 0.6),   # > 16384 candidates: global-memory sort + global-scratch greedy paths
    (2, 1500, 2, "random", 0.2, 0.5),     # 2 keys per thread
    (1, 4000, 1, "clustered", 0.125, 0.5),  # 4 keys per thread, many score ties
]


@pytest.mark.parametrize("b,a,nc,kind,conf,thr", CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_nms_vs_torchvision_and_oracle(cuda, b, a, nc, kind, conf, thr, mode):
    from oracle import nms as onms
    pred = make_pred(b, a, nc, 100 + a + nc, kind)
    pg = pred.clone().to(cuda)
    det, idx, cnt, cand = run_dcfa_nms(pg, conf, thr, mode)
    # (1) in-place xywh -> xyxy side effect (utils/utils_bbox.py:97)
    exp_corner = torch.stack((pred[..., 0] - pred[..., 2] / 2, pred[..., 1] - pred[..., 3] / 2,
                              pred[..., 0] + pred[..., 2] / 2, pred[..., 1] + pred[..., 3] / 2), -1)
    assert torch.equal(pg[..., :4].cpu(), exp_corner)
    # (2) real torchvision op on the matching device class
    ref = reference_nms_torchvision(pred, conf, thr, cuda if mode == 1 else torch.device("cpu"))
    for i in range(b):
        ridx, rdet = ref[i]
        assert cnt[i] == len(ridx), "image %d: kept %d vs torchvision %d" % (i, cnt[i], len(ridx))
        assert np.array_equal(idx[i, :cnt[i]].astype(np.int64), ridx), "image %d kept indices differ" % i
        assert np.array_equal(det[i, :cnt[i]], rdet.astype(np.float32))
    # (3) the C oracle
    pn = pred.numpy().copy()
    odet, oidx, ocnt, ocand = onms.nms_raw(pn, conf, thr, mode)
    assert np.array_equal(cnt, ocnt) and np.array_equal(cand, ocand)
    for i in range(b):
        assert np.array_equal(idx[i, :cnt[i]], oidx[i, :cnt[i]])
        assert np.array_equal(det[i, :cnt[i]], odet[i, :cnt[i]])


def test_pipelined_batches_on_one_decodebox_do_not_overwrite_each_other(cuda):
    """INTEGRATION.md's pipeline on a single DecodeBox: nms_device + start_fetch of batch A, then of batch B (same shape),
    and only then fetch_detections(A) -- A's rows must still be A's (the workspaces rotate through DecodeBox.ring)."""
    from utils.utils_bbox import DecodeBox
    dec = DecodeBox(2, (64, 64))
    shape = np.array([64, 64])
    preds = [make_pred(2, 600, 2, seed, "random") for seed in (1, 2, 3)]
    want = [DecodeBox(2, (64, 64)).non_max_suppression(p.clone().to(cuda), 2, [64, 64], shape, True, 0.3, 0.45) for p in preds]
    ws = []
    for p in preds[:2]:
        w = dec.nms_device(p.clone().to(cuda), 0.3, 0.45)
        dec.start_fetch(w)
        ws.append(w)
    assert ws[0] is not ws[1]
    got0 = dec.fetch_detections(ws[0], [64, 64], shape, True)
    w2 = dec.nms_device(preds[2].clone().to(cuda), 0.3, 0.45)
    dec.start_fetch(w2)
    got1 = dec.fetch_detections(ws[1], [64, 64], shape, True)
    got2 = dec.fetch_detections(w2, [64, 64], shape, True)
    for got, ref in zip((got0, got1, got2), want):
        for a, b in zip(got, ref):
            assert (a is None) == (b is None)
            if a is not None:
                np.testing.assert_array_equal(a, b)
    # more batches in flight than the ring holds is an error, not silent corruption
    dec2 = DecodeBox(2, (64, 64))
    dec2.ring = 2
    for i in range(2):
        dec2.start_fetch(dec2.nms_device(preds[i].clone().to(cuda), 0.3, 0.45))
    with pytest.raises(RuntimeError, match="in flight"):
        dec2.nms_device(preds[2].clone().to(cuda), 0.3, 0.45)
