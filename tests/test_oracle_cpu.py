"""CPU-side pinning of the oracle against the real dependency the reference calls (torchvision.ops.nms, CPU op)."""
import numpy as np
import pytest
import torch

from test_nms_gpu import CASES, make_pred, reference_nms_torchvision


@pytest.mark.parametrize("b,a,nc,kind,conf,thr", CASES)
def test_c_oracle_nms_matches_torchvision_cpu(b, a, nc, kind, conf, thr):
    from oracle import nms as onms
    pred = make_pred(b, a, nc, 100 + a + nc, kind)
    ref = reference_nms_torchvision(pred, conf, thr, torch.device("cpu"))
    det, idx, cnt, cand = onms.nms_raw(pred.numpy().copy(), conf, thr, 0)
    for i in range(b):
        ridx, rdet = ref[i]
        assert cnt[i] == len(ridx)
        assert np.array_equal(idx[i, :cnt[i]].astype(np.int64), ridx)
        assert np.array_equal(det[i, :cnt[i]], rdet.astype(np.float32))
