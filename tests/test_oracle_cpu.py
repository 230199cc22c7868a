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


# ------------------------------------------------------------------------------------------------------
# forward / decode / NMS restatement vs golden vectors produced by the real reference (oracle/make_golden.py)
# ------------------------------------------------------------------------------------------------------
import glob
import json
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# model goldens (oracle/make_golden.py); facade_*.npz hold the host-helper goldens of oracle/make_golden_facade.py,
# loss_*.npz those of the validation-loss criterion (oracle/make_golden_loss.py, tests/test_loss_*.py)
GOLDEN_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz"))
                      if not os.path.basename(p).startswith(("facade_", "loss_")))


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    keys = json.load(open(os.path.join(GOLD, "state_dict_keys.json")))["%s_nc%d" % (meta["phi"], meta["nc"])]
    return z, meta, keys


def golden_state_dict(meta, keys):
    from oracle import forward as O
    template = {k: torch.empty(tuple(s), device="meta") for k, s in keys.items()}
    return O.synth_state_dict(template, meta["seed"], meta["mode"])


def compare_x_maps(z, x, atol, rtol):
    for i, xi in enumerate(x):
        xi = xi.detach().float().cpu().numpy()
        assert tuple(z["x%d_full_shape" % i]) == xi.shape
        g = z["x%d" % i]
        sub = xi if g.shape == xi.shape else xi[:, :, ::4, ::4]
        np.testing.assert_allclose(sub, g, atol=atol, rtol=rtol, err_msg="head map %d" % i)
        # checksum over EVERY element of the full-resolution map (the stored map may be a stride-4 subsample): the mean
        # signed error and the mean magnitude must agree well inside the per-element tolerance -- a systematic bias or a
        # wrong region outside the subsample grid shows up here
        s_ref, a_ref = z["x%d_sum" % i]
        x64 = xi.astype(np.float64)
        assert abs(x64.sum() - s_ref) <= xi.size * atol / 4 + rtol / 4 * a_ref, "head map %d: sum" % i
        assert abs(np.abs(x64).sum() - a_ref) <= xi.size * atol / 4 + rtol / 4 * a_ref, "head map %d: abs sum" % i


def test_golden_set_is_complete():
    assert set(GOLDEN_CASES) >= {"n640_default", "n640_stress", "s128_stress", "m64x96_stress_nc3", "l64_stress",
                                 "n96_default_b3", "s640_stress", "l1280_stress", "x64_stress"}


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_forward_decode_nms_match_reference_goldens(name):
    from oracle import forward as O
    from oracle import nms as onms
    z, meta, keys = load_golden(name)
    sd = golden_state_dict(meta, keys)
    rgb, nir = O.synth_inputs(meta["B"], meta["H"], meta["W"], meta["seed"] + 1000)
    out = O.yolo_forward(sd, meta["phi"], rgb, nir, meta["nc"])
    dbox, cls, x, anchors, strides = out
    np.testing.assert_allclose(dbox.numpy(), z["dbox"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(cls.numpy(), z["cls"], atol=2e-4, rtol=1e-4)
    compare_x_maps(z, x, 2e-4, 1e-4)
    assert np.array_equal(anchors.numpy(), z["anchors"]) and np.array_equal(strides.numpy(), z["strides"])
    # decode on the REFERENCE's own head outputs: 1e-5 (north_star tolerance for this stage)
    gold_out = (torch.from_numpy(z["dbox"]), torch.from_numpy(z["cls"]), None, torch.from_numpy(z["anchors"]),
                torch.from_numpy(z["strides"]))
    dec = O.decode_box(gold_out, (meta["H"], meta["W"]))
    np.testing.assert_allclose(dec.numpy(), z["decoded"], atol=1e-5, rtol=0)
    # NMS on the reference's decoded boxes: bit-exact rows
    res = onms.non_max_suppression(z["decoded"].copy(), [meta["H"], meta["W"]], np.array([meta["H"], meta["W"]]), True,
                                   meta["conf"], meta["iou"], iou_mode=0)
    for i, r in enumerate(res):
        g = z["nms%d" % i]
        if r is None:
            assert g.shape[0] == 0
        else:
            assert r.shape == g.shape
            np.testing.assert_array_equal(r[:, 4:], g[:, 4:])
            np.testing.assert_allclose(r[:, :4], g[:, :4], atol=1e-4, rtol=1e-6)
