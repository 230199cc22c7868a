"""The loss oracle (oracle/loss.py, numpy) against the goldens the REAL reference criterion produced
(oracle/make_golden_loss.py -> tests/golden/loss_*.npz): total, the three gained components, the assigner's foreground
mask and ground-truth index per anchor.  Also: the C-ABI argument checks of dcfa_yolo_loss (no compute without a GPU) and
the host-side target bookkeeping of the drop-in Loss against the oracle's restatement of Loss.preprocess."""
import ast
import glob
import os

import numpy as np
import pytest
import torch

from oracle import loss as OL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "loss_*.npz")))


def load_case(path):
    g = np.load(path)
    kw = ast.literal_eval(str(g["args"]))
    feats, targets = OL.synth_case(**kw)
    assert np.array_equal(targets, g["targets"])
    fg = np.unpackbits(g["fg"])[:int(np.prod(g["fg_shape"]))].reshape(g["fg_shape"]).astype(bool)
    return kw, feats, targets, g, fg


def test_goldens_present():
    assert len(GOLDEN) == 6


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_criterion(path):
    kw, feats, targets, g, fg = load_case(path)
    det = {}
    out = OL.loss_forward(feats, targets, nc=kw["nc"], details=det)
    assert np.array_equal(det["fg"], fg)
    assert np.array_equal(det["gt_idx"][fg], g["gt_idx"].astype(np.int64)[fg])
    assert abs(float(det["target_scores_sum"]) - float(g["tss"])) <= 1e-5 * float(g["tss"])
    np.testing.assert_allclose(out[:3], g["parts"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(out[3], g["total"], rtol=2e-5, atol=1e-5)


def test_loss_abi_argument_checks():
    from dcfa_b200 import _lib
    lib = _lib.lib
    assert lib.dcfa_loss_workspace_bytes(2, 8400, 1, 6) > 2 * 8400 * 24
    assert lib.dcfa_loss_workspace_bytes(0, 8400, 1, 6) == -1
    assert lib.dcfa_yolo_loss(None, None, None, 1, 1, None, None, None, 0, None, None, 0, None) == -1
    assert b"yolo_loss" in lib.dcfa_last_error()


def test_dropin_preprocess_matches_oracle():
    from nets.yolo_training import Loss

    class M:
        stride = torch.tensor([8., 16., 32.])
        num_classes = 3
        no = 67
        reg_max = 16
    crit = Loss(M())
    _, targets = OL.synth_case(seed=5, B=4, nc=3, hw0=(40, 56), n_targets=3, tiny=2, empty_image=2)
    targets = targets[np.random.RandomState(0).permutation(len(targets))]   # rows of one image need not be adjacent
    scale = np.array([448., 320., 448., 320.], np.float32)
    mine = crit.preprocess(torch.from_numpy(targets), 4, scale)
    assert np.array_equal(mine, OL.preprocess(targets, 4, scale))
    assert crit.preprocess(torch.zeros(0, 6), 4, scale).shape == (4, 0, 5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        crit([torch.zeros(1, 67, 8, 8), torch.zeros(1, 67, 4, 4), torch.zeros(1, 67, 2, 2)], torch.zeros(0, 6))
