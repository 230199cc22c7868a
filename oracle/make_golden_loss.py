"""TEST INFRASTRUCTURE -- generates tests/golden/loss_*.npz by running the REAL reference criterion
(`Loss`, `TaskAlignedAssigner`, `BboxLoss` imported from /root/reference/nets/yolo_training.py, read-only) on
deterministic synthetic head maps and targets.  Run in the authoring container:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_loss.py

Head maps are NOT stored: oracle.loss.synth_case regenerates them from the seed (numpy MT19937).  Stored per case: the
generator arguments, the target rows, the reference's loss (sum and the three gained components, taken from the
criterion's own sub-modules through forward hooks), the assigner's foreground mask and ground-truth index per anchor,
and target_scores_sum.  The script also checks that the unspecified tie order of torch.topk cannot have influenced a
stored result (oracle.loss docstring): the restatement is run with both tie orders and both must reproduce the
reference's assignment exactly.
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DCFA_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from nets.yolo_training import Loss  # noqa: E402  (reference)

from oracle import loss as OL  # noqa: E402

CASES = {
    # name: dict(seed, B, nc, hw0, n_targets, tiny, crowd, empty_image, spread)
    "loss_n640_b2": dict(seed=31, B=2, nc=1, hw0=(80, 80), n_targets=6),
    "loss_320_nc3_tiny_empty": dict(seed=32, B=3, nc=3, hw0=(40, 40), n_targets=4, tiny=3, empty_image=1),
    "loss_rect_crowd_nc2": dict(seed=33, B=2, nc=2, hw0=(32, 48), n_targets=2, crowd=8),
    "loss_rollout_70gts": dict(seed=34, B=2, nc=1, hw0=(40, 40), n_targets=70),
    "loss_no_targets": dict(seed=35, B=2, nc=1, hw0=(16, 16), n_targets=0),
    "loss_640_b4_peaky": dict(seed=36, B=4, nc=1, hw0=(80, 80), n_targets=9, tiny=2, crowd=3, spread=3.0),
}


class _Model:
    def __init__(self, nc):
        self.stride = torch.tensor([8., 16., 32.])
        self.num_classes = nc
        self.reg_max = 16
        self.no = nc + 64


def run_reference(feats, targets, nc):
    crit = Loss(_Model(nc))
    seen = {}
    crit.assigner.register_forward_hook(lambda m, i, o: seen.__setitem__("assign", o))
    crit.bbox_loss.register_forward_hook(lambda m, i, o: seen.__setitem__("bbox", o))
    # the BCE term is read back through a wrapper around the criterion's own module
    bce = crit.bce

    def bce_spy(x, t):
        out = bce(x, t)
        seen["bce_sum"] = out.sum()
        seen["tscore"] = t
        return out
    crit.bce = bce_spy
    tf = [torch.from_numpy(f) for f in feats]
    dummy = torch.zeros(1)
    with torch.no_grad():
        total = crit((dummy, dummy, tf, None, None), torch.from_numpy(targets))
    _, _, tscore, fg, gt_idx = seen["assign"]
    tss = max(float(tscore.sum()), 1.0)
    box = float(seen["bbox"][0]) * 7.5 if "bbox" in seen else 0.0
    dfl = float(seen["bbox"][1]) * 1.5 if "bbox" in seen else 0.0
    cls = float(seen["bce_sum"]) / tss * 0.5
    return dict(total=np.float32(total), parts=np.array([box, cls, dfl], np.float32), fg=fg.numpy().astype(bool),
                gt_idx=gt_idx.numpy().astype(np.int64), tss=np.float32(tss))


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, kw in CASES.items():
        feats, targets = OL.synth_case(**kw)
        ref = run_reference(feats, targets, kw["nc"])
        # tie-order independence of the stored vectors
        for tie in ("low", "high"):
            saved = OL.topk_lowest_index_first
            if tie == "high":
                OL.topk_lowest_index_first = lambda v, k: np.lexsort((-np.arange(v.size), -v.astype(np.float64)))[:k]
            det = {}
            mine = OL.loss_forward(feats, targets, nc=kw["nc"], details=det)
            OL.topk_lowest_index_first = saved
            assert np.array_equal(det["fg"], ref["fg"]), (name, tie, "foreground mask differs")
            assert np.array_equal(det["gt_idx"][ref["fg"]], ref["gt_idx"][ref["fg"]]), (name, tie, "gt index differs")
            assert abs(mine[3] - ref["total"]) <= 2e-5 * abs(ref["total"]) + 1e-5, (name, tie, mine, ref["total"])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"),
                            args=np.array(repr(kw)), targets=targets, total=ref["total"], parts=ref["parts"],
                            fg=np.packbits(ref["fg"]), fg_shape=np.array(ref["fg"].shape),
                            gt_idx=ref["gt_idx"].astype(np.int16), tss=ref["tss"])
        print("%-28s total %.6f parts %s fg %d tss %.4f  (oracle %.6f)" % (
            name, ref["total"], ref["parts"], ref["fg"].sum(), ref["tss"], mine[3]))


if __name__ == "__main__":
    main()
