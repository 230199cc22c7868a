#!/usr/bin/env python
"""Run the validation-loss criterion (dcfa_yolo_loss) at the bench batch for ncu / timing:
    python tools/prof_loss.py [--batch 32 --size 640 --targets 10 --iters 20]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dcfa-yolo_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402

from nets.yolo_training import Loss  # noqa: E402
from oracle import loss as OL  # noqa: E402


class M:
    stride = torch.tensor([8., 16., 32.])
    num_classes = 1
    no = 65
    reg_max = 16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--targets", type=int, default=10)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    feats, targets = OL.synth_case(seed=105, B=a.batch, nc=1, hw0=(a.size // 8, a.size // 8), n_targets=a.targets)
    maps = [torch.from_numpy(f).to(dev) for f in feats]
    tgt = torch.from_numpy(targets)
    crit = Loss(M())
    for _ in range(3):
        crit(maps, tgt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.iters):
        v = crit(maps, tgt)
    e1.record()
    torch.cuda.synchronize()
    print("loss %.5f  %.4f ms per call (public API, host included)" % (float(v), e0.elapsed_time(e1) / a.iters))
    # device time alone: targets staged once, the four launches enqueued back to back
    gt = crit.preprocess(tgt, a.batch, [a.size] * 4)
    gt_dev = torch.from_numpy(gt).to(dev)
    out = torch.empty(8, dtype=torch.float32, device=dev)
    for _ in range(3):
        crit.launch(maps, gt_dev, gt.shape[1], out)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.iters):
        crit.launch(maps, gt_dev, gt.shape[1], out)
    e1.record()
    torch.cuda.synchronize()
    print("launch only %.4f ms per call (4 launches, targets resident, G = %d)  loss %.5f" % (
        e0.elapsed_time(e1) / a.iters, gt.shape[1], float(out[3])))

if __name__ == "__main__":
    main()
