#!/usr/bin/env python
"""Run decode + NMS of the bench workload in isolation (for ncu / timing).

    python tools/prof_nms.py [--phi s --batch 32 --size 640 --iters 5]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dcfa-yolo_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from utils.utils_bbox import DecodeBox  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--phi", default="s")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    net = bench.build_model(a.phi, a.size, dev)
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(a.batch, 3, a.size, a.size, generator=g).to(dev)
    nir = torch.rand(a.batch, 3, a.size, a.size, generator=g).to(dev)
    dec = DecodeBox(1, (a.size, a.size))
    out = net(rgb, nir)
    y0 = dec.decode_box(out)
    torch.cuda.synchronize()
    st = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(a.iters):
        y = y0.clone()
        torch.cuda.synchronize()
        e0.record(st)
        ws = dec.nms_device(y, bench.CONF, bench.IOU)
        e1.record(st)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("nms ms:", ["%.4f" % t for t in ts], "cand", ws.cand[:4].tolist(), "kept", ws.cnt[:4].tolist())


if __name__ == "__main__":
    main()
