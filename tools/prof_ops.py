#!/usr/bin/env python
"""Run selected ops of the compiled forward plan in isolation (for ncu / timing).

    python tools/prof_ops.py --ops stem,dark2.0,dark2.1.pw1 [--phi s --batch 32 --size 640 --iters 3]

A full forward runs first (so every op sees real inputs), then each selected op is launched `iters` times.
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dcfa-yolo_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from dcfa_b200 import _lib, abi  # noqa: E402
from dcfa_b200 import plan as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="stem")
    ap.add_argument("--phi", default="s")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--u8", action="store_true", help="uint8 NHWC inputs (DCFA_STEM_FLAG_U8)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    net = bench.build_model(a.phi, a.size, dev)
    eng = net._engine(a.batch, a.size, a.size, dev, input_u8=a.u8)
    if a.u8:
        rgb = torch.randint(0, 256, (a.batch, a.size, a.size, 3), dtype=torch.uint8, device=dev)
        nir = torch.randint(0, 256, (a.batch, a.size, a.size, 3), dtype=torch.uint8, device=dev)
    else:
        rgb = torch.rand(a.batch, 3, a.size, a.size, device=dev)
        nir = torch.rand(a.batch, 3, a.size, a.size, device=dev)
    eng.run(rgb, nir)
    torch.cuda.synchronize()
    st = torch.cuda.current_stream(dev)
    want = a.ops.split(",")
    torch.cuda.profiler.start()   # ncu --profile-from-start off: only the selected ops are captured
    for name in want:
        count = 1
        if "+" in name:   # "dark2.1.pw1+3": three consecutive records in one call (lets the dispatcher fuse them)
            name, cnt = name.split("+")
            count = int(cnt)
        idx = [i for i, n in enumerate(eng.plan.op_names) if n == name]
        if not idx:
            print("no op named", name, "; have:", eng.plan.op_names)
            continue
        i = idx[0]
        op1 = (abi.Op * count)(*eng.plan.ops[i:i + count])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(a.iters):
            e0.record(st)
            _lib.check(_lib.lib.dcfa_run_ops(op1, count, eng.last_bufs, P.NUM_BUFS, C.c_void_p(st.cuda_stream)))
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("%-16s %s ms" % (name, ["%.4f" % t for t in ts]))
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
