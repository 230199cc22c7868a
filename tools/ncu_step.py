#!/usr/bin/env python
"""Launch ONE step of the bench workload (forward units in plan order, then decode_box and NMS) inside a
cudaProfilerStart/Stop range, and write the manifest that names each launch, so an `ncu --profile-from-start off`
capture of this program can be joined with the plan:

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_step \
        python tools/ncu_step.py --manifest gpurun_out/r2_step_units.json
    ncu -i gpurun_out/r2_step.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_join.py raw.csv units.json
"""
import argparse
import ctypes as C
import json
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
    ap.add_argument("--phi", default="s")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--manifest", default="gpurun_out/step_units.json")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    net = bench.build_model(a.phi, a.size, dev)
    eng = net._engine(a.batch, a.size, a.size, dev)
    rgb = torch.rand(a.batch, 3, a.size, a.size, device=dev)
    nir = torch.rand(a.batch, 3, a.size, a.size, device=dev)
    from utils.utils_bbox import DecodeBox
    dec = DecodeBox(1, (a.size, a.size))
    out = eng.run(rgb, nir)
    full = (out[0], out[1], out[2], eng.anchors, eng.strides)
    dec.nms_device(dec.decode_box(full), bench.CONF, bench.IOU)
    torch.cuda.synchronize()
    st = torch.cuda.current_stream(dev)
    groups = bench.plan_units(eng.plan)
    arrays = [(abi.Op * cnt)(*eng.plan.ops[i0:i0 + cnt]) for (i0, cnt, _, _) in groups]
    units = []
    torch.cuda.profiler.start()
    for gi, (i0, cnt, kind, name) in enumerate(groups):
        n0 = _lib.lib.dcfa_launch_count()
        _lib.check(_lib.lib.dcfa_run_ops(arrays[gi], cnt, eng.last_bufs, P.NUM_BUFS, C.c_void_p(st.cuda_stream)))
        op = eng.plan.ops[i0]
        flops = 0
        if kind == "conv":
            flops = 2 * op.n_img * op.Ho * op.Wo * op.Cout * op.K_real
            if op.flags & abi.CONV_FLAG_DFL:
                flops = 2 * op.n_img * op.Ho * op.Wo * (64 * 64 + op.nc * (op.Cin - 64))
        units.append({"name": name, "kind": kind, "launches": int(_lib.lib.dcfa_launch_count() - n0), "flops": flops,
                      "bytes": int(bench.algorithmic_bytes(kind, op, cnt, a.batch, eng.plan,
                                                           extra_res=(kind == "ghost" and eng.plan.ops[i0 + 1].x2.buf >= 0)))})
    n0 = _lib.lib.dcfa_launch_count()
    y = dec.decode_box(full)
    units.append({"name": "decode_box", "kind": "decode", "launches": int(_lib.lib.dcfa_launch_count() - n0), "flops": 0,
                  "bytes": int(a.batch * eng.plan.A * (5 + eng.plan.nc) * 4 * 2)})
    n0 = _lib.lib.dcfa_launch_count()
    dec.nms_device(y, bench.CONF, bench.IOU)
    units.append({"name": "nms", "kind": "nms", "launches": int(_lib.lib.dcfa_launch_count() - n0), "flops": 0, "bytes": 0})
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    json.dump({"phi": a.phi, "batch": a.batch, "size": a.size, "units": units}, open(a.manifest, "w"), indent=1)
    print("units", len(units), "launches", sum(u["launches"] for u in units))


if __name__ == "__main__":
    main()
