#!/usr/bin/env python
"""Debug: clock64 timeline of CTA 0 of the halo-strip conv kernel (library built with `make EXTRA=-DDCFA_TIMELINE`).
    python tools/tl_strip.py head0.0 head0.cls1"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dcfa-yolo_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np, torch
import bench
from dcfa_b200 import _lib, abi, plan as P
dev = torch.device("cuda:0")
net = bench.build_model("s", 640, dev)
eng = net._engine(32, 640, 640, dev)
rgb = torch.rand(32, 3, 640, 640, device=dev); nir = torch.rand(32, 3, 640, 640, device=dev)
eng.run(rgb, nir); torch.cuda.synchronize()
st = torch.cuda.current_stream(dev)
for name in sys.argv[1:]:
    i = eng.plan.op_names.index(name)
    op = eng.plan.ops[i]
    op1 = (abi.Op * 1)(op)
    for _ in range(2):
        _lib.check(_lib.lib.dcfa_run_ops(op1, 1, eng.last_bufs, P.NUM_BUFS, C.c_void_p(st.cuda_stream)))
    torch.cuda.synchronize()
    buf = np.zeros((8, 4096), np.int64)
    _lib.lib.dcfa_debug_clear_strip_timeline()
    _lib.check(_lib.lib.dcfa_run_ops(op1, 1, eng.last_bufs, P.NUM_BUFS, C.c_void_p(st.cuda_stream)))
    torch.cuda.synchronize()
    _lib.lib.dcfa_debug_read_strip_timeline(buf.ctypes.data, buf.nbytes)
    cb = op.Cin // 64
    t0 = buf[5, 0]
    n = lambda r: int((buf[r] != 0).sum())
    print("== %s: cblocks %d, BN %d; marks: strips %d, W tiles %d, tiles %d" % (name, cb, op.BN, n(0), n(2), n(5)))
    nw = n(4)
    wi, wl, wm = buf[2, :nw] - t0, buf[3, :nw] - t0, buf[4, :nw] - t0
    print(" W tile: issue->landed-seen median %d | MMA k-block interval median %d | producer issue interval median %d" % (
        np.median(wl - wi), np.median(np.diff(wm)), np.median(np.diff(wi))))
    ns = n(1)
    si, sl = buf[0, :ns] - t0, buf[1, :ns] - t0
    print(" strip: issue->seen median %d | first strips issue %s seen %s" % (np.median(sl - si), si[:4].tolist(), sl[:4].tolist()))
    nt = n(5)
    ts = buf[5, :nt] - t0
    print(" tile starts (MMA) %s ... interval median %d" % (ts[:6].tolist(), np.median(np.diff(ts)) if nt > 1 else -1))
    ne = n(6) // 2
    e = (buf[6, :2 * ne] - t0).reshape(-1, 2)
    print(" epilogue (tfull, done) first 4: %s ; work median %d" % (e[:4].tolist(), np.median(e[:, 1] - e[:, 0]) if ne else -1))
    k = 9 * cb
    for t in (2, 3):
        if nw >= (t + 1) * k and nt > t and ns >= (t + 1) * cb:
            print(" tile %d boundary: last MMA of previous tile %d | tempty passed %d | strip seen %d | first W seen %d | first MMA issued %d | W issue times of first 3 tiles of this tile %s" % (
                t, wm[t * k - 1], ts[t], sl[t * cb], wl[t * k], wm[t * k], wi[t * k:t * k + 3].tolist()))
    print(" per-k-block MMA issue times of tile 2:", (wm[2 * k:3 * k] - wm[2 * k]).tolist() if nw >= 3 * k else "-")
    print(" W landed-seen minus MMA-ready gap (wait on W) tile 2:", (wl[2 * k:3 * k] - np.concatenate(([wm[2 * k - 1]], wm[2 * k:3 * k - 1]))).tolist() if nw >= 3 * k else "-")
