#!/usr/bin/env python
"""Debug: per-role clock64 timeline of CTA 0 of one TMA conv op.
Needs a library built with `make -C dcfa-yolo_b200/csrc EXTRA=-DDCFA_TIMELINE` (never the shipped build)."""
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
    op1 = (abi.Op * 1)(eng.plan.ops[i])
    for _ in range(2):
        _lib.check(_lib.lib.dcfa_run_ops(op1, 1, eng.last_bufs, P.NUM_BUFS, C.c_void_p(st.cuda_stream)))
    torch.cuda.synchronize()
    buf = np.zeros((8, 2048), np.int64)
    _lib.lib.dcfa_debug_clear_timeline()
    _lib.check(_lib.lib.dcfa_run_ops(op1, 1, eng.last_bufs, P.NUM_BUFS, C.c_void_p(st.cuda_stream)))
    torch.cuda.synchronize()
    _lib.lib.dcfa_debug_read_timeline(buf.ctypes.data, buf.nbytes)
    kb = eng.plan.ops[i].k_blocks
    t0 = buf[0, 0]
    b = buf - t0
    print("== %s: k_blocks/tile %d" % (name, kb))
    print(" entry %d  pdl_wait passed %d  first producer issue %d  (cycles after kernel entry of CTA 0)" % (0, buf[7, 1] - buf[7, 0], buf[0, 0] - buf[7, 0]))
    for t in range(0, 4):   # the first tiles of CTA 0: launch latency, pipeline fill, epilogue drain
        k0 = t * kb
        if buf[0, k0] == 0 and t > 0:
            break
        g, lt = t & 1, t >> 1
        e = buf[4 + g, 4 * lt: 4 * lt + 4] - buf[7, 0]
        print(" first tiles: tile %d prod_issue[first,last] %d %d | full_ok[first,last] %d %d | mma issued[last] %d | epi start %d tfull %d done %d end %d" % (
            t, buf[0, k0] - buf[7, 0], buf[0, k0 + kb - 1] - buf[7, 0], buf[1, k0] - buf[7, 0], buf[1, k0 + kb - 1] - buf[7, 0],
            buf[2, k0 + kb - 1] - buf[7, 0], e[0], e[1], e[2], e[3]))
    T = 40  # look at tiles around T
    for t in range(T, T + 4):
        k0 = t * kb
        print(" tile %d: prod_issue[first,last] %d %d | mma tempty_ok %d full_ok[first,last] %d %d issued[last] %d" % (
            t, b[0, k0], b[0, k0 + kb - 1], b[3, t], b[1, k0], b[1, k0 + kb - 1], b[2, k0 + kb - 1]))
        g, lt = t & 1, t >> 1
        e = b[4 + g, 4 * lt: 4 * lt + 4]
        print("          epi(group %d): start %d tfull_ok %d chunks_done %d end %d  | wait %d work %d" % (g, e[0], e[1], e[2], e[3], e[1] - e[0], e[2] - e[1]))
    n = min(2048, 46 * kb)
    print(" producer interval/kb: median %d | mma issue interval/kb: median %d | full lag median %d" % (
        np.median(np.diff(b[0, :n])), np.median(np.diff(b[2, :n])), np.median(b[1, :n] - b[0, :n])))
