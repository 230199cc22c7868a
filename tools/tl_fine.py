import ctypes as C, os, sys
ROOT = "/root/repo"
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
    t = buf[6]; n = int((t != 0).sum()); t = t[:n] - buf[7, 0]
    print("==", name, "fine marks", n)
    for k in range(0, min(n, 64), 4):
        a = t[k:k + 4]
        if len(a) == 4:
            print("  iter %2d: top %6d | ld_wait %4d | compute %4d | ldtm+emit %4d | next top +%d" % (k // 4, a[0], a[1] - a[0], a[2] - a[1], a[3] - a[2], (t[k + 4] - a[3]) if k + 4 < n else -1))
