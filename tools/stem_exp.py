import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dcfa-yolo_b200"), ROOT):
    sys.path.insert(0, p)
import torch
import bench
from dcfa_b200 import _lib, abi, plan as P
dev = torch.device("cuda:0")
net = bench.build_model("s", 640, dev)
eng = net._engine(32, 640, 640, dev)
rgb = torch.rand(32, 3, 640, 640, device=dev); nir = torch.rand(32, 3, 640, 640, device=dev)
eng.run(rgb, nir); torch.cuda.synchronize()
st = torch.cuda.current_stream(dev)
i = eng.plan.op_names.index("stem")
op1 = (abi.Op * 1)(eng.plan.ops[i])
def run():
    _lib.check(_lib.lib.dcfa_run_ops(op1, 1, eng._bufs, P.NUM_BUFS, C.c_void_p(st.cuda_stream)))
def timed(pre=None):
    if pre: pre()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); run(); e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
print("plain x6      ", ["%.3f" % timed() for _ in range(6)])
big = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
print("after 256MB memset", ["%.3f" % timed(lambda: big.zero_()) for _ in range(4)])
print("after full fwd ", ["%.3f" % timed(lambda: eng.run(rgb, nir)) for _ in range(4)])
print("after sleep 0.2", ["%.3f" % timed(lambda: time.sleep(0.2)) for _ in range(3)])
# back-to-back x10 in one timing
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(10): run()
e1.record(st); torch.cuda.synchronize()
print("10 back-to-back: %.3f ms each" % (e0.elapsed_time(e1) / 10))
rgb2 = torch.rand(32, 3, 640, 640, device=dev)
print("plain again    ", ["%.3f" % timed() for _ in range(3)])
