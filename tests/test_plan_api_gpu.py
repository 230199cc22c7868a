"""The C-side plan object (include/dcfa_b200.h: dcfa_plan_create / _run / _load / _forward) on a B200:
  * a host-built op list run through dcfa_plan_* via ctypes equals dcfa_run_ops bit for bit, also after the caller moves
    its input / output buffers (only the ops touching them are re-prepared), and launches the same number of kernels;
  * a plan FILE written by Plan.save and executed by a C program that links only the header, the library and the CUDA
    runtime (examples/plan_demo.c) reproduces YoloBody.forward's dbox / cls bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from test_forward_gpu import build_model
from test_oracle_cpu import golden_state_dict, load_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _alloc_io(plan, device):
    b = plan.B
    x = [torch.zeros(b, plan.no, h, w, dtype=torch.float32, device=device) for (h, w) in plan.level_shapes]
    return x, torch.zeros(b, 4, plan.A, device=device), torch.zeros(b, plan.nc, plan.A, device=device)


def test_plan_object_equals_run_ops_and_survives_moving_buffers(cuda):
    from dcfa_b200 import _lib
    from dcfa_b200 import plan as P
    from oracle import forward as O
    z, meta, keys = load_golden("s128_stress")
    sd = golden_state_dict(meta, keys)
    plan = P.Plan(sd, "s", 1, 2, 128, 128)
    blob = plan.blob_tensor.to(cuda)
    arena = torch.zeros(plan.arena_bytes + 256, dtype=torch.uint8, device=cuda)
    rgb, nir = (t.to(cuda) for t in O.synth_inputs(2, 128, 128, 5))
    st = torch.cuda.current_stream(cuda).cuda_stream

    def bufs_for(rgb, nir, x, dbox, cls, bound_only=False):
        a = (C.c_void_p * P.NUM_BUFS)()
        a[P.BUF_BLOB], a[P.BUF_ARENA] = blob.data_ptr(), arena.data_ptr()
        if not bound_only:
            a[P.BUF_RGB], a[P.BUF_NIR] = rgb.data_ptr(), nir.data_ptr()
            a[P.BUF_X0], a[P.BUF_X1], a[P.BUF_X2] = (t.data_ptr() for t in x)
            a[P.BUF_DBOX], a[P.BUF_CLS] = dbox.data_ptr(), cls.data_ptr()
        return a

    x0, d0, c0 = _alloc_io(plan, cuda)
    n0 = _lib.launch_count()
    _lib.check(_lib.lib.dcfa_run_ops(plan.op_array, len(plan.ops), bufs_for(rgb, nir, x0, d0, c0), P.NUM_BUFS, C.c_void_p(st)))
    eager_launches = _lib.launch_count() - n0
    torch.cuda.synchronize()

    handle = C.c_void_p()
    _lib.check(_lib.lib.dcfa_plan_create(plan.op_array, len(plan.ops), bufs_for(None, None, None, None, None, True), P.NUM_BUFS,
                                         C.byref(handle)))
    try:
        for trial in range(3):   # fresh input / output buffers on every trial: their ops are re-prepared, the rest replayed
            x1, d1, c1 = _alloc_io(plan, cuda)
            r2, n2 = rgb.clone(), nir.clone()
            n0 = _lib.launch_count()
            _lib.check(_lib.lib.dcfa_plan_run(handle, bufs_for(r2, n2, x1, d1, c1), P.NUM_BUFS, C.c_void_p(st)))
            assert _lib.launch_count() - n0 == eager_launches == _lib.lib.dcfa_plan_num_launches(handle)
            torch.cuda.synchronize()
            assert torch.equal(d1, d0) and torch.equal(c1, c0)
            for a, b in zip(x1, x0):
                assert torch.equal(a, b)
        # a bound buffer cannot be swapped behind the plan's back
        bad = bufs_for(rgb, nir, x0, d0, c0)
        bad[P.BUF_ARENA] = rgb.data_ptr()
        assert _lib.lib.dcfa_plan_run(handle, bad, P.NUM_BUFS, C.c_void_p(st)) < 0
        assert b"bound" in _lib.lib.dcfa_last_error()
    finally:
        _lib.lib.dcfa_plan_destroy(handle)


@pytest.mark.parametrize("u8,plane", [(False, False), (True, True)])
def test_c_only_caller_runs_forward_from_a_plan_file(cuda, tmp_path, u8, plane):
    from dcfa_b200 import plan as P
    from oracle import forward as O
    demo = os.path.join(ROOT, "dcfa-yolo_b200", "lib", "plan_demo")
    assert os.path.exists(demo), "build it with `make -C dcfa-yolo_b200/csrc` (graft build())"
    z, meta, keys = load_golden("n96_default_b3")
    sd = golden_state_dict(meta, keys)
    net = build_model(meta, sd, cuda)
    b, h, w, nc = 2, 96, 96, meta["nc"]
    if u8:
        g = torch.Generator().manual_seed(4)
        rgb = torch.randint(0, 256, (b, h, w, 3), generator=g, dtype=torch.uint8)
        nir = torch.randint(0, 256, (b, h, w), generator=g, dtype=torch.uint8)
    else:
        rgb, nir = O.synth_inputs(b, h, w, 9)
    out = net(rgb.to(cuda), nir.to(cuda))
    torch.cuda.synchronize()
    P.Plan(sd, "n", nc, b, h, w, input_u8=u8, depth_plane=plane).save(str(tmp_path / "plan.bin"))
    rgb.numpy().tofile(str(tmp_path / "rgb.bin"))
    nir.numpy().tofile(str(tmp_path / "dep.bin"))
    r = subprocess.run([demo, str(tmp_path / "plan.bin"), str(tmp_path / "rgb.bin"), str(tmp_path / "dep.bin"),
                        str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(str(tmp_path / "out.bin"), dtype=np.float32)
    a = out[0].shape[-1]
    assert np.array_equal(got[:b * 4 * a].reshape(b, 4, a), out[0].cpu().numpy())
    assert np.array_equal(got[b * 4 * a:].reshape(b, nc, a), out[1].cpu().numpy())
