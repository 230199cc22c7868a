"""dcfa_yolo_loss (csrc/loss.cu) on the GPU: against the goldens of the REAL reference criterion, against the numpy
oracle on further seeded cases (ragged targets, empty images, crowded boxes, > 64 boxes per image, no targets, the
bench batch), against the reference criterion itself run on the same GPU (oracle/_ref, when staged), end to end behind
YoloBody.forward, and for run-to-run determinism.  Tolerance: 1e-4 relative on every loss component (fp32 sums in a
different order; atanf / powf / expf within a few ulp of torch's); foreground-anchor count exact."""
import contextlib
import glob
import io
import os
import sys

import numpy as np
import pytest
import torch

from oracle import loss as OL
from test_loss_cpu import GOLDEN, load_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL, ATOL = 1e-4, 1e-5


class _Model:
    def __init__(self, nc):
        self.stride = torch.tensor([8., 16., 32.])
        self.num_classes = nc
        self.no = nc + 64
        self.reg_max = 16


def run_device(feats, targets, nc, cuda, targets_on_device=False):
    from nets.yolo_training import Loss
    crit = Loss(_Model(nc))
    x = [torch.from_numpy(f).to(cuda) for f in feats]
    t = torch.from_numpy(targets)
    total = crit((None, None, x, None, None), t.to(cuda) if targets_on_device else t)
    out = crit.last.cpu().numpy()
    assert float(total) == out[3]
    return out


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_loss_matches_reference_goldens(cuda, path):
    kw, feats, targets, g, fg = load_case(path)
    out = run_device(feats, targets, kw["nc"], cuda)
    np.testing.assert_allclose(out[:3], g["parts"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out[3], g["total"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out[4], g["tss"], rtol=RTOL)
    assert int(out[5]) == int(fg.sum())


CASES = [
    dict(seed=101, B=1, nc=1, hw0=(80, 80), n_targets=1),
    dict(seed=102, B=5, nc=4, hw0=(24, 40), n_targets=5, tiny=4, crowd=5, empty_image=3),
    dict(seed=103, B=2, nc=80, hw0=(20, 20), n_targets=12),
    dict(seed=104, B=3, nc=2, hw0=(40, 40), n_targets=90, crowd=20),          # > 64 boxes: the reference's roll-out path
    dict(seed=105, B=32, nc=1, hw0=(80, 80), n_targets=8, tiny=2, crowd=2),   # the bench batch (BASELINE configs[1])
    dict(seed=106, B=2, nc=1, hw0=(160, 160), n_targets=10, spread=2.0),      # l@1280: A = 33 600
    dict(seed=107, B=2, nc=1, hw0=(16, 16), n_targets=0),
]


@pytest.mark.parametrize("kw", CASES, ids=["seed%d" % c["seed"] for c in CASES])
def test_loss_matches_oracle(cuda, kw):
    feats, targets = OL.synth_case(**kw)
    det = {}
    ref = OL.loss_forward(feats, targets, nc=kw["nc"], details=det)
    out = run_device(feats, targets, kw["nc"], cuda, targets_on_device=(kw["seed"] % 2 == 0))
    np.testing.assert_allclose(out[:4], ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out[4], det["target_scores_sum"], rtol=RTOL)
    assert int(out[5]) == int(det["fg"].sum())


def test_loss_is_deterministic_and_leaves_inputs_alone(cuda):
    from nets.yolo_training import Loss
    feats, targets = OL.synth_case(seed=108, B=4, nc=2, hw0=(40, 40), n_targets=7, crowd=6)
    x = [torch.from_numpy(f).to(cuda) for f in feats]
    keep = [t.clone() for t in x]
    crit = Loss(_Model(2))
    outs = []
    for _ in range(3):
        crit(x, torch.from_numpy(targets))
        outs.append(crit.last.cpu().numpy().copy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert all(torch.equal(a, b) for a, b in zip(x, keep))


def test_loss_matches_reference_criterion_on_this_gpu(cuda):
    """The reference's own Loss (staged under oracle/_ref by `make -C oracle ref`) run on CUDA tensors: same device class
    as the drop-in, so this also pins torch.topk's CUDA behaviour at the top-10 boundary on inputs with tiny boxes."""
    ref_root = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_root, "nets", "yolo_training.py")):
        pytest.skip("oracle/_ref not staged")
    import subprocess
    code = r"""
import sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %r); sys.path.insert(1, %r)
import numpy as np, torch
from nets.yolo_training import Loss
from oracle import loss as OL
class M:
    def __init__(s, nc):
        s.stride = torch.tensor([8., 16., 32.]); s.num_classes = nc; s.no = nc + 64; s.reg_max = 16
for kw in %r:
    feats, targets = OL.synth_case(**kw)
    x = [torch.from_numpy(f).cuda() for f in feats]
    with torch.no_grad():
        v = Loss(M(kw["nc"]))((x[0], x[0], x), torch.from_numpy(targets).cuda())
    print("REF %%d %%.9g" %% (kw["seed"], float(v)))
""" % (ref_root, ROOT, CASES[:5])
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    ref = {int(l.split()[1]): float(l.split()[2]) for l in res.stdout.splitlines() if l.startswith("REF ")}
    assert len(ref) == 5
    for kw in CASES[:5]:
        feats, targets = OL.synth_case(**kw)
        out = run_device(feats, targets, kw["nc"], cuda)
        assert abs(out[3] - ref[kw["seed"]]) <= RTOL * abs(ref[kw["seed"]]) + ATOL, (kw, out[3], ref[kw["seed"]])


def test_validation_step_behind_forward(cuda):
    """utils/utils_fit_mul.py:78-92: eval-mode forward, then the criterion on its outputs -- forward's head maps go
    straight into dcfa_yolo_loss; the value must match the oracle criterion fed the same maps."""
    from nets.yolo_mul import YoloBody
    from nets.yolo_training import Loss
    with contextlib.redirect_stdout(io.StringIO()):
        net = YoloBody([128, 128], 2, 'n').eval().to(cuda)
    g = torch.Generator().manual_seed(3)
    rgb = torch.rand(3, 3, 128, 128, generator=g).to(cuda)
    nir = torch.rand(3, 3, 128, 128, generator=g).to(cuda)
    _, targets = OL.synth_case(seed=109, B=3, nc=2, hw0=(16, 16), n_targets=3)
    with torch.no_grad():
        outputs = net(rgb, nir)
        value = Loss(net)(outputs, torch.from_numpy(targets))
    ref = OL.loss_forward([t.cpu().numpy() for t in outputs[2]], targets, nc=2)
    assert abs(float(value) - float(ref[3])) <= RTOL * abs(float(ref[3])) + ATOL


def test_validate_one_epoch(cuda):
    """The reference's validation loop (utils/utils_fit_mul.py:78-94) over three batches, one of them without targets:
    the running mean equals the mean of the oracle criterion over the same forward outputs."""
    from nets.yolo_mul import YoloBody
    from nets.yolo_training import Loss
    from utils.utils_fit_mul import fit_one_epoch, validate_one_epoch
    with contextlib.redirect_stdout(io.StringIO()):
        net = YoloBody([96, 96], 1, 'n').eval().to(cuda)
    g = torch.Generator().manual_seed(4)
    batches = []
    for i in range(3):
        _, t = OL.synth_case(seed=120 + i, B=2, nc=1, hw0=(12, 12), n_targets=0 if i == 1 else 2)
        batches.append((torch.rand(2, 3, 96, 96, generator=g), torch.rand(2, 3, 96, 96, generator=g), torch.from_numpy(t)))
    got = validate_one_epoch(net, Loss(net), batches, 3, True, 0)
    want = 0.0
    with torch.no_grad():
        for rgb, nir, t in batches:
            maps = net(rgb.to(cuda), nir.to(cuda))[2]
            want += float(OL.loss_forward([m.cpu().numpy() for m in maps], t.numpy(), nc=1)[3])
    assert abs(got - want / 3) <= RTOL * abs(want / 3) + ATOL
    with pytest.raises(NotImplementedError):
        fit_one_epoch()


def test_loss_through_the_c_abi(cuda):
    """dcfa_yolo_loss called with raw pointers, as the header declares it (no Python wrapper in between)."""
    import ctypes as C
    from dcfa_b200 import _lib
    kw = dict(seed=130, B=2, nc=3, hw0=(24, 32), n_targets=4, tiny=1)
    feats, targets = OL.synth_case(**kw)
    ref = OL.loss_forward(feats, targets, nc=3)
    B, nc = 2, 3
    x = [torch.from_numpy(f).to(cuda) for f in feats]
    hw = [int(v) for f in feats for v in f.shape[2:]]
    A = sum(hw[2 * i] * hw[2 * i + 1] for i in range(3))
    gt = OL.preprocess(targets, B, np.array([hw[1] * 8, hw[0] * 8, hw[1] * 8, hw[0] * 8], np.float32))
    G = gt.shape[1]
    gt_dev = torch.from_numpy(gt).to(cuda)
    need = _lib.lib.dcfa_loss_workspace_bytes(B, A, nc, G)
    ws = torch.empty(need, dtype=torch.uint8, device=cuda)
    out = torch.zeros(8, device=cuda)
    st = torch.cuda.current_stream(cuda).cuda_stream
    args = (x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr(), B, nc, (C.c_int32 * 6)(*hw), (C.c_float * 3)(8., 16., 32.),
            gt_dev.data_ptr(), G, out.data_ptr(), ws.data_ptr())
    assert _lib.lib.dcfa_yolo_loss(*args, need, C.c_void_p(st)) == 0
    np.testing.assert_allclose(out.cpu().numpy()[:4], ref, rtol=RTOL, atol=ATOL)
    # a workspace that is too small is refused with a message, nothing is launched
    assert _lib.lib.dcfa_yolo_loss(*args, need - 256, C.c_void_p(st)) == -1
    assert b"workspace" in _lib.lib.dcfa_last_error()
