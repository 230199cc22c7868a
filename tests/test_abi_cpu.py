"""The C-ABI library loads without a GPU and exports every symbol include/dcfa_b200.h declares; the ctypes mirror
matches the C struct layout; the product path refuses to run without CUDA (no CPU fallback)."""
import contextlib
import ctypes as C
import io
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dcfa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcfa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from dcfa_b200 import _lib, abi
    syms = declared_symbols()
    assert set(syms) == set(_lib.SYMBOLS), (syms, _lib.SYMBOLS)
    for s in syms:
        assert hasattr(_lib.lib, s), s
    assert _lib.lib.dcfa_abi_version() == abi.ABI_VERSION
    assert _lib.lib.dcfa_sizeof_view() == C.sizeof(abi.View) == 40
    assert _lib.lib.dcfa_sizeof_op() == C.sizeof(abi.Op)
    assert _lib.lib.dcfa_nms_workspace_bytes(2, 8400) > 2 * 8400 * 16
    assert _lib.lib.dcfa_nms_workspace_bytes(0, 8400) == 0


def test_error_reporting_without_a_gpu():
    """Bad arguments come back as codes + messages, never as exceptions / aborts across the ABI."""
    from dcfa_b200 import _lib
    rc = _lib.lib.dcfa_run_ops(None, 1, None, 0, None)
    assert rc == -1 and b"run_ops" in _lib.lib.dcfa_last_error()
    rc = _lib.lib.dcfa_decode_box(None, None, 0, None, 0, 0, None, 1, 1, 1, 1.0, 1.0, None, None)
    assert rc == -1 and b"decode_box" in _lib.lib.dcfa_last_error()
    rc = _lib.lib.dcfa_nms(None, 1, 1, 1, 0.5, 0.5, 0, None, None, None, None, None, 0, None)
    assert rc == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a machine without CUDA")
def test_no_cpu_fallback():
    import numpy as np
    from nets.yolo_mul import YoloBody
    from utils.utils_bbox import DecodeBox
    with contextlib.redirect_stdout(io.StringIO()):
        net = YoloBody([64, 64], 1, 'n').eval()
    x = torch.rand(1, 3, 64, 64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net(x, x)
    dec = DecodeBox(1, (64, 64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        dec.decode_box((torch.rand(1, 4, 84), torch.rand(1, 1, 84), None, torch.rand(2, 84), torch.rand(1, 84)))
    with pytest.raises(RuntimeError, match="no CPU path"):
        dec.non_max_suppression(torch.rand(1, 84, 5), 1, [64, 64], np.array([64, 64]), True)


def test_state_dict_keys_match_the_reference():
    """Parameter / buffer names and shapes are the drop-in contract (reference checkpoints must load unchanged)."""
    import json
    from nets.yolo_mul import YoloBody
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    for tag, ref in keys.items():
        phi, nc = tag.split("_nc")
        with contextlib.redirect_stdout(io.StringIO()):
            net = YoloBody([640, 640], int(nc), phi)
        got = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert got == ref, tag
        assert net.stride.tolist() == [8.0, 16.0, 32.0] and net.no == 64 + int(nc) and net.reg_max == 16
        assert torch.equal(net.dfl.conv.weight.reshape(-1), torch.arange(16.0)) and not net.dfl.conv.weight.requires_grad
