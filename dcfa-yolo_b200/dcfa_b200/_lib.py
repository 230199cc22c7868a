"""Loader for lib/libdcfa_b200.so (the C-ABI kernel library).  There is no fallback: if the library is
missing or its ABI does not match, importing this module raises."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libdcfa_b200.so")

# every symbol include/dcfa_b200.h declares
SYMBOLS = ("dcfa_abi_version", "dcfa_sizeof_view", "dcfa_sizeof_op", "dcfa_last_error", "dcfa_device_check",
           "dcfa_launch_count", "dcfa_run_ops", "dcfa_decode_box", "dcfa_nms_workspace_bytes", "dcfa_nms",
           "dcfa_letterbox_workspace_bytes", "dcfa_letterbox_u8", "dcfa_pack_detections",
           "dcfa_plan_create", "dcfa_plan_run", "dcfa_plan_num_launches", "dcfa_plan_destroy", "dcfa_plan_load",
           "dcfa_plan_get_info", "dcfa_plan_forward", "dcfa_loss_workspace_bytes", "dcfa_yolo_loss")


class DcfaError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("dcfa_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or `make -C dcfa-yolo_b200/csrc`); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for s in SYMBOLS:
        if not hasattr(lib, s):
            raise ImportError("dcfa_b200: %s does not export %s" % (LIB_PATH, s))
    lib.dcfa_last_error.restype = C.c_char_p
    lib.dcfa_launch_count.restype = C.c_int64
    lib.dcfa_nms_workspace_bytes.restype = C.c_int64
    lib.dcfa_nms_workspace_bytes.argtypes = [C.c_int, C.c_int]
    lib.dcfa_device_check.argtypes = [C.c_int]
    lib.dcfa_run_ops.argtypes = [C.POINTER(abi.Op), C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    lib.dcfa_decode_box.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    lib.dcfa_nms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double, C.c_int,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.dcfa_letterbox_workspace_bytes.restype = C.c_int64
    lib.dcfa_letterbox_workspace_bytes.argtypes = [C.c_int] * 5
    lib.dcfa_letterbox_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_int64, C.c_void_p]
    lib.dcfa_pack_detections.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p, C.c_void_p]
    lib.dcfa_plan_create.argtypes = [C.POINTER(abi.Op), C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
    lib.dcfa_plan_run.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    lib.dcfa_plan_num_launches.argtypes = [C.c_void_p]
    lib.dcfa_plan_destroy.argtypes = [C.c_void_p]
    lib.dcfa_plan_destroy.restype = None
    lib.dcfa_plan_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.dcfa_plan_get_info.argtypes = [C.c_void_p, C.POINTER(abi.PlanInfo)]
    lib.dcfa_plan_forward.argtypes = [C.c_void_p] * 9
    lib.dcfa_loss_workspace_bytes.restype = C.c_int64
    lib.dcfa_loss_workspace_bytes.argtypes = [C.c_int] * 4
    lib.dcfa_yolo_loss.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32),
                                   C.POINTER(C.c_float), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    if lib.dcfa_abi_version() != abi.ABI_VERSION:
        raise ImportError("dcfa_b200: ABI version %d != %d" % (lib.dcfa_abi_version(), abi.ABI_VERSION))
    if lib.dcfa_sizeof_view() != C.sizeof(abi.View) or lib.dcfa_sizeof_op() != C.sizeof(abi.Op):
        raise ImportError("dcfa_b200: struct layout mismatch (view %d/%d, op %d/%d)" % (
            lib.dcfa_sizeof_view(), C.sizeof(abi.View), lib.dcfa_sizeof_op(), C.sizeof(abi.Op)))
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise DcfaError("dcfa_b200 error %d: %s" % (rc, lib.dcfa_last_error().decode("utf-8", "replace")))


def run_ops(ops, bufs, stream):
    """ops: ctypes array of abi.Op (or list); bufs: list of device pointers (ints, 0 for unused); stream: int."""
    if isinstance(ops, (list, tuple)):
        ops = (abi.Op * len(ops))(*ops)
    arr = (C.c_void_p * len(bufs))(*[C.c_void_p(int(b)) if b else None for b in bufs])
    check(lib.dcfa_run_ops(ops, len(ops), arr, len(bufs), C.c_void_p(int(stream))))


def launch_count():
    return int(lib.dcfa_launch_count())
