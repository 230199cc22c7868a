"""ctypes mirror of include/dcfa_b200.h (struct layouts and enum values only; no library load)."""
import ctypes as C

ABI_VERSION = 1

# op kinds
OP_STEM, OP_CONV, OP_DWCONV, OP_CBAM_POOL, OP_CBAM_MLP, OP_CBAM_STATS, OP_CBAM_APPLY, OP_MAXPOOL5, OP_UPSAMPLE, OP_DFL = range(1, 11)
STEM_FLAG_U8 = 0x100   # DCFA_STEM_FLAG_U8
STEM_FLAG_X2_PLANE = 0x200   # DCFA_STEM_FLAG_X2_PLANE
CONV_FLAG_CHAIN_HEAD = 0x400   # DCFA_CONV_FLAG_CHAIN_HEAD
CONV_FLAG_PAIR = 0x800         # DCFA_CONV_FLAG_PAIR
CONV_FLAG_GHOST_HEAD = 0x1000  # DCFA_CONV_FLAG_GHOST_HEAD
CONV_FLAG_DFL = 0x2000         # DCFA_CONV_FLAG_DFL
OP_NAMES = {OP_STEM: "stem", OP_CONV: "conv", OP_DWCONV: "dwconv", OP_CBAM_POOL: "cbam_pool", OP_CBAM_MLP: "cbam_mlp",
            OP_CBAM_STATS: "cbam_stats", OP_CBAM_APPLY: "cbam_apply", OP_MAXPOOL5: "maxpool5", OP_UPSAMPLE: "upsample",
            OP_DFL: "dfl"}
ACT_NONE, ACT_RELU, ACT_SILU = 0, 1, 2
OUT_BF16_NHWC, OUT_F32_NCHW = 0, 1
IOU_TV_CPU, IOU_TV_CUDA = 0, 1
E_INVALID, E_CUDA, E_ARCH = -1, -2, -3


class View(C.Structure):
    _fields_ = [("buf", C.c_int32), ("ld", C.c_int32), ("off", C.c_int64), ("img_stride", C.c_int64),
                ("gstride", C.c_int64), ("gi", C.c_int32), ("pad_", C.c_int32)]


def no_view():
    return View(-1, 0, 0, 0, 0, 0, 0)


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("act", C.c_int32), ("out_mode", C.c_int32), ("flags", C.c_int32),
                ("x", View), ("x2", View), ("y", View), ("w", View), ("scale", View), ("bias", View),
                ("a0", View), ("a1", View), ("a2", View),
                ("n_img", C.c_int32), ("group_imgs", C.c_int32),
                ("Hi", C.c_int32), ("Wi", C.c_int32), ("Cin", C.c_int32),
                ("Ho", C.c_int32), ("Wo", C.c_int32), ("Cout", C.c_int32),
                ("ksize", C.c_int32), ("stride", C.c_int32),
                ("BN", C.c_int32), ("n_tiles", C.c_int32), ("k_blocks", C.c_int32), ("K_real", C.c_int32),
                ("hidden", C.c_int32), ("parts", C.c_int32), ("nc", C.c_int32), ("A", C.c_int32),
                ("out_ctot", C.c_int32), ("out_coff", C.c_int32),
                ("w_gstride", C.c_int64), ("sb_gstride", C.c_int64),
                ("f0", C.c_float), ("f1", C.c_float), ("f2", C.c_float), ("f3", C.c_float)]


def new_op(kind, **kw):
    op = Op()
    op.kind = kind
    for name in ("x", "x2", "y", "w", "scale", "bias", "a0", "a1", "a2"):
        setattr(op, name, no_view())
    op.f0 = 1.0
    for k, v in kw.items():
        if not hasattr(op, k):
            raise AttributeError("dcfa_op has no field %r" % k)
        setattr(op, k, v)
    return op


class PlanInfo(C.Structure):
    """dcfa_plan_info"""
    _fields_ = [("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("num_classes", C.c_int32),
                ("anchors", C.c_int32), ("no", C.c_int32), ("level_hw", (C.c_int32 * 2) * 3), ("input_u8", C.c_int32),
                ("depth_plane", C.c_int32), ("reserved", C.c_int32 * 3)]


class PlanFileHeader(C.Structure):
    """dcfa_plan_file_header"""
    _fields_ = [("magic", C.c_char * 8), ("abi_version", C.c_int32), ("sizeof_op", C.c_int32), ("n_ops", C.c_int32),
                ("nbufs", C.c_int32), ("blob_bytes", C.c_int64), ("arena_bytes", C.c_int64), ("info", PlanInfo)]
