"""Drop-in for the reference's nets/repghost.py: the RepGhost neck blocks as PARAMETER CONTAINERS.

Module / parameter names and shapes are identical to the reference (so its checkpoints load unchanged):
  RepGhostModule      nets/repghost.py:70-115   primary_conv.{0,1}, fusion_bn.0, cheap_operation.{0,1}
  RepGhostBottleneck  nets/repghost.py:178-279  ghost1, ghost2 (stride 1, no SE, identity shortcut)
  C2f_repghost        nets/repghost.py:308-320  cv1, cv2, m[i]
  Conv                nets/repghost.py:291-305  conv, bn (eps 1e-5)
None of them computes anything: the arithmetic runs in lib/libdcfa_b200.so, driven by YoloBody.forward, which
reads these parameters through the plan compiler (dcfa_b200/plan.py).  There is no eager / CPU path.
"""
import torch.nn as nn


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("%s is a parameter container of the dcfa_b200 drop-in; only YoloBody.forward(rgb, nir) "
                           "executes (on a CUDA sm_100a device)" % type(self).__name__)


class Conv(_Holder):
    """conv(bias=False, pad=k//2) + BatchNorm2d(eps=1e-5) + SiLU."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2 if p is None else p, groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2)


class RepGhostModule(_Holder):
    """1x1 conv + BN (+SiLU), then depthwise 3x3 + BN plus a BN-only identity branch (+SiLU when relu=True)."""

    def __init__(self, inp, oup, relu=True):
        super().__init__()
        self.has_act = bool(relu)
        self.primary_conv = nn.Sequential(nn.Conv2d(inp, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup),
                                          nn.SiLU(inplace=True) if relu else nn.Sequential())
        self.fusion_conv = nn.Sequential(nn.Identity())
        self.fusion_bn = nn.Sequential(nn.BatchNorm2d(oup))
        self.cheap_operation = nn.Sequential(nn.Conv2d(oup, oup, 3, 1, 1, groups=oup, bias=False), nn.BatchNorm2d(oup))


class RepGhostBottleneck(_Holder):
    def __init__(self, in_chs, mid_chs, out_chs, dw_kernel_size=3):
        super().__init__()
        if not (in_chs == mid_chs == out_chs):
            raise ValueError("the DCFA-YOLO neck only uses RepGhostBottleneck(c, c, c)")
        self.ghost1 = RepGhostModule(in_chs, mid_chs, relu=True)
        self.ghost2 = RepGhostModule(mid_chs, out_chs, relu=False)
        self.shortcut = nn.Sequential()


class C2f_repghost(_Holder):
    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(RepGhostBottleneck(self.c, self.c, self.c) for _ in range(n))
