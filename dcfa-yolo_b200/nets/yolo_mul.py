"""Drop-in for the reference's nets/yolo_mul.py.

`YoloBody(input_shape, num_classes, phi, pretrained=False)` keeps the reference's constructor, attributes,
state_dict key set and `forward(rgb, nir) -> (dbox, cls, x, anchors, strides)` contract
(reference nets/yolo_mul.py:328-462), but the forward pass is executed by hand-written sm_100a kernels
(lib/libdcfa_b200.so) through a compiled plan.  The sub-modules below only hold parameters under the
reference's names; there is no eager PyTorch path and no CPU fallback.

Generalisation: the reference hard-codes +128/+64/+256 channels (:361,:364,:376) and (40,40)/(80,80)
interpolate sizes (:426,:433), valid only for phi='n' at 640x640.  Here they are 8*bc, 4*bc, C5 and the lateral
feature's size; for phi='n' @ 640 this is the identity, and it makes phi in {s,m,l,x} and other sizes runnable.
"""
import torch
import torch.nn as nn

from nets.repghost import C2f_repghost, _Holder
from nets.yolo_training import weights_init

_DEPTH = {'n': 0.33, 's': 0.33, 'm': 0.67, 'l': 1.00, 'x': 1.00}
_WIDTH = {'n': 0.25, 's': 0.50, 'm': 0.75, 'l': 1.00, 'x': 1.25}
_DEEP = {'n': 1.00, 's': 1.00, 'm': 0.75, 'l': 0.50, 'x': 0.50}


class Conv(_Holder):
    """conv(bias=False, pad=k//2) + BatchNorm2d(eps=1e-3, momentum=0.03) + SiLU  (reference :190-204)."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2 if p is None else p, groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=0.001, momentum=0.03, affine=True, track_running_stats=True)


class ChannelAttention(_Holder):
    def __init__(self, in_planes, ratio=8):
        super().__init__()
        self.fc1 = nn.Conv2d(in_planes, in_planes // ratio, 1, bias=False)
        self.fc2 = nn.Conv2d(in_planes // ratio, in_planes, 1, bias=False)


class SpatialAttention(_Holder):
    def __init__(self, kernel_size=7):
        super().__init__()
        if kernel_size != 7:
            raise ValueError("the DCFA-YOLO path only uses 7x7 spatial attention")
        self.conv1 = nn.Conv2d(2, 1, kernel_size, padding=3, bias=False)


class CBAM(_Holder):
    def __init__(self, channel, ratio=8, kernel_size=7):
        super().__init__()
        self.channelattention = ChannelAttention(channel, ratio=ratio)
        self.spatialattention = SpatialAttention(kernel_size=kernel_size)


class Concat_BiFPN(_Holder):
    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension
        self.w = nn.Parameter(torch.ones(3, dtype=torch.float32), requires_grad=True)
        self.epsilon = 0.0001


class SPPF_CBAM(_Holder):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        for i in range(1, 5):
            setattr(self, 'cbam%d' % i, CBAM(c_, c_))   # ratio = c_  ->  hidden width 1, as in the reference (:18-21)


class Conv_maxpool(_Holder):
    def __init__(self, c1, c2):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(c1, c2, 3, 1, 1, bias=False), nn.BatchNorm2d(c2), nn.ReLU(inplace=True))


class ShuffleNetV2(_Holder):
    def __init__(self, inp, oup, stride):
        super().__init__()
        if stride != 1 or inp != oup:
            raise ValueError("the DCFA-YOLO backbone only uses stride-1 ShuffleNetV2 units")
        bf = oup // 2
        self.branch1 = nn.Sequential()
        self.branch2 = nn.Sequential(
            nn.Conv2d(bf, bf, 1, 1, 0, bias=False), nn.BatchNorm2d(bf), nn.ReLU(inplace=True),
            nn.Conv2d(bf, bf, 3, 1, 1, groups=bf), nn.BatchNorm2d(bf),
            nn.Conv2d(bf, bf, 1, 1, 0, bias=False), nn.BatchNorm2d(bf), nn.ReLU(inplace=True))


class Backbone(_Holder):
    def __init__(self, base_channels, base_depth, deep_mul, phi, pretrained=False):
        super().__init__()
        if pretrained:
            raise NotImplementedError("pretrained backbone download (reference :283-293) needs network access")
        bc, c5 = base_channels, int(base_channels * 16 * deep_mul)
        self.stem = Conv_maxpool(3, bc)
        self.dark2 = nn.Sequential(Conv(bc, bc * 2, 3, 2), ShuffleNetV2(bc * 2, bc * 2, 1))
        self.dark3 = nn.Sequential(Conv(bc * 2, bc * 4, 3, 2), ShuffleNetV2(bc * 4, bc * 4, 1))
        self.dark4 = nn.Sequential(Conv(bc * 4, bc * 8, 3, 2), ShuffleNetV2(bc * 8, bc * 8, 1))
        self.dark5 = nn.Sequential(Conv(bc * 8, c5, 3, 2), ShuffleNetV2(c5, c5, 1), SPPF_CBAM(c5, c5, k=5))


class DFL(_Holder):
    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1


class YoloBody(nn.Module):
    def __init__(self, input_shape, num_classes, phi, pretrained=False):
        super().__init__()
        dep_mul, wid_mul, deep_mul = _DEPTH[phi], _WIDTH[phi], _DEEP[phi]
        bc = int(wid_mul * 64)
        base_depth = max(round(dep_mul * 3), 1)
        c3, c4, c5 = bc * 4, bc * 8, int(bc * 16 * deep_mul)
        self.phi = phi

        self.backbone_rgb = Backbone(bc, base_depth, deep_mul, phi, pretrained=pretrained)
        self.backbone_nir = Backbone(bc, base_depth, deep_mul, phi, pretrained=pretrained)
        self.bi_fpn = Concat_BiFPN(dimension=1)
        self.cbam_rgb_feat1, self.cbam_nir_feat1 = CBAM(c3), CBAM(c3)
        self.cbam_rgb_feat2, self.cbam_nir_feat2 = CBAM(c4), CBAM(c4)
        self.cbam_rgb_feat3, self.cbam_nir_feat3 = CBAM(c5, ratio=8, kernel_size=7), CBAM(c5, ratio=8, kernel_size=7)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")   # unused by forward, kept for attribute parity (:358)

        self.conv3_for_upsample1 = C2f_repghost(c5 + 2 * c4, c4, base_depth, shortcut=False)
        self.conv3_for_upsample2 = C2f_repghost(c4 + 2 * c3, c3, base_depth, shortcut=False)
        self.down_sample1 = Conv(c3, c3, 3, 2)
        self.conv3_for_downsample1 = C2f_repghost(c4 + c3, c4, base_depth, shortcut=False)
        self.down_sample2 = Conv(c4, c4, 3, 2)
        self.conv3_for_downsample2 = C2f_repghost(c4 + 2 * c5, c5, base_depth, shortcut=False)

        ch = [c3, c4, c5]
        self.shape = None
        self.nl = len(ch)
        self.stride = torch.tensor([8.0, 16.0, 32.0])   # the reference derives these from a dummy forward (:382)
        self.reg_max = 16
        self.no = num_classes + self.reg_max * 4
        self.num_classes = num_classes
        c2h, c3h = max(16, ch[0] // 4, self.reg_max * 4), max(ch[0], num_classes)
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2h, 3), Conv(c2h, c2h, 3), nn.Conv2d(c2h, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(Conv(x, c3h, 3), Conv(c3h, c3h, 3), nn.Conv2d(c3h, num_classes, 1)) for x in ch)
        if not pretrained:
            weights_init(self)
        self.dfl = DFL(self.reg_max)

        self._engines = {}
        self._weights_dirty = True

    # ------------------------------------------------------------------ plan / weight-pack invalidation
    def invalidate_plan(self):
        """Call after modifying parameters or buffers in place; load_state_dict / .to() / .train() do it for you."""
        self._engines = {}
        self._weights_dirty = True

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_plan()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, '_engines'):
            self.invalidate_plan()
        return out

    def train(self, mode=True):
        out = super().train(mode)
        if hasattr(self, '_engines'):
            self.invalidate_plan()
        return out

    def _engine(self, batch, height, width, device, input_u8=False, depth_plane=False):
        from dcfa_b200.engine import Engine
        key = (batch, height, width, str(device), bool(input_u8), bool(depth_plane))
        eng = self._engines.get(key)
        if eng is None:
            eng = Engine(self.state_dict(), self.phi, self.num_classes, batch, height, width, device, input_u8=input_u8,
                         depth_plane=depth_plane)
            self._engines[key] = eng
        return eng

    def forward(self, rgb, nir):
        if self.training:
            raise NotImplementedError("dcfa_b200 accelerates the inference path only: call .eval() first "
                                      "(training-mode BatchNorm / backward are out of scope)")
        if not (torch.is_tensor(rgb) and torch.is_tensor(nir) and rgb.is_cuda and nir.is_cuda):
            raise RuntimeError("dcfa_b200 has no CPU path: YoloBody.forward needs CUDA tensors on an sm_100a device")
        if rgb.dtype == torch.uint8 or nir.dtype == torch.uint8:
            # Extension of the reference signature: raw uint8 images [B,H,W,3] (what cvtColor/resize_image produce,
            # utils/utils.py:9-37).  preprocess_input's /255 and the HWC->CHW transpose (yolo_mul.py:76) happen
            # inside the stem kernel; the upload is 4x smaller than the fp32 tensor.  The depth image may also be
            # given as the single plane [B,H,W] (or [B,H,W,1]) that cvtColor replicates to three channels
            # (utils/utils.py:14-19): the stem replicates it, the result is identical, the upload a third.
            plane = nir.dim() == 3 or (nir.dim() == 4 and nir.shape[3] == 1)
            ok = (rgb.dtype == nir.dtype and rgb.dim() == 4 and rgb.shape[3] == 3 and
                  (tuple(nir.shape[:3]) == tuple(rgb.shape[:3])) and (plane or nir.shape == rgb.shape))
            if not ok:
                raise ValueError("expected uint8 (B,H,W,3) images (depth optionally (B,H,W)), got %s %s and %s %s" % (
                    rgb.dtype, tuple(rgb.shape), nir.dtype, tuple(nir.shape)))
            b, h, w, _ = rgb.shape
            if plane:
                nir = nir.reshape(b, h, w)
            eng = self._engine(b, h, w, rgb.device, input_u8=True, depth_plane=plane)
            with torch.no_grad():
                dbox, cls, x = eng.run(rgb.detach().contiguous(), nir.detach().contiguous())
            self.anchors, self.strides, self.shape = eng.anchors, eng.strides, (b, x[0].shape[1], x[0].shape[2], x[0].shape[3])
            return dbox, cls, x, eng.anchors, eng.strides
        if rgb.shape != nir.shape or rgb.dim() != 4 or rgb.shape[1] != 3:
            raise ValueError("expected two (B,3,H,W) tensors, got %s and %s" % (tuple(rgb.shape), tuple(nir.shape)))
        b, _, h, w = rgb.shape
        eng = self._engine(b, h, w, rgb.device)
        with torch.no_grad():
            rgb = rgb.detach().float().contiguous()
            nir = nir.detach().float().contiguous()
            dbox, cls, x = eng.run(rgb, nir)
        self.anchors, self.strides, self.shape = eng.anchors, eng.strides, (b, x[0].shape[1], x[0].shape[2], x[0].shape[3])
        return dbox, cls, x, eng.anchors, eng.strides
