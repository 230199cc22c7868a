"""TEST INFRASTRUCTURE -- loads the REAL reference modules (heitieya/DCFA-YOLO) and, for configurations the shipped
code cannot run, the five-constant generalisation of SURVEY F1 as a SUBCLASS of the reference's own YoloBody.

The reference's packages are called `nets` and `utils`, like the drop-in's, so this module must run in a process
whose sys.path resolves them to the reference: oracle/make_golden.py (authoring container, /root/reference) and
oracle/ref_runner.py (a subprocess of bench.py; oracle/_ref/ staged by `make -C oracle ref`).  Nothing under
dcfa-yolo_b200/ imports it.
"""
import contextlib
import io
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    """$DCFA_REFERENCE, else /root/reference (authoring container), else the staged copy oracle/_ref (GPU box)."""
    for p in (os.environ.get("DCFA_REFERENCE"), "/root/reference", os.path.join(HERE, "_ref")):
        if p and os.path.exists(os.path.join(p, "nets", "yolo_mul.py")):
            return p
    return None


def load(ref=None):
    """Put the reference first on sys.path and import its hot-path modules.  Returns a namespace with YoloBody (as
    shipped), GeneralisedYoloBody, DecodeBox, make_anchors and the root used."""
    ref = ref or reference_root()
    if ref is None:
        raise ImportError("reference modules not found (neither /root/reference nor oracle/_ref)")
    for m in [k for k in sys.modules if k in ("nets", "utils") or k.startswith(("nets.", "utils."))]:
        if not getattr(sys.modules[m], "__file__", "").startswith(ref):
            raise ImportError("module %s is already imported from elsewhere: run the reference in its own process" % m)
    sys.path.insert(0, ref)
    import torch
    import torch.nn.functional as F
    from nets.repghost import C2f_repghost
    from nets.yolo_mul import YoloBody
    from utils.utils_bbox import DecodeBox, make_anchors

    sys.path.insert(1, os.path.dirname(HERE))
    from oracle import forward as O

    class GeneralisedYoloBody(YoloBody):
        """Reference YoloBody with the five hard-coded constants generalised (identity for phi='n' @ 640):
        nets/yolo_mul.py:361,:364,:376 (+128/+64/+256 input channels) and :426,:433 ((40,40)/(80,80) sizes)."""

        def __init__(self, input_shape, num_classes, phi):
            super().__init__(input_shape, num_classes, phi)
            bc, depth, c3, c4, c5 = O.dims(phi)
            self.conv3_for_upsample1 = C2f_repghost(c5 + 2 * c4, c4, depth, shortcut=False)
            self.conv3_for_upsample2 = C2f_repghost(c4 + 2 * c3, c3, depth, shortcut=False)
            self.conv3_for_downsample2 = C2f_repghost(c4 + 2 * c5, c5, depth, shortcut=False)

        def forward(self, rgb, nir):
            f1r, f2r, f3r = self.backbone_rgb.forward(rgb)
            f1n, f2n, f3n = self.backbone_nir.forward(nir)
            f1r, f1n = self.cbam_rgb_feat1(f1r), self.cbam_nir_feat1(f1n)
            f2r, f2n = self.cbam_rgb_feat2(f2r), self.cbam_nir_feat2(f2n)
            f3r, f3n = self.cbam_rgb_feat3(f3r), self.cbam_nir_feat3(f3n)
            feat3 = f3r + f3n
            p5_up = F.interpolate(feat3, size=f2r.shape[-2:], mode='bilinear', align_corners=True)
            p4 = self.conv3_for_upsample1(self.bi_fpn([p5_up, f2r, f2n]))
            p4_up = F.interpolate(p4, size=f1r.shape[-2:], mode='bilinear', align_corners=True)
            p3 = self.conv3_for_upsample2(self.bi_fpn([p4_up, f1r, f1n]))
            p4 = self.conv3_for_downsample1(torch.cat([self.down_sample1(p3), p4], 1))
            p5 = self.conv3_for_downsample2(self.bi_fpn([self.down_sample2(p4), f3r, f3n]))
            shape = p3.shape
            x = [p3, p4, p5]
            for i in range(self.nl):
                x[i] = torch.cat((self.cv2[i](x[i]), self.cv3[i](x[i])), 1)
            self.anchors, self.strides = (t.transpose(0, 1) for t in make_anchors(x, self.stride, 0.5))
            box, cls = torch.cat([xi.view(shape[0], self.no, -1) for xi in x], 2).split((self.reg_max * 4, self.num_classes), 1)
            return self.dfl(box), cls, x, self.anchors, self.strides

    class NS:
        pass

    ns = NS()
    ns.root, ns.YoloBody, ns.GeneralisedYoloBody, ns.DecodeBox, ns.make_anchors, ns.O = ref, YoloBody, GeneralisedYoloBody, DecodeBox, make_anchors, O
    return ns


def build(ns, phi, h, w, nc, shipped):
    """shipped=True: the unmodified reference class (only valid for phi='n' at 640x640)."""
    with contextlib.redirect_stdout(io.StringIO()):  # weights_init prints
        net = ns.YoloBody([h, w], nc, phi) if shipped else ns.GeneralisedYoloBody([h, w], nc, phi)
    return net.eval()
