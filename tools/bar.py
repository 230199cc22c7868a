#!/usr/bin/env python
"""Calibration on the GPU box: (a) device copy bandwidth as this process sees it, (b) the "library bar":
the oracle's torch restatement of the reference forward run eagerly on the GPU (cuDNN), fp32 and bf16
autocast, channels_last, same batch as bench.py.  Not part of the product path."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dcfa-yolo_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import forward as O  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    n = 512 * 1024 * 1024
    a = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
    b = torch.empty_like(a)
    ms = timeit(lambda: b.copy_(a))
    print("copy 1 GiB bf16: %.3f ms -> %.0f GB/s (read+write)" % (ms, 2 * n * 2 / ms / 1e6))
    for mb in (64, 256):
        m = mb * 1024 * 1024 // 2
        ms = timeit(lambda: b[:m].copy_(a[:m]), iters=50)
        print("copy %d MiB: %.4f ms -> %.0f GB/s" % (mb, ms, 2 * m * 2 / ms / 1e6))
    del a, b
    B, S = 32, 640
    net = bench.build_model("s", S)
    sd = {k: v.to(dev) for k, v in net.state_dict().items() if v.dtype.is_floating_point}
    rgb = torch.rand(B, 3, S, S, device=dev).contiguous(memory_format=torch.channels_last)
    nir = torch.rand(B, 3, S, S, device=dev).contiguous(memory_format=torch.channels_last)

    def fwd():
        return O_forward_gpu(sd, rgb, nir)

    import torch.nn.functional as F  # noqa: F401

    def O_forward_gpu(sd, rgb, nir):
        # oracle.forward.yolo_forward moves the state_dict to CPU; re-implement the entry with device tensors
        f = O
        _, depth, _, _, _ = f.dims("s")
        f1r, f2r, f3r = f.backbone(rgb, sd, 'backbone_rgb')
        f1n, f2n, f3n = f.backbone(nir, sd, 'backbone_nir')
        f1r, f1n = f.cbam(f1r, sd, 'cbam_rgb_feat1'), f.cbam(f1n, sd, 'cbam_nir_feat1')
        f2r, f2n = f.cbam(f2r, sd, 'cbam_rgb_feat2'), f.cbam(f2n, sd, 'cbam_nir_feat2')
        f3r, f3n = f.cbam(f3r, sd, 'cbam_rgb_feat3'), f.cbam(f3n, sd, 'cbam_nir_feat3')
        feat3 = f3r + f3n
        p5_up = F.interpolate(feat3, size=f2r.shape[-2:], mode='bilinear', align_corners=True)
        p4 = f.c2f_repghost(f.bifpn_concat([p5_up, f2r, f2n], sd), sd, 'conv3_for_upsample1', depth)
        p4_up = F.interpolate(p4, size=f1r.shape[-2:], mode='bilinear', align_corners=True)
        p3 = f.c2f_repghost(f.bifpn_concat([p4_up, f1r, f1n], sd), sd, 'conv3_for_upsample2', depth)
        p4 = f.c2f_repghost(torch.cat([f.conv_bn_silu(p3, sd, 'down_sample1', 2), p4], 1), sd, 'conv3_for_downsample1', depth)
        p5 = f.c2f_repghost(f.bifpn_concat([f.conv_bn_silu(p4, sd, 'down_sample2', 2), f3r, f3n], sd), sd, 'conv3_for_downsample2', depth)
        return [torch.cat((f.head_branch(p, sd, 'cv2.%d' % i), f.head_branch(p, sd, 'cv3.%d' % i)), 1) for i, p in enumerate((p3, p4, p5))]

    torch.backends.cudnn.benchmark = True
    with torch.no_grad():
        ms32 = timeit(fwd, iters=5, warm=3)
        print("eager cuDNN fp32 channels_last forward, B=%d: %.2f ms -> %.0f pairs/s" % (B, ms32, B / ms32 * 1e3))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            msbf = timeit(fwd, iters=5, warm=3)
        print("eager cuDNN bf16 autocast forward,       B=%d: %.2f ms -> %.0f pairs/s" % (B, msbf, B / msbf * 1e3))


if __name__ == "__main__":
    main()
