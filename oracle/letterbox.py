"""TEST INFRASTRUCTURE -- numpy restatement of the reference's resize_image (utils/utils.py:24-37): Pillow's
`image.resize((nw, nh), Image.BICUBIC)` for 8-bit images pasted on a grey (128) canvas.

Pillow is an un-vendored dependency of the reference (requirements.txt, unpinned; 11.x here); its algorithm for 8-bit images
(libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc,
ImagingResampleInner) is restated below and pinned bit-exactly against the installed Pillow by tests/test_letterbox_cpu.py
over many sizes, both letterbox modes, RGB and single-plane images.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x):
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1, np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))


def precompute_coeffs(in_size, out_size):
    """-> (ksize, bounds [out,2] int, kk [out,ksize] int32)  (Resample.c: precompute_coeffs + normalize_coeffs_8bpc, box = whole axis)"""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = bicubic_filter((np.arange(xmax, dtype=np.float64) + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                       # sequential accumulation, as the C loop does
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        f = w * float(1 << PRECISION_BITS)
        kk[xx, :xmax] = np.where(w < 0, np.trunc(-0.5 + f), np.trunc(0.5 + f)).astype(np.int32)
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_axis(img, out_size, bounds, kk, axis):
    """one fixed-point pass along `axis` (0: vertical, 1: horizontal) of a uint8 [H,W,C] image"""
    src = img.astype(np.int64)
    shape = list(img.shape)
    shape[axis] = out_size
    out = np.zeros(shape, np.uint8)
    for o in range(out_size):
        lo, n = bounds[o]
        k = kk[o, :n].astype(np.int64)
        if axis == 1:
            acc = (src[:, lo:lo + n, :] * k[None, :, None]).sum(1) + (1 << (PRECISION_BITS - 1))
            out[:, o, :] = _clip8(acc)
        else:
            acc = (src[lo:lo + n, :, :] * k[:, None, None]).sum(0) + (1 << (PRECISION_BITS - 1))
            out[o, :, :] = _clip8(acc)
    return out


def pil_resize_bicubic(img, nw, nh):
    """uint8 [H,W,C] -> [nh,nw,C]; ImagingResampleInner: horizontal pass first (over the rows the vertical pass reads)"""
    h, w = img.shape[:2]
    need_h, need_v = nw != w, nh != h
    if need_v:
        _, bv, kv = precompute_coeffs(h, nh)
        first, last = int(bv[0, 0]), int(bv[-1, 0] + bv[-1, 1])
    if need_h:
        _, bh, kh = precompute_coeffs(w, nw)
        if need_v:
            img = resample_axis(img[first:last], nw, bh, kh, 1)
            bv = bv.copy()
            bv[:, 0] -= first
        else:
            img = resample_axis(img, nw, bh, kh, 1)
    if need_v:
        img = resample_axis(img, nh, bv, kv, 0)
    return img.copy()


def resize_image(img, size, letterbox_image):
    """utils/utils.py:24-37 on a uint8 [H,W,C] array; size = (w, h) as in the reference.  -> uint8 [h,w,C]"""
    ih, iw = img.shape[:2]
    w, h = size
    if not letterbox_image:
        return pil_resize_bicubic(img, w, h)
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    out = np.full((h, w, img.shape[2]), 128, np.uint8)
    out[(h - nh) // 2:(h - nh) // 2 + nh, (w - nw) // 2:(w - nw) // 2 + nw] = pil_resize_bicubic(img, nw, nh)
    return out
