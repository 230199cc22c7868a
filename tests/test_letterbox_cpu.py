"""The oracle's restatement of Pillow's 8-bit BICUBIC resample + the reference's letterbox (utils/utils.py:24-37) against the
installed Pillow itself -- bit-exact (integer work).  The GPU kernel is compared with the same oracle in test_letterbox_gpu.py."""
import numpy as np
import pytest
from PIL import Image

from oracle import letterbox as L

SIZES = [(640, 480), (480, 640), (1920, 1080), (333, 517), (64, 64), (640, 640), (1280, 720), (97, 1003), (700, 641), (2, 3),
         (641, 640), (640, 641), (31, 640), (1000, 640)]


def ref_resize_image(arr, size, letterbox_image):
    """the reference function, restated with PIL calls only (utils/utils.py:24-37)"""
    image = Image.fromarray(arr if arr.shape[2] == 3 else arr[..., 0])
    iw, ih = image.size
    w, h = size
    if letterbox_image:
        scale = min(w / iw, h / ih)
        nw, nh = int(iw * scale), int(ih * scale)
        image = image.resize((nw, nh), Image.BICUBIC)
        new_image = Image.new(image.mode, size, (128, 128, 128) if image.mode == 'RGB' else 128)
        new_image.paste(image, ((w - nw) // 2, (h - nh) // 2))
    else:
        new_image = image.resize((w, h), Image.BICUBIC)
    out = np.asarray(new_image)
    return out if out.ndim == 3 else out[..., None]


@pytest.mark.parametrize("iw,ih", SIZES)
@pytest.mark.parametrize("letterbox", [True, False])
def test_oracle_letterbox_matches_pillow(iw, ih, letterbox):
    rng = np.random.RandomState(iw * 7 + ih)
    for c, target in ((3, (640, 640)), (1, (640, 640)), (3, (96, 160))):
        arr = rng.randint(0, 256, (ih, iw, c), dtype=np.uint8)
        if c == 3 and iw * ih < 200000:      # smooth content too (weights of both signs matter differently)
            arr = (np.add.outer(np.arange(ih), np.arange(iw))[..., None] * np.array([1, 2, 3]) % 256).astype(np.uint8)
        got = L.resize_image(arr, target, letterbox)
        want = ref_resize_image(arr, target, letterbox)
        assert got.shape == want.shape
        assert np.array_equal(got, want), "max diff %d" % np.abs(got.astype(int) - want.astype(int)).max()
