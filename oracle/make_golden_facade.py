"""TEST INFRASTRUCTURE -- golden vectors for the host-side helpers of the prediction facade.

Run in the authoring container only (needs /root/reference):  python oracle/make_golden_facade.py
Imports the REAL reference helpers (utils/utils.py: cvtColor, resize_image, preprocess_input; utils/utils_bbox.py:
DecodeBox.yolo_correct_boxes) and records their outputs on seeded inputs into tests/golden/facade_helpers.npz.
The inputs are regenerated from the seeds by tests/test_facade_cpu.py (facade_inputs below), not stored.
"""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (seed, mode, (w, h) of the source image, (w, h) target, letterbox)
RESIZE_CASES = [
    (1, 'RGB', (37, 23), (64, 64), True),
    (2, 'RGB', (90, 120), (64, 48), True),
    (3, 'L', (50, 50), (32, 64), True),
    (4, 'RGB', (64, 64), (64, 64), True),
    (5, 'RGB', (31, 77), (48, 48), False),
    (6, 'RGBA', (45, 33), (64, 32), True),
]
# (seed, n boxes, input_shape (h, w), image_shape (h, w), letterbox)
BOX_CASES = [
    (11, 7, (640, 640), (480, 640), True),
    (12, 5, (640, 640), (1080, 1920), True),
    (13, 4, (128, 96), (77, 300), True),
    (14, 6, (640, 640), (333, 500), False),
]


def facade_inputs():
    imgs = []
    for seed, mode, (w, h), _, _ in RESIZE_CASES:
        rng = np.random.RandomState(seed)
        ch = {'RGB': 3, 'L': 1, 'RGBA': 4}[mode]
        a = rng.randint(0, 256, size=(h, w, ch) if ch > 1 else (h, w)).astype(np.uint8)
        imgs.append(Image.fromarray(a, mode))
    boxes = []
    for seed, n, _, _, _ in BOX_CASES:
        rng = np.random.RandomState(seed)
        xy = rng.rand(n, 2).astype(np.float32)
        wh = (rng.rand(n, 2) * 0.4).astype(np.float32)
        boxes.append((xy, wh))
    return imgs, boxes


def main():
    sys.path.insert(0, '/root/reference')
    from utils.utils import cvtColor, preprocess_input, resize_image          # the reference's own helpers
    from utils.utils_bbox import DecodeBox
    imgs, boxes = facade_inputs()
    out = {}
    for i, ((seed, mode, _, size, lb), im) in enumerate(zip(RESIZE_CASES, imgs)):
        r = resize_image(cvtColor(im), size, lb)
        out['resize_%d' % i] = np.array(r, dtype=np.uint8)
        out['pre_%d' % i] = np.transpose(preprocess_input(np.array(r, dtype='float32')), (2, 0, 1))
    dec = DecodeBox(1, (640, 640))
    for i, ((seed, n, ishape, imshape, lb), (xy, wh)) in enumerate(zip(BOX_CASES, boxes)):
        out['boxes_%d' % i] = dec.yolo_correct_boxes(xy, wh, list(ishape), np.array(imshape), lb)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'facade_helpers.npz'), **out)
    print('wrote', len(out), 'arrays')


if __name__ == '__main__':
    main()
