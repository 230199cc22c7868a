"""Drop-in for the inference helpers of the reference's utils/utils.py (cvtColor :9-20, resize_image :24-37,
get_classes :41-45, preprocess_input :76-79, show_config :81-89), plus the batched letterbox the facade's
detect_images uses.  Training helpers (get_lr, seed_everything, worker_init_fn, download_weights) are out of scope.
"""
import numpy as np
from PIL import Image

LETTERBOX_FILL = (128, 128, 128)   # the reference pads with mid-grey (utils/utils.py:33)


def cvtColor(image):
    """Anything that is not already a 3-channel image becomes RGB (grey-scale depth/NIR frames, RGBA, palette)."""
    shape = np.shape(image)
    if len(shape) == 3 and shape[2] == 3:
        return image
    return image.convert('RGB')


def resize_image(image, size, letterbox_image):
    """PIL image -> PIL image of `size` = (w, h).  Letterbox: BICUBIC resize by the limiting side's scale (truncated
    to whole pixels), pasted centred on a grey canvas; otherwise a plain BICUBIC stretch."""
    w, h = size
    if not letterbox_image:
        return image.resize((w, h), Image.BICUBIC)
    iw, ih = image.size
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    canvas = Image.new('RGB', size, LETTERBOX_FILL)
    canvas.paste(image.resize((nw, nh), Image.BICUBIC), ((w - nw) // 2, (h - nh) // 2))
    return canvas


def letterbox_batch(images, input_shape, letterbox_image, out=None):
    """List of PIL images -> uint8 array [B, H, W, 3] (the stem kernel's uint8 NHWC input) and their original
    (h, w) shapes [B, 2].  `out` may be a preallocated (e.g. pinned) array to fill."""
    h, w = int(input_shape[0]), int(input_shape[1])
    if out is None:
        out = np.empty((len(images), h, w, 3), dtype=np.uint8)
    shapes = np.empty((len(images), 2), dtype=np.int64)
    for i, im in enumerate(images):
        shapes[i] = np.shape(im)[0:2]
        out[i] = np.asarray(resize_image(cvtColor(im), (w, h), letterbox_image), dtype=np.uint8)
    return out, shapes


def get_classes(classes_path):
    with open(classes_path, encoding='utf-8') as f:
        names = [line.strip() for line in f.readlines()]
    return names, len(names)


def preprocess_input(image):
    """In place, like the reference: float pixels -> [0, 1]."""
    image /= 255.0
    return image


def show_config(**kwargs):
    bar = '-' * 70
    print('Configurations:')
    print(bar)
    print('|%25s | %40s|' % ('keys', 'values'))
    print(bar)
    for key, value in kwargs.items():
        print('|%25s | %40s|' % (str(key), str(value)))
    print(bar)
