"""Batched prediction facade on the GPU: YOLO.detect_images must give, for every image of a batch
with mixed original sizes, exactly what the single-image path gives, and what the public pieces (letterbox ->
uint8 forward -> decode_box -> non_max_suppression with that image's shape) give when called by hand."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


def _facade(**kw):
    import importlib.util
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dcfa-yolo_b200")
    spec = importlib.util.spec_from_file_location("facade_yolo_mul", os.path.join(here, "yolo_mul.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(0)
        return mod.YOLO(model_path=None, **kw)


def _images(sizes, seed):
    rng = np.random.RandomState(seed)
    return [Image.fromarray(rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8), "RGB") for (w, h) in sizes]


def test_detect_images_batch_equals_single_and_manual(cuda, tmp_path):
    from utils.utils import cvtColor, resize_image
    yolo = _facade(phi="n", input_shape=[128, 128], confidence=0.5, nms_iou=0.3)
    sizes = [(150, 100), (90, 200), (128, 128), (300, 77)]
    rgb, nir = _images(sizes, 1), _images(sizes, 2)
    nir[1] = nir[1].convert("L")                      # a single-channel depth frame goes through cvtColor
    batch = yolo.detect_images(rgb, nir)
    assert len(batch) == 4 and any(r is not None for r in batch)
    for i in range(4):
        single = yolo.detect_images([rgb[i]], [nir[i]])[0]
        if batch[i] is None:
            assert single is None
            continue
        assert batch[i].dtype == np.float32 and batch[i].shape[1] == 6
        assert np.array_equal(batch[i], single), "image %d: batched result differs from the single-image call" % i
        # by hand, through the reference-compatible pieces
        a = np.asarray(resize_image(cvtColor(rgb[i]), (128, 128), True), dtype=np.uint8)[None]
        b = np.asarray(resize_image(cvtColor(nir[i]), (128, 128), True), dtype=np.uint8)[None]
        out = yolo.net(torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda))
        y = yolo.bbox_util.decode_box(out)
        manual = yolo.bbox_util.non_max_suppression(y, 1, [128, 128], np.array([sizes[i][1], sizes[i][0]]), True,
                                                    conf_thres=0.5, nms_thres=0.3)[0]
        assert np.array_equal(batch[i], manual)
        # boxes are in the image's own pixel frame
        assert batch[i][:, 4].min() >= 0.5
    # the reference-style single-pair API: annotated image out, detection file, timing
    with contextlib.redirect_stdout(io.StringIO()):
        drawn = yolo.detect_image(rgb[0].copy(), nir[0])
    assert drawn.size == rgb[0].size
    os.makedirs(tmp_path / "detection-results")
    yolo.get_map_txt("img0", rgb[0], nir[0], yolo.class_names, str(tmp_path))
    lines = open(tmp_path / "detection-results" / "img0.txt").read().strip().splitlines()
    assert len(lines) == len(batch[0]) and lines[0].startswith("cherry_tomato ")
    assert yolo.get_FPS(rgb[0], nir[0], 3) > 0.0
    with pytest.raises(NotImplementedError):
        yolo.detect_heatmap(rgb[0], nir[0], "x.png")
