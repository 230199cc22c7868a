"""Host-side helpers of the prediction facade (no GPU) against outputs of the REAL reference helpers recorded in
tests/golden/facade_helpers.npz by oracle/make_golden_facade.py: letterbox resize, colour conversion,
preprocess_input, and the un-letterbox of boxes -- including the batched per-image-shape form the facade uses."""
import os

import numpy as np
import pytest

from oracle.make_golden_facade import BOX_CASES, RESIZE_CASES, facade_inputs

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "facade_helpers.npz"))


def test_resize_cvtcolor_preprocess_match_reference():
    from utils.utils import cvtColor, preprocess_input, resize_image
    imgs, _ = facade_inputs()
    for i, ((seed, mode, _, size, lb), im) in enumerate(zip(RESIZE_CASES, imgs)):
        r = resize_image(cvtColor(im), size, lb)
        got = np.array(r, dtype=np.uint8)
        assert got.shape == GOLD["resize_%d" % i].shape
        assert np.array_equal(got, GOLD["resize_%d" % i]), "resize case %d (%s)" % (i, mode)
        pre = np.transpose(preprocess_input(np.array(r, dtype="float32")), (2, 0, 1))
        assert np.array_equal(pre, GOLD["pre_%d" % i])


def test_letterbox_batch_equals_per_image_resize():
    from utils.utils import letterbox_batch
    imgs, _ = facade_inputs()
    sel = [0, 1, 3, 5]
    batch, shapes = letterbox_batch([imgs[i] for i in sel], (64, 64), True)
    assert batch.dtype == np.uint8 and batch.shape == (4, 64, 64, 3)
    assert np.array_equal(batch[0], GOLD["resize_0"]) and np.array_equal(batch[2], GOLD["resize_3"])
    assert shapes.tolist() == [[23, 37], [120, 90], [64, 64], [33, 45]]


def test_yolo_correct_boxes_matches_reference_and_batched_form():
    from utils.utils_bbox import DecodeBox
    dec = DecodeBox(1, (640, 640))
    _, boxes = facade_inputs()
    for i, ((seed, n, ishape, imshape, lb), (xy, wh)) in enumerate(zip(BOX_CASES, boxes)):
        got = dec.yolo_correct_boxes(xy, wh, list(ishape), np.array(imshape), lb)
        assert np.array_equal(got, GOLD["boxes_%d" % i]), "box case %d" % i
        # one shape row per box (the batched facade): same numbers
        rows = np.repeat(np.array(imshape)[None, :], n, axis=0)
        assert np.array_equal(dec.yolo_correct_boxes(xy, wh, list(ishape), rows, lb), got)
    # two images of different shapes in one call == two separate calls
    (xy0, wh0), (xy1, wh1) = boxes[0], boxes[1]
    rows = np.concatenate([np.repeat(np.array([[480, 640]]), len(xy0), 0), np.repeat(np.array([[1080, 1920]]), len(xy1), 0)])
    both = dec.yolo_correct_boxes(np.concatenate([xy0, xy1]), np.concatenate([wh0, wh1]), [640, 640], rows, True)
    assert np.array_equal(both[:len(xy0)], GOLD["boxes_0"]) and np.array_equal(both[len(xy0):], GOLD["boxes_1"])


def test_get_classes_and_defaults():
    from utils.utils import get_classes
    import importlib.util
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dcfa-yolo_b200")
    names, n = get_classes(os.path.join(here, "model_data", "voc_classes.txt"))
    assert n == 1 and names == ["cherry_tomato"]
    spec = importlib.util.spec_from_file_location("facade_yolo_mul", os.path.join(here, "yolo_mul.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.YOLO.get_defaults("confidence") == 0.5 and mod.YOLO.get_defaults("nms_iou") == 0.3
    assert mod.YOLO.get_defaults("letterbox_image") is True and mod.YOLO.get_defaults("input_shape") == [640, 640]
    assert mod.YOLO.get_defaults("nope").startswith("Unrecognized")
    with pytest.raises(RuntimeError, match="no CPU path"):
        mod.YOLO(model_path=None, cuda=False)
