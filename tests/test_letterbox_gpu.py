"""Device-side pre/post-processing on a B200, bit-exact (integer / IEEE-double work):
  * dcfa_letterbox_u8 vs the oracle's restatement of the reference's resize_image (PIL BICUBIC + grey letterbox,
    utils/utils.py:24-37), which tests/test_letterbox_cpu.py pins against Pillow itself;
  * dcfa_pack_detections' un-letterbox vs DecodeBox.yolo_correct_boxes (numpy, pinned to the reference's function by
    tests/golden/facade_helpers.npz) in the reference's call pattern (utils/utils_bbox.py:170-173)."""
import numpy as np
import pytest
import torch

from test_letterbox_cpu import SIZES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("iw,ih", SIZES)
@pytest.mark.parametrize("letterbox", [True, False])
def test_device_letterbox_is_bit_exact(cuda, iw, ih, letterbox):
    from dcfa_b200.engine import letterbox_u8
    from oracle import letterbox as L
    rng = np.random.RandomState(iw * 7 + ih)
    ws = None
    for c, (tw, th) in ((3, (640, 640)), (1, (640, 640)), (3, (96, 160))):
        arr = rng.randint(0, 256, (ih, iw, c), dtype=np.uint8)
        want = L.resize_image(arr, (tw, th), letterbox)
        src = torch.from_numpy(arr if c == 3 else arr[..., 0].copy()).to(cuda)
        dst = torch.zeros((th, tw, 3) if c == 3 else (th, tw), dtype=torch.uint8, device=cuda)
        ws = letterbox_u8(src, dst, letterbox, ws)
        got = dst.cpu().numpy().reshape(want.shape)
        assert np.array_equal(got, want), "c=%d: %d pixels differ, max %d" % (
            c, int((got != want).sum()), int(np.abs(got.astype(int) - want.astype(int)).max()))


@pytest.mark.parametrize("letterbox", [True, False])
def test_device_unletterbox_matches_numpy_bit_exactly(cuda, letterbox):
    from dcfa_b200 import engine
    from utils.utils_bbox import DecodeBox
    b, a, k = 5, 300, 64
    g = torch.Generator().manual_seed(3)
    xy = torch.rand(b, a, 2, generator=g)
    wh = torch.rand(b, a, 2, generator=g) * 0.4
    det = torch.cat([xy - wh / 2, xy + wh / 2, torch.rand(b, a, 1, generator=g), torch.randint(0, 3, (b, a, 1), generator=g).float()], 2)
    cnt = torch.tensor([0, 1, 64, 17, 300], dtype=torch.int32)
    shapes = np.array([[480, 640], [1080, 1920], [333, 517], [640, 640], [97, 1003]])

    class WS:
        pass
    ws = WS()
    ws.b, ws.a, ws.det, ws.cnt = b, a, det.to(cuda), cnt.to(cuda)
    out = torch.empty(b, 1 + 6 * k, dtype=torch.float32, device=cuda)
    hw = torch.from_numpy(shapes.astype(np.int32)).to(cuda)
    engine.pack_detections(ws, k, out, hw, (640, 512), letterbox)
    got = out.cpu().numpy()
    dec = DecodeBox(3, (640, 512))
    for i in range(b):
        n = min(int(cnt[i]), k)
        assert got[i, 0] == float(cnt[i])
        rows = got[i, 1:].reshape(k, 6)
        d = det[i, :n].numpy().copy()
        if n:
            box_xy, box_wh = (d[:, 0:2] + d[:, 2:4]) / 2, d[:, 2:4] - d[:, 0:2]     # utils/utils_bbox.py:172
            d[:, :4] = dec.yolo_correct_boxes(box_xy, box_wh, (640, 512), shapes[i], letterbox)
            assert np.array_equal(rows[:n], d), "image %d" % i
        assert not rows[n:].any()
    # without shapes: rows pass through unchanged
    engine.pack_detections(ws, k, out, None, None, True)
    raw = out.cpu().numpy()
    assert np.array_equal(raw[2, 1:].reshape(k, 6), det[2, :k].numpy())


def test_facade_device_letterbox_equals_host_pil_path(cuda):
    """YOLO.detect_images with the letterbox / un-letterbox on the device gives exactly the rows of the PIL + numpy path."""
    from test_facade_gpu import _facade, _images
    sizes = [(150, 100), (90, 200), (128, 128), (300, 77), (640, 480)]
    rgb, nir = _images(sizes, 1), _images(sizes, 2)
    res = {}
    for dev_lb in (True, False):
        yolo = _facade(phi="n", input_shape=[128, 128], confidence=0.5, nms_iou=0.3, device_letterbox=dev_lb)
        res[dev_lb] = yolo.detect_images(rgb, nir)
        # single-channel depth frames: the device path keeps ONE plane, the host path replicates it (cvtColor)
        res[(dev_lb, "L")] = yolo.detect_images(rgb, [im.convert("L") for im in nir])
    for key in (True, (True, "L")):
        other = False if key is True else (False, "L")
        assert any(r is not None for r in res[key])
        for a, b in zip(res[key], res[other]):
            assert (a is None) == (b is None)
            if a is not None:
                assert np.array_equal(a, b)
