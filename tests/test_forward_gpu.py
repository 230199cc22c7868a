"""End-to-end parity of the drop-in API on a B200 against the golden vectors of the real reference
(tests/golden/*.npz, produced by oracle/make_golden.py) and against the CPU oracle.

Tolerances (BASELINE.json north_star): head maps / cls / dbox  |err| <= 2e-2 + 1e-2*|ref| (bf16 path vs fp32
reference); decode <= 1e-5 on the reference's own head outputs; NMS rows bit-exact on the reference's decoded boxes.
"""
import contextlib
import io

import numpy as np
import pytest
import torch

from test_oracle_cpu import GOLDEN_CASES, compare_x_maps, golden_state_dict, load_golden

pytestmark = pytest.mark.gpu


def build_model(meta, sd, device):
    from nets.yolo_mul import YoloBody
    with contextlib.redirect_stdout(io.StringIO()):
        net = YoloBody([meta["H"], meta["W"]], meta["nc"], meta["phi"])
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return net.to(device).eval()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_forward_matches_reference_golden(cuda, name):
    from oracle import forward as O
    z, meta, keys = load_golden(name)
    sd = golden_state_dict(meta, keys)
    net = build_model(meta, sd, cuda)
    rgb, nir = O.synth_inputs(meta["B"], meta["H"], meta["W"], meta["seed"] + 1000)
    dbox, cls, x, anchors, strides = net(rgb.to(cuda), nir.to(cuda))
    torch.cuda.synchronize()
    # API contract (reference nets/yolo_mul.py:462): shapes, dtypes, layouts
    b, a, nc = meta["B"], z["dbox"].shape[-1], meta["nc"]
    assert dbox.shape == (b, 4, a) and cls.shape == (b, nc, a) and anchors.shape == (2, a) and strides.shape == (1, a)
    assert all(t.dtype == torch.float32 for t in [dbox, cls, anchors, strides] + list(x))
    assert all(t.is_contiguous() for t in x) and dbox.is_contiguous()
    assert np.array_equal(anchors.cpu().numpy(), z["anchors"]) and np.array_equal(strides.cpu().numpy(), z["strides"])
    compare_x_maps(z, x, 2e-2, 1e-2)
    np.testing.assert_allclose(cls.cpu().numpy(), z["cls"], atol=2e-2, rtol=1e-2)
    np.testing.assert_allclose(dbox.cpu().numpy(), z["dbox"], atol=2e-2, rtol=1e-2)
    # the raw maps and the gathered (dbox-input, cls) views are consistent (nets/yolo_mul.py:459-460)
    cat = torch.cat([xi.view(b, 64 + nc, -1) for xi in x], 2)
    assert torch.equal(cat[:, 64:], cls)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_decode_and_nms_on_reference_outputs(cuda, name):
    """decode_box on the reference's head outputs (1e-5) and non_max_suppression on the reference's decoded
    boxes (bit-exact rows, identical kept set to the C oracle)."""
    from oracle import nms as onms
    from utils.utils_bbox import DecodeBox
    z, meta, _ = load_golden(name)
    dec = DecodeBox(meta["nc"], (meta["H"], meta["W"]))
    inputs = tuple(torch.from_numpy(z[k]).to(cuda) for k in ("dbox", "cls")) + (None,) + tuple(
        torch.from_numpy(z[k]).to(cuda) for k in ("anchors", "strides"))
    y = dec.decode_box(inputs)
    assert y.shape == z["decoded"].shape
    np.testing.assert_allclose(y.cpu().numpy(), z["decoded"], atol=1e-5, rtol=0)

    pred = torch.from_numpy(z["decoded"].copy()).to(cuda)
    res = dec.non_max_suppression(pred, meta["nc"], [meta["H"], meta["W"]], np.array([meta["H"], meta["W"]]), True,
                                  conf_thres=meta["conf"], nms_thres=meta["iou"])
    # side effect of the reference (utils/utils_bbox.py:97): boxes become corners in place
    d = z["decoded"]
    assert np.array_equal(pred[..., 0].cpu().numpy(), d[..., 0] - d[..., 2] / 2)
    for i, r in enumerate(res):
        g = z["nms%d" % i]
        if r is None:
            assert g.shape[0] == 0
            continue
        assert r.shape == g.shape, "image %d: kept %s vs reference %s" % (i, r.shape, g.shape)
        np.testing.assert_array_equal(r[:, 4:], g[:, 4:])
        np.testing.assert_allclose(r[:, :4], g[:, :4], atol=1e-4, rtol=1e-6)
    # kept anchor indices vs the C oracle
    ws = dec.nms_device(torch.from_numpy(z["decoded"].copy()).to(cuda), meta["conf"], meta["iou"])
    odet, oidx, ocnt, _ = onms.nms_raw(z["decoded"].copy(), meta["conf"], meta["iou"], 0)
    cnt = ws.cnt.cpu().numpy()
    assert np.array_equal(cnt, ocnt)
    for i in range(len(cnt)):
        assert np.array_equal(ws.idx[i, :cnt[i]].cpu().numpy(), oidx[i, :cnt[i]])


def _check_against_golden(z, meta, out, sel):
    """dbox / cls / head maps of the batch rows `sel` against the golden pair(s), north_star tolerance."""
    dbox, cls, x = out[0], out[1], out[2]
    np.testing.assert_allclose(dbox[sel].cpu().numpy(), z["dbox"], atol=2e-2, rtol=1e-2)
    np.testing.assert_allclose(cls[sel].cpu().numpy(), z["cls"], atol=2e-2, rtol=1e-2)
    compare_x_maps(z, [xi[sel] for xi in x], 2e-2, 1e-2)


@pytest.mark.parametrize("name,batch,positions", [("s640_stress", 32, (0, 13, 30)), ("l1280_stress", 4, (0, 3))])
def test_golden_pairs_inside_the_bench_batch(cuda, name, batch, positions):
    """Numeric parity at the batch composition the bench runs (BASELINE.json configs[1]: s, 640x640, B=32; configs[3]:
    l, 1280x1280): the golden pair(s) of the real reference are placed at several positions of a full batch of other
    images, so the tile schedule, cluster sizes and fused/unfused kernel choices of the real run are the ones compared.
    Head maps, cls and dbox within 2e-2 + 1e-2*|ref|; decode + NMS of the batch rows reproduce the reference's rows."""
    from oracle import forward as O
    from utils.utils_bbox import DecodeBox
    z, meta, keys = load_golden(name)
    sd = golden_state_dict(meta, keys)
    net = build_model(meta, sd, cuda)
    h, w, gb = meta["H"], meta["W"], meta["B"]
    grgb, gnir = O.synth_inputs(gb, h, w, meta["seed"] + 1000)
    g = torch.Generator().manual_seed(99)
    rgb = torch.rand(batch, 3, h, w, generator=g)
    nir = torch.rand(batch, 3, h, w, generator=g)
    for p0 in positions:
        rgb[p0:p0 + gb], nir[p0:p0 + gb] = grgb, gnir
    out = net(rgb.to(cuda), nir.to(cuda))
    torch.cuda.synchronize()
    for p0 in positions:
        _check_against_golden(z, meta, out, slice(p0, p0 + gb))
    # identical pairs give identical results wherever they sit in the batch
    for p0 in positions[1:]:
        assert torch.equal(out[0][p0:p0 + gb], out[0][positions[0]:positions[0] + gb])
    dec = DecodeBox(meta["nc"], (h, w))
    y = dec.decode_box(out)
    p0 = positions[-1]
    np.testing.assert_allclose(y[p0:p0 + gb].cpu().numpy(), z["decoded"], atol=2e-3, rtol=1e-2)
    # NMS of OUR decoded boxes: bit-exact against the C oracle fed the same tensor (the golden scores of these weights
    # sit within the bf16 tolerance of the confidence threshold, so the kept SET is only comparable on identical inputs;
    # NMS on the reference's own decoded boxes is test_decode_and_nms_on_reference_outputs, bit-exact, same goldens)
    from oracle import nms as onms
    y_host = y.cpu().numpy().copy()
    res = dec.non_max_suppression(y, meta["nc"], [h, w], np.array([h, w]), True, conf_thres=meta["conf"], nms_thres=meta["iou"])
    ref = onms.non_max_suppression(y_host, [h, w], np.array([h, w]), True, meta["conf"], meta["iou"], 0)
    assert len(res) == batch
    for r, o in zip(res, ref):
        assert (r is None) == (o is None)
        if r is not None:
            assert r.shape == o.shape and np.array_equal(r[:, 4:], o[:, 4:])
            np.testing.assert_allclose(r[:, :4], o[:, :4], atol=1e-3, rtol=1e-6)


@pytest.mark.parametrize("var,val", [("DCFA_CBAM_FUSED", "0"), ("DCFA_CBAM_FUSED", "2"), ("DCFA_CHAIN", "0"), ("DCFA_CHAIN", "2"),
                                     ("DCFA_GHOST", "0"), ("DCFA_GHOST", "2"), ("DCFA_PDL", "1"), ("DCFA_SPPF_FUSED", "0"), ("DCFA_CONV_FAST", "0")])
def test_kernel_path_switches_end_to_end_at_640(cuda, monkeypatch, var, val):
    """Every alternative kernel path (four-kernel / always-fused CBAM, three-kernel / always-fused ShuffleNet branch)
    against the s@640x640 golden of the real reference, end to end."""
    from oracle import forward as O
    monkeypatch.setenv(var, val)
    z, meta, keys = load_golden("s640_stress")
    sd = golden_state_dict(meta, keys)
    net = build_model(meta, sd, cuda)
    rgb, nir = O.synth_inputs(meta["B"], meta["H"], meta["W"], meta["seed"] + 1000)
    out = net(rgb.to(cuda), nir.to(cuda))
    torch.cuda.synchronize()
    _check_against_golden(z, meta, out, slice(0, meta["B"]))


def test_full_pipeline_vs_oracle_and_shard_invariance(cuda):
    """forward -> decode -> NMS through the public API, vs the fp32 CPU oracle on the same weights; any batch
    shard gives bit-identical results (the multi-GPU path shards the batch with no collective)."""
    from oracle import forward as O
    from utils.utils_bbox import DecodeBox
    z, meta, keys = load_golden("s128_stress")
    sd = golden_state_dict(meta, keys)
    net = build_model(meta, sd, cuda)
    rgb, nir = O.synth_inputs(4, 128, 128, 77)
    out4 = net(rgb.to(cuda), nir.to(cuda))
    out2 = net(rgb[2:].to(cuda), nir[2:].to(cuda))
    assert torch.equal(out4[0][2:], out2[0]) and torch.equal(out4[1][2:], out2[1])
    for a4, a2 in zip(out4[2], out2[2]):
        assert torch.equal(a4[2:], a2)
    ref = O.yolo_forward(sd, "s", rgb, nir, 1)
    np.testing.assert_allclose(out4[0].cpu().numpy(), ref[0].numpy(), atol=2e-2, rtol=1e-2)
    dec = DecodeBox(1, (128, 128))
    y = dec.decode_box(out4)
    yref = O.decode_box(ref, (128, 128))
    np.testing.assert_allclose(y.cpu().numpy(), yref.numpy(), atol=2e-3, rtol=1e-2)
    res = dec.non_max_suppression(y, 1, [128, 128], np.array([128, 128]), True, conf_thres=0.5, nms_thres=0.3)
    assert len(res) == 4


def test_uint8_nhwc_inputs_match_oracle_on_preprocessed_pixels(cuda):
    """net(uint8 [B,H,W,3]) == oracle(preprocess_input(pixels) as NCHW): the /255 and the transpose of
    yolo_mul.py:76 run inside the stem kernel."""
    from oracle import forward as O
    z, meta, keys = load_golden("s128_stress")
    sd = golden_state_dict(meta, keys)
    net = build_model(meta, sd, cuda)
    g = torch.Generator().manual_seed(11)
    rgb8 = torch.randint(0, 256, (3, 128, 128, 3), generator=g, dtype=torch.uint8)
    nir8 = torch.randint(0, 256, (3, 128, 128, 3), generator=g, dtype=torch.uint8)
    out = net(rgb8.to(cuda), nir8.to(cuda))
    ref = O.yolo_forward(sd, "s", rgb8.permute(0, 3, 1, 2).float() / 255.0, nir8.permute(0, 3, 1, 2).float() / 255.0, 1)
    np.testing.assert_allclose(out[0].cpu().numpy(), ref[0].numpy(), atol=2e-2, rtol=1e-2)
    np.testing.assert_allclose(out[1].cpu().numpy(), ref[1].numpy(), atol=2e-2, rtol=1e-2)
    # and it agrees with the fp32 entry point fed the same pixels (only the input rounding differs)
    out_f = net((rgb8.permute(0, 3, 1, 2).float() / 255.0).to(cuda), (nir8.permute(0, 3, 1, 2).float() / 255.0).to(cuda))
    np.testing.assert_allclose(out[0].cpu().numpy(), out_f[0].cpu().numpy(), atol=4e-2, rtol=2e-2)
    with pytest.raises(ValueError):
        net(rgb8.permute(0, 3, 1, 2).contiguous().to(cuda), nir8.permute(0, 3, 1, 2).contiguous().to(cuda))
    # the depth image as the single plane cvtColor would replicate (utils/utils.py:14-19): bit-identical outputs
    plane = nir8[..., 0].contiguous()
    rep = plane[..., None].expand(-1, -1, -1, 3).contiguous()
    out_rep = net(rgb8.to(cuda), rep.to(cuda))
    for shape in (plane, plane[..., None]):
        out_pl = net(rgb8.to(cuda), shape.to(cuda))
        assert torch.equal(out_pl[0], out_rep[0]) and torch.equal(out_pl[1], out_rep[1])


def test_full_size_properties_at_the_bench_workload(cuda):
    """BASELINE.json configs[1] shape (s, 640x640, constructor-init weights, ~8000 NMS candidates per image), checked
    through size-independent properties: batch-shard invariance of every output (the multi-GPU path shards the
    batch), kept rows sorted by descending score per image, NMS idempotence (the kept set is a fixed point), no kept
    pair above the IoU threshold, and agreement of the uint8 and fp32 entry points on the same pixels."""
    from nets.yolo_mul import YoloBody
    from utils.utils_bbox import DecodeBox
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        net = YoloBody([640, 640], 1, 's').eval().to(cuda)
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (2, 6, 640, 640, 3), generator=g, dtype=torch.uint8)
    rgb = (u8[0].permute(0, 3, 1, 2).float() / 255.0).to(cuda)
    nir = (u8[1].permute(0, 3, 1, 2).float() / 255.0).to(cuda)
    full = net(rgb, nir)
    part = net(rgb[4:], nir[4:])
    assert torch.equal(full[0][4:], part[0]) and torch.equal(full[1][4:], part[1])
    for a, b in zip(full[2], part[2]):
        assert torch.equal(a[4:], b)
    dec = DecodeBox(1, (640, 640))
    y = dec.decode_box(full)
    ws = dec.nms_device(y.clone(), 0.5, 0.3)
    cnt, cand = ws.cnt.cpu().numpy(), ws.cand.cpu().numpy()
    assert cand.min() > 1000 and cnt.min() > 0          # the heavy-NMS regime the bench runs in
    det = ws.det.cpu().numpy()
    for i in range(6):
        rows = det[i, :cnt[i]]
        assert np.all(np.diff(rows[:, 4]) <= 0), "kept rows are not in descending score order"
        x1, y1, x2, y2 = rows[:, 0:1], rows[:, 1:2], rows[:, 2:3], rows[:, 3:4]
        iw = np.clip(np.minimum(x2, x2.T) - np.maximum(x1, x1.T), 0, None)
        ih = np.clip(np.minimum(y2, y2.T) - np.maximum(y1, y1.T), 0, None)
        area = (x2 - x1) * (y2 - y1)
        iou = iw * ih / (area + area.T - iw * ih)
        np.fill_diagonal(iou, 0.0)
        assert iou.max() <= 0.3 + 1e-6, "two kept boxes overlap above the threshold"
    # idempotence: feed the kept boxes (as xywh rows with their scores) back in -> all of them are kept again
    k = int(cnt.max())
    again = torch.zeros(6, k, 5, device=cuda)
    for i in range(6):
        r = ws.det[i, :cnt[i]]
        again[i, :cnt[i], 0] = (r[:, 0] + r[:, 2]) / 2
        again[i, :cnt[i], 1] = (r[:, 1] + r[:, 3]) / 2
        again[i, :cnt[i], 2] = r[:, 2] - r[:, 0]
        again[i, :cnt[i], 3] = r[:, 3] - r[:, 1]
        again[i, :cnt[i], 4] = r[:, 4]
    ws2 = DecodeBox(1, (640, 640)).nms_device(again, 0.5, 0.3)
    assert np.array_equal(ws2.cnt.cpu().numpy(), cnt)
    # uint8 entry point on the same pixels
    out8 = net(u8[0].to(cuda), u8[1].to(cuda))
    np.testing.assert_allclose(out8[0].cpu().numpy(), full[0].cpu().numpy(), atol=4e-2, rtol=2e-2)


def test_no_cpu_fallback_and_training_mode(cuda):
    from nets.yolo_mul import YoloBody
    from utils.utils_bbox import DecodeBox
    with contextlib.redirect_stdout(io.StringIO()):
        net = YoloBody([64, 64], 1, 'n')
    x = torch.rand(1, 3, 64, 64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.eval()(x, x)
    with pytest.raises(NotImplementedError):
        net.train().to(cuda)(x.to(cuda), x.to(cuda))
    with pytest.raises(RuntimeError, match="no CPU path"):
        DecodeBox(1, (64, 64)).non_max_suppression(torch.rand(1, 84, 5), 1, [64, 64], np.array([64, 64]), True)
    # default-init constructor path runs end to end and invalidation picks up new weights
    net = net.eval()
    out_a = net(x.to(cuda), x.to(cuda))[0].clone()
    sd = net.state_dict()
    sd['cv2.0.2.bias'] = sd['cv2.0.2.bias'] + 1.0
    net.load_state_dict(sd)
    out_b = net(x.to(cuda), x.to(cuda))[0]
    assert not torch.equal(out_a, out_b)
