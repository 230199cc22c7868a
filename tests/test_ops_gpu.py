"""Op-level parity of every CUDA kernel against a plain PyTorch fp32 reference of the same op
(inputs/weights rounded to bf16 first, so only accumulation order and the final bf16 store differ)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import bf16_round, conv_op, flat_view, nhwc_view, ptrs

pytestmark = pytest.mark.gpu


def _run(ops, bufs):
    from dcfa_b200 import _lib
    _lib.run_ops(ops, ptrs(bufs), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()


def _launches():
    from dcfa_b200 import _lib
    return int(_lib.lib.dcfa_launch_count())


def _act(v, act):
    from dcfa_b200 import abi
    if act == abi.ACT_RELU:
        return F.relu(v)
    if act == abi.ACT_SILU:
        return v * torch.sigmoid(v)
    return v


def _bf16_close(got, ref, what):
    got, ref = got.float(), ref.float()
    tol = 1e-2 * ref.abs() + 2e-2 * max(ref.abs().max().item(), 1e-6) / 4
    bad = (got - ref).abs() > tol
    assert not bad.any(), "%s: %d mismatches, max abs err %.4g (ref max %.4g)" % (
        what, int(bad.sum()), (got - ref).abs().max().item(), ref.abs().max().item())


CONV_CASES = [
    # n, h, w, cin, cout, k, s, act, groups
    (2, 8, 8, 64, 64, 1, 1, 2, 1),      # one tile, one k-block
    (1, 16, 16, 32, 32, 1, 1, 1, 1),    # Cin < 64: K padding
    (2, 20, 20, 128, 64, 1, 1, 2, 1),   # M tail (800 rows), 2 k-blocks
    (2, 16, 16, 64, 128, 3, 1, 2, 1),   # 3x3, padding
    (2, 16, 16, 32, 64, 3, 2, 2, 1),    # 3x3 stride 2, taps straddling 64-wide k-blocks
    (1, 12, 20, 16, 16, 3, 2, 0, 1),    # tiny channels (phi=n), N=16
    (2, 10, 10, 256, 512, 1, 1, 2, 1),  # two N tiles of 256
    (4, 12, 12, 64, 192, 3, 1, 2, 2),   # two weight groups (modalities), N=192
    (3, 40, 40, 128, 128, 3, 1, 2, 1),  # many tiles > pipeline depth, persistent loop
    (2, 9, 7, 48, 80, 3, 1, 1, 1),      # odd sizes, Cin not a multiple of 64
    (2, 160, 160, 32, 64, 3, 2, 2, 1),  # backbone-like: 64-byte swizzle, stride 2, wide tiles
    (1, 21, 37, 64, 64, 3, 2, 2, 1),    # odd input size with stride 2
    (2, 20, 20, 512, 256, 1, 1, 2, 1),  # 20x20 maps: partial spatial tiles
    (2, 40, 40, 128, 128, 3, 2, 2, 2),  # stride 2, two k-blocks per tap, two groups
    (1, 22, 38, 64, 32, 3, 2, 0, 1),    # stride 2, even width, partial tiles on the right / bottom edge
    (4, 20, 20, 64, 64, 1, 1, 2, 2),    # 1x1, two weight groups: flattened pixel tiling, 7 tiles per group (partial last)
    (6, 40, 40, 128, 256, 1, 1, 2, 2),  # 1x1 at 40x40 (the neck's shape): 38 flattened tiles per group instead of 3 x 14
    (2, 80, 80, 128, 192, 3, 1, 2, 1),  # head0.0's shape: halo-strip path, 5-row strips, 4 weight stages
    (2, 20, 20, 512, 192, 3, 1, 2, 1),  # head2.0's shape: 8 channel blocks per tile, 9-row strips
    (3, 13, 27, 64, 64, 3, 1, 2, 1),    # strip path, odd sizes: tiles that start and end inside a row, N = 64 (groups alternate tiles)
    (2, 40, 40, 256, 128, 3, 1, 2, 2),  # strip path, two weight groups, 4 channel blocks
]


@pytest.mark.parametrize("path", ["tma", "tma-staged", "gather", "tma-pairs", "tma-nostrip"])
@pytest.mark.parametrize("n,h,w,cin,cout,k,s,act,groups", CONV_CASES)
def test_conv_bf16_nhwc(cuda, monkeypatch, n, h, w, cin, cout, k, s, act, groups, path):
    """both A-operand paths: TMA box loads (one k-block per tap x channel block) and the cp.async gather (flat K);
    and both store paths of the TMA kernel: 256-bit sector stores (default) and the staged tile + TMA store
    ("tma-staged": DCFA_ST256=0, the path taken by outputs that are not 32-byte aligned); "tma-pairs": the opt-in
    CTA-pair schedule with multicast weight tiles (DCFA_CONV_MC); "tma-nostrip": DCFA_CONV_STRIP=0 (by default the 3x3 stride-1
    SiLU cases with Cin % 64 == 0 run on the halo-strip kernel)."""
    from dcfa_b200 import abi
    if path == "tma-staged":
        monkeypatch.setenv("DCFA_ST256", "0")
        path = "tma"
    if path == "tma-nostrip":   # 3x3 stride-1 SiLU layers on the tap-box kernel instead of the halo-strip kernel (conv_strip.cu)
        monkeypatch.setenv("DCFA_CONV_STRIP", "0")
        path = "tma"
    if path == "tma-pairs":   # CTA pairs (cluster of 2) multicasting each W k-block to both rings (opt-in, DCFA_CONV_MC)
        monkeypatch.setenv("DCFA_CONV_MC", "16")
        path = "tma"
    g = torch.Generator().manual_seed(1234 + cin + cout + k + s)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    ws = [bf16_round(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5) for _ in range(groups)]
    scs = [torch.rand(cout, generator=g) + 0.5 for _ in range(groups)]
    bis = [torch.randn(cout, generator=g) * 0.1 for _ in range(groups)]
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(cuda)
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    y = torch.full((n, ho, wo, cout), 7.0, dtype=torch.bfloat16, device=cuda)
    gi = n // groups
    op, bufs = conv_op(x_nhwc, ws, scs, bis, y, ksize=k, stride=s, act=act, cin=cin, group_imgs=gi,
                       bk=None if path == "tma" else 0)
    assert (op.flags != 0) == (path == "tma")
    _run([op], bufs)
    refs = []
    for gg in range(groups):
        r = F.conv2d(x[gg * gi:(gg + 1) * gi].to(cuda), ws[gg].to(cuda), None, s, pad)
        r = r * scs[gg].to(cuda)[None, :, None, None] + bis[gg].to(cuda)[None, :, None, None]
        refs.append(_act(r, act))
    ref = torch.cat(refs).permute(0, 2, 3, 1)
    _bf16_close(y, ref, "conv")


@pytest.mark.parametrize("n,h,w,cin,cout,groups", [(4, 20, 20, 64, 64, 2), (3, 40, 40, 128, 256, 1), (2, 10, 10, 256, 512, 1)])
def test_conv_1x1_flattened_tiling_equals_spatial(cuda, monkeypatch, n, h, w, cin, cout, groups):
    """1x1 stride-1 convs tile the flattened pixel axis of each weight group (launch_conv_tma); DCFA_CONV_FLAT=0 keeps the
    spatial tiles.  Same MMAs on the same rows in a different tile order: the outputs must be bit-identical."""
    from dcfa_b200 import abi
    g = torch.Generator().manual_seed(77 + cin + cout)
    x = bf16_round(torch.randn(n, cin, h, w, generator=g)).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(cuda)
    ws = [bf16_round(torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5) for _ in range(groups)]
    scs = [torch.rand(cout, generator=g) + 0.5 for _ in range(groups)]
    bis = [torch.randn(cout, generator=g) * 0.1 for _ in range(groups)]
    outs = []
    for flat in ("1", "0"):
        monkeypatch.setenv("DCFA_CONV_FLAT", flat)
        y = torch.full((n, h, w, cout), 7.0, dtype=torch.bfloat16, device=cuda)
        op, bufs = conv_op(x, ws, scs, bis, y, ksize=1, stride=1, act=abi.ACT_SILU, cin=cin, group_imgs=n // groups)
        _run([op], bufs)
        outs.append(y)
    assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16))


def test_conv_views_residual_postscale(cuda):
    """channel-offset input view, channel-offset output slot, residual add and post-activation scale."""
    from dcfa_b200 import abi
    g = torch.Generator().manual_seed(7)
    n, h, w, ctot, c_off, cin, cout = 2, 14, 14, 96, 32, 64, 32
    x = bf16_round(torch.randn(n, h, w, ctot, generator=g))
    wt = bf16_round(torch.randn(cout, cin, 1, 1, generator=g) / 8)
    sc, bi = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    res = bf16_round(torch.randn(n, h, w, 48, generator=g))
    y = torch.full((n, h, w, 80), 3.0, dtype=torch.bfloat16, device=cuda)
    op, bufs = conv_op(x.to(torch.bfloat16).to(cuda), [wt], [sc], [bi], y, ksize=1, stride=1, act=abi.ACT_SILU, cin=cin,
                       c_off_in=c_off, c_off_out=40, res=res.to(torch.bfloat16).to(cuda), c_off_res=16, post_scale=0.5)
    _run([op], bufs)
    xin = x[..., c_off:c_off + cin].permute(0, 3, 1, 2)
    r = F.conv2d(xin, wt) * sc[None, :, None, None] + bi[None, :, None, None]
    r = (r * torch.sigmoid(r)) * 0.5 + res[..., 16:16 + cout].permute(0, 3, 1, 2)
    _bf16_close(y[..., 40:72].cpu(), r.permute(0, 2, 3, 1), "conv view")
    assert (y[..., :40] == 3.0).all() and (y[..., 72:] == 3.0).all(), "conv wrote outside its channel slot"


def test_conv_stride2_on_channel_subview(cuda):
    """stride-2 3x3 conv reading a channel sub-view (Cin < ld): the element-strided tensor map must step over the
    neighbouring channels (down1/down2 read P3/P4 out of wider buffers)."""
    from dcfa_b200 import abi
    g = torch.Generator().manual_seed(17)
    n, h, w, ctot, c_off, cin, cout = 2, 24, 36, 160, 64, 64, 96
    x = bf16_round(torch.randn(n, h, w, ctot, generator=g))
    wt = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) / 24)
    sc, bi = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    y = torch.full((n, h // 2, w // 2, cout), 3.0, dtype=torch.bfloat16, device=cuda)
    op, bufs = conv_op(x.to(torch.bfloat16).to(cuda), [wt], [sc], [bi], y, ksize=3, stride=2, act=abi.ACT_SILU, cin=cin,
                       c_off_in=c_off)
    assert op.flags != 0
    _run([op], bufs)
    r = F.conv2d(x[..., c_off:c_off + cin].permute(0, 3, 1, 2), wt, None, 2, 1) * sc[None, :, None, None] + bi[None, :, None, None]
    _bf16_close(y.cpu(), (r * torch.sigmoid(r)).permute(0, 2, 3, 1), "conv s2 sub-view")


@pytest.mark.parametrize("n,h,w,cout,groups", [(2, 16, 16, 64, 1), (4, 160, 160, 64, 2), (1, 23, 38, 48, 1), (2, 40, 36, 128, 1)])
def test_conv_stride2_pixel_pair_blocks(cuda, n, h, w, cout, groups):
    """DCFA_CONV_FLAG_PAIR: 3x3 stride-2 conv with Cin = 32 read as six 64-channel pixel-pair k-blocks per tile."""
    from dcfa_b200 import abi, pack
    g = torch.Generator().manual_seed(77 + cout)
    cin, gi = 32, n // groups
    x = bf16_round(torch.randn(n, cin, h, w, generator=g))
    ws = [bf16_round(torch.randn(cout, cin, 3, 3, generator=g) / 17.0) for _ in range(groups)]
    scs = [torch.rand(cout, generator=g) + 0.5 for _ in range(groups)]
    bis = [torch.randn(cout, generator=g) * 0.1 for _ in range(groups)]
    packed = [pack.pack_conv_weight_pair(t) for t in ws]
    m = packed[0][1]
    npad = m["BN"] * m["n_tiles"]
    xg = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(cuda)
    W = torch.stack([t[0] for t in packed]).to(cuda)
    S = torch.stack([pack.pad_channels(t, npad) for t in scs]).to(cuda)
    B = torch.stack([pack.pad_channels(t, npad) for t in bis]).to(cuda)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.full((n, ho, wo, cout), 7.0, dtype=torch.bfloat16, device=cuda)
    op = abi.new_op(abi.OP_CONV, act=abi.ACT_SILU, x=nhwc_view(xg, 0), w=flat_view(1), scale=flat_view(2), bias=flat_view(3),
                    y=nhwc_view(y, 4), n_img=n, group_imgs=gi, Hi=h, Wi=w, Cin=cin, Ho=ho, Wo=wo, Cout=cout, ksize=3, stride=2,
                    BN=m["BN"], n_tiles=m["n_tiles"], k_blocks=m["k_blocks"], K_real=m["K_real"], w_gstride=packed[0][0].numel(),
                    sb_gstride=npad, flags=m["bk"] | abi.CONV_FLAG_PAIR)
    _run([op], [xg, W, S, B, y])
    refs = []
    for gg in range(groups):
        r = F.conv2d(x[gg * gi:(gg + 1) * gi], ws[gg], None, 2, 1) * scs[gg].view(1, -1, 1, 1) + bis[gg].view(1, -1, 1, 1)
        refs.append(r * torch.sigmoid(r))
    _bf16_close(y.cpu(), torch.cat(refs).permute(0, 2, 3, 1), "conv pair")


def test_conv_f32_nchw_out(cuda):
    """head-style output: fp32 NCHW at a channel offset, Cout not a multiple of 16 (nc = 1 and 64)."""
    from dcfa_b200 import abi
    g = torch.Generator().manual_seed(11)
    n, h, w, cin = 2, 10, 10, 64
    x = bf16_round(torch.randn(n, h, w, cin, generator=g))
    out = torch.zeros(n, 65, h, w, dtype=torch.float32, device=cuda)
    xg = x.to(torch.bfloat16).to(cuda)
    for cout, coff in ((64, 0), (1, 64)):
        wt = bf16_round(torch.randn(cout, cin, 1, 1, generator=g) / 8)
        bi = torch.randn(cout, generator=g)
        op, bufs = conv_op(xg, [wt], [torch.ones(cout)], [bi], out, ksize=1, stride=1, act=abi.ACT_NONE, cin=cin,
                           out_mode=abi.OUT_F32_NCHW, out_ctot=65, out_coff=coff)
        _run([op], bufs)
        r = F.conv2d(x.permute(0, 3, 1, 2), wt, bi)
        got = out[:, coff:coff + cout].cpu()
        assert torch.allclose(got, r, atol=2e-4, rtol=1e-4), (got - r).abs().max()


@pytest.mark.parametrize("b,h,w,c0,groups", [(1, 32, 64, 16, 1), (2, 40, 72, 32, 2), (1, 70, 130, 32, 1), (3, 33, 47, 48, 2),
                                            (1, 64, 64, 80, 1), (2, 96, 128, 64, 2), (1, 640, 640, 32, 1), (2, 5, 12, 32, 2),
                                            (1, 100, 36, 128, 1)])
def test_stem(cuda, b, h, w, c0, groups):
    """fused conv3x3+BN+ReLU+maxpool3/2 vs F.conv2d + F.max_pool2d (nets/yolo_mul.py:104-115)."""
    from dcfa_b200 import abi
    g = torch.Generator().manual_seed(5)
    xs = [torch.rand(b, 3, h, w, generator=g) for _ in range(groups)]
    ws = [torch.randn(c0, 3, 3, 3, generator=g) * 0.3 for _ in range(groups)]
    bs = [torch.randn(c0, generator=g) * 0.2 for _ in range(groups)]
    from dcfa_b200 import pack
    ws = [bf16_round(wt) for wt in ws]
    xs = [bf16_round(t) for t in xs]
    scs = [torch.rand(c0, generator=g) - 0.3 for _ in range(groups)]   # some negative BN scales (sign folding)
    packed, sks, bks, c0pad = [], [], [], 32
    for i in range(groups):
        pk, sca, bia, c0pad = pack.pack_stem(ws[i], scs[i], bs[i])
        packed.append(pk); sks.append(sca); bks.append(bia)
    wk = torch.stack(packed).to(cuda)
    sk = torch.stack(sks).to(cuda)
    bk = torch.stack(bks).to(cuda)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.zeros(groups * b, ho, wo, c0, dtype=torch.bfloat16, device=cuda)
    xg = [t.to(cuda) for t in xs]
    bufs = [xg[0], xg[1] if groups == 2 else None, wk, bk, y, sk]
    op = abi.new_op(abi.OP_STEM, x=flat_view(0), x2=flat_view(1) if groups == 2 else abi.no_view(), w=flat_view(2),
                    bias=flat_view(3), scale=flat_view(5), y=nhwc_view(y, 4), n_img=groups * b, group_imgs=b, Hi=h, Wi=w,
                    Ho=ho, Wo=wo, Cout=c0, BN=c0pad, n_tiles=1, k_blocks=1, K_real=27, w_gstride=packed[0].numel(), sb_gstride=c0pad)
    _run([op], bufs)
    bs = [bs[i] for i in range(groups)]
    ws = [ws[i] * 1.0 for i in range(groups)]
    ref = torch.cat([F.max_pool2d(F.relu(F.conv2d(xs[i], ws[i], None, 1, 1) * scs[i].view(1, -1, 1, 1)
                                         + bs[i].view(1, -1, 1, 1)), 3, 2, 1) for i in range(groups)])
    _bf16_close(y.cpu(), ref.permute(0, 2, 3, 1), "stem")


@pytest.mark.parametrize("b,h,w,c0", [(2, 48, 64, 32), (1, 70, 130, 16), (2, 33, 47, 64), (1, 640, 640, 32)])
def test_stem_uint8_depth_plane(cuda, b, h, w, c0):
    """DCFA_STEM_FLAG_X2_PLANE: group 1 is ONE uint8 plane per image; bit-identical to the replicated 3-channel image
    (what cvtColor produces, utils/utils.py:14-19) through the plain uint8 path."""
    from dcfa_b200 import abi, pack
    g = torch.Generator().manual_seed(8)
    rgb = torch.randint(0, 256, (b, h, w, 3), generator=g, dtype=torch.uint8)
    plane = torch.randint(0, 256, (b, h, w), generator=g, dtype=torch.uint8)
    ws = [bf16_round(torch.randn(c0, 3, 3, 3, generator=g) * 0.3) for _ in range(2)]
    bs = [torch.randn(c0, generator=g) * 0.2 for _ in range(2)]
    scs = [torch.rand(c0, generator=g) - 0.3 for _ in range(2)]
    packed, sks, bks, c0pad = [], [], [], 32
    for i in range(2):
        pk, sca, bia, c0pad = pack.pack_stem(ws[i], scs[i], bs[i], u8=True)
        packed.append(pk); sks.append(sca); bks.append(bia)
    wk, sk, bk = torch.stack(packed).to(cuda), torch.stack(sks).to(cuda), torch.stack(bks).to(cuda)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    outs = []
    for flag, x2 in ((abi.STEM_FLAG_X2_PLANE, plane), (0, plane[..., None].expand(b, h, w, 3).contiguous())):
        y = torch.zeros(2 * b, ho, wo, c0, dtype=torch.bfloat16, device=cuda)
        bufs = [rgb.to(cuda), x2.to(cuda), wk, bk, y, sk]
        op = abi.new_op(abi.OP_STEM, x=flat_view(0), x2=flat_view(1), w=flat_view(2), bias=flat_view(3), scale=flat_view(5),
                        y=nhwc_view(y, 4), n_img=2 * b, group_imgs=b, Hi=h, Wi=w, Ho=ho, Wo=wo, Cout=c0, BN=c0pad, n_tiles=1,
                        k_blocks=1, K_real=27, w_gstride=packed[0].numel(), sb_gstride=c0pad, flags=abi.STEM_FLAG_U8 | flag)
        _run([op], bufs)
        outs.append(y.cpu())
    assert torch.equal(outs[0], outs[1])
    x1 = plane[:, None].expand(b, 3, h, w).float() / 255.0
    ref = F.max_pool2d(F.relu(F.conv2d(x1, ws[1], None, 1, 1) * scs[1].view(1, -1, 1, 1) + bs[1].view(1, -1, 1, 1)), 3, 2, 1)
    _bf16_close(outs[0][b:], ref.permute(0, 2, 3, 1), "stem_u8_plane")


@pytest.mark.parametrize("b,h,w,c0,groups", [(1, 32, 64, 16, 1), (2, 40, 80, 32, 2), (1, 70, 130, 32, 1), (3, 33, 47, 48, 2),
                                            (2, 96, 128, 64, 2), (1, 640, 640, 32, 1), (1, 64, 64, 80, 1), (2, 9, 7, 32, 2)])
def test_stem_uint8_nhwc(cuda, b, h, w, c0, groups):
    """DCFA_STEM_FLAG_U8: raw uint8 NHWC pixels; the kernel folds preprocess_input's /255 (utils/utils.py:76-79).
    Widths whose rows are 16-byte multiples take the TMA path, the others the plain-load path."""
    from dcfa_b200 import abi, pack
    g = torch.Generator().manual_seed(6)
    xs = [torch.randint(0, 256, (b, h, w, 3), generator=g, dtype=torch.uint8) for _ in range(groups)]
    ws = [bf16_round(torch.randn(c0, 3, 3, 3, generator=g) * 0.3) for _ in range(groups)]
    bs = [torch.randn(c0, generator=g) * 0.2 for _ in range(groups)]
    scs = [torch.rand(c0, generator=g) - 0.3 for _ in range(groups)]
    packed, sks, bks, c0pad = [], [], [], 32
    for i in range(groups):
        pk, sca, bia, c0pad = pack.pack_stem(ws[i], scs[i], bs[i], u8=True)
        packed.append(pk); sks.append(sca); bks.append(bia)
    wk, sk, bk = torch.stack(packed).to(cuda), torch.stack(sks).to(cuda), torch.stack(bks).to(cuda)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.zeros(groups * b, ho, wo, c0, dtype=torch.bfloat16, device=cuda)
    xg = [t.to(cuda) for t in xs]
    bufs = [xg[0], xg[1] if groups == 2 else None, wk, bk, y, sk]
    op = abi.new_op(abi.OP_STEM, x=flat_view(0), x2=flat_view(1) if groups == 2 else abi.no_view(), w=flat_view(2),
                    bias=flat_view(3), scale=flat_view(5), y=nhwc_view(y, 4), n_img=groups * b, group_imgs=b, Hi=h, Wi=w,
                    Ho=ho, Wo=wo, Cout=c0, BN=c0pad, n_tiles=1, k_blocks=1, K_real=27, w_gstride=packed[0].numel(), sb_gstride=c0pad,
                    flags=abi.STEM_FLAG_U8)
    _run([op], bufs)
    ref = torch.cat([F.max_pool2d(F.relu(F.conv2d(xs[i].permute(0, 3, 1, 2).float() / 255.0, ws[i], None, 1, 1)
                                         * scs[i].view(1, -1, 1, 1) + bs[i].view(1, -1, 1, 1)), 3, 2, 1) for i in range(groups)])
    _bf16_close(y.cpu(), ref.permute(0, 2, 3, 1), "stem_u8")


@pytest.mark.parametrize("n,h,w,c,act,groups,use_res", [(2, 9, 11, 16, 0, 1, False), (4, 20, 20, 64, 2, 2, True),
                                                      (2, 13, 6, 128, 2, 1, True), (2, 40, 40, 32, 0, 2, False),
                                                      (3, 33, 50, 256, 2, 1, True), (2, 160, 160, 32, 0, 2, False),
                                                      (2, 21, 19, 24, 2, 1, True)])
def test_dwconv(cuda, n, h, w, c, act, groups, use_res):
    from dcfa_b200 import abi
    g = torch.Generator().manual_seed(9)
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    ws = [torch.randn(c, 1, 3, 3, generator=g) * 0.3 for _ in range(groups)]
    bs = [torch.randn(c, generator=g) * 0.2 for _ in range(groups)]
    res = bf16_round(torch.randn(n, c, h, w, generator=g))
    wk = torch.stack([wt.reshape(c, 9).t().contiguous() for wt in ws]).to(cuda)  # [G][9][C]
    bk = torch.stack(bs).to(cuda)
    xg = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(cuda)
    rg = res.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(cuda)
    y = torch.zeros(n, h, w, c, dtype=torch.bfloat16, device=cuda)
    bufs = [xg, rg, wk, bk, y]
    op = abi.new_op(abi.OP_DWCONV, act=act, x=nhwc_view(xg, 0), x2=nhwc_view(rg, 1) if use_res else abi.no_view(),
                    w=flat_view(2), bias=flat_view(3), y=nhwc_view(y, 4), n_img=n, group_imgs=n // groups, Hi=h, Wi=w, Cin=c)
    _run([op], bufs)
    gi = n // groups
    ref = torch.cat([F.conv2d(x[i * gi:(i + 1) * gi], ws[i], bs[i], 1, 1, groups=c) for i in range(groups)])
    ref = _act(ref, act)
    if use_res:
        ref = ref + res
    _bf16_close(y.cpu(), ref.permute(0, 2, 3, 1), "dwconv")


@pytest.mark.parametrize("n,h,w,c,groups", [(2, 16, 32, 32, 1), (4, 21, 37, 64, 2), (2, 40, 40, 128, 2), (2, 160, 160, 32, 2),
                                            (1, 8, 16, 64, 1)])
def test_shuffle_branch_chain_fused_vs_three_kernels(cuda, monkeypatch, n, h, w, c, groups):
    """1x1 conv + BN + ReLU -> depthwise 3x3 + BN -> 1x1 conv + BN + ReLU (ShuffleNetV2 branch 2, nets/yolo_mul.py:
    138-162): the fused tcgen05 kernel must reproduce the three-kernel path (same bf16 rounding points) and both must
    match torch; the input is a channel sub-view, the output a channel slot of a wider tensor."""
    from dcfa_b200 import abi, pack
    g = torch.Generator().manual_seed(31 + c)
    gi = n // groups
    xfull = bf16_round(torch.randn(n, h, w, c + 16, generator=g))          # the chain reads channels [16, 16 + c)
    w1 = [bf16_round(torch.randn(c, c, 1, 1, generator=g) / c ** 0.5) for _ in range(groups)]
    w2 = [bf16_round(torch.randn(c, c, 1, 1, generator=g) / c ** 0.5) for _ in range(groups)]
    wd = [torch.randn(c, 1, 3, 3, generator=g) * 0.3 for _ in range(groups)]
    bd = [torch.randn(c, generator=g) * 0.2 for _ in range(groups)]
    s1 = [torch.rand(c, generator=g) + 0.5 for _ in range(groups)]
    b1 = [torch.randn(c, generator=g) * 0.2 for _ in range(groups)]
    s2 = [torch.rand(c, generator=g) + 0.5 for _ in range(groups)]
    b2 = [torch.randn(c, generator=g) * 0.2 for _ in range(groups)]
    p1 = [pack.pack_conv_weight(t) for t in w1]
    p2 = [pack.pack_conv_weight(t) for t in w2]
    m1, m2 = p1[0][1], p2[0][1]
    xg = xfull.to(torch.bfloat16).to(cuda)
    W1, W2 = torch.stack([t[0] for t in p1]).to(cuda), torch.stack([t[0] for t in p2]).to(cuda)
    S1, B1 = torch.stack(s1).to(cuda), torch.stack(b1).to(cuda)
    S2, B2 = torch.stack(s2).to(cuda), torch.stack(b2).to(cuda)
    WD = torch.stack([t.reshape(c, 9).t().contiguous() for t in wd]).to(cuda)
    BD = torch.stack(bd).to(cuda)
    outs = []
    for fused in ("2", "0"):   # "2": fused even for the shapes the launch heuristic leaves to the three kernels
        monkeypatch.setenv("DCFA_CHAIN", fused)
        t1 = torch.zeros(n, h, w, c, dtype=torch.bfloat16, device=cuda)
        t2 = torch.zeros(n, h, w, c, dtype=torch.bfloat16, device=cuda)
        y = torch.full((n, h, w, 2 * c), 3.0, dtype=torch.bfloat16, device=cuda)   # output slot: channels [c, 2c)
        bufs = [xg, W1, S1, B1, t1, WD, BD, t2, W2, S2, B2, y]
        common = dict(n_img=n, group_imgs=gi, Hi=h, Wi=w, Ho=h, Wo=w, Cin=c, Cout=c)
        ops = [
            abi.new_op(abi.OP_CONV, act=abi.ACT_RELU, x=nhwc_view(xg, 0, 16), w=flat_view(1), scale=flat_view(2), bias=flat_view(3),
                       y=nhwc_view(t1, 4), ksize=1, stride=1, BN=m1["BN"], n_tiles=m1["n_tiles"], k_blocks=m1["k_blocks"],
                       K_real=m1["K_real"], w_gstride=p1[0][0].numel(), sb_gstride=c, flags=m1["bk"] | abi.CONV_FLAG_CHAIN_HEAD,
                       **common),
            abi.new_op(abi.OP_DWCONV, act=abi.ACT_NONE, x=nhwc_view(t1, 4), w=flat_view(5), bias=flat_view(6), y=nhwc_view(t2, 7),
                       n_img=n, group_imgs=gi, Hi=h, Wi=w, Cin=c),
            abi.new_op(abi.OP_CONV, act=abi.ACT_RELU, x=nhwc_view(t2, 7), w=flat_view(8), scale=flat_view(9), bias=flat_view(10),
                       y=nhwc_view(y, 11, c), ksize=1, stride=1, BN=m2["BN"], n_tiles=m2["n_tiles"], k_blocks=m2["k_blocks"],
                       K_real=m2["K_real"], w_gstride=p2[0][0].numel(), sb_gstride=c, flags=m2["bk"], **common),
        ]
        _run(ops, bufs)
        if fused == "2":
            assert float(t1.float().abs().max()) == 0.0 and float(t2.float().abs().max()) == 0.0, "fused path not taken"
        assert (y[..., :c] == 3.0).all(), "chain wrote outside its channel slot"
        outs.append(y[..., c:].float().cpu())
    assert torch.equal(outs[0], outs[1]), "fused chain differs from the three-kernel path: max %g" % (outs[0] - outs[1]).abs().max()
    x = xfull[..., 16:16 + c].permute(0, 3, 1, 2)
    refs = []
    for gg in range(groups):
        t = F.relu(F.conv2d(x[gg * gi:(gg + 1) * gi], w1[gg]) * s1[gg].view(1, -1, 1, 1) + b1[gg].view(1, -1, 1, 1))
        t = bf16_round(t)
        t = bf16_round(F.conv2d(t, wd[gg], bd[gg], 1, 1, groups=c))
        refs.append(F.relu(F.conv2d(t, w2[gg]) * s2[gg].view(1, -1, 1, 1) + b2[gg].view(1, -1, 1, 1)))
    _bf16_close(outs[0], torch.cat(refs).permute(0, 2, 3, 1), "chain")


@pytest.mark.parametrize("n,h,w,c,act,use_res", [(2, 16, 32, 64, 2, False), (3, 21, 37, 64, 0, True), (2, 40, 40, 128, 2, False),
                                                 (2, 40, 40, 128, 0, True), (1, 8, 16, 32, 2, True), (2, 80, 80, 64, 0, True)])
def test_repghost_module_fused_vs_two_kernels(cuda, monkeypatch, n, h, w, c, act, use_res):
    """RepGhostModule in deploy algebra (nets/repghost.py:98-123, :263-279): 1x1 conv + BN (+SiLU) -> depthwise 3x3 (+SiLU)
    (+ residual).  The fused kernel (DCFA_CONV_FLAG_GHOST_HEAD) against the two-kernel path and against torch; input,
    output and residual are channel slots of wider tensors, as in C2f_repghost's concat buffer."""
    from dcfa_b200 import abi, pack
    g = torch.Generator().manual_seed(41 + c + h)
    cat = bf16_round(torch.randn(n, h, w, 3 * c, generator=g))       # [other | module input / residual | output slot]
    w1 = bf16_round(torch.randn(c, c, 1, 1, generator=g) / c ** 0.5)
    wd = torch.randn(c, 1, 3, 3, generator=g) * 0.3
    bd = torch.randn(c, generator=g) * 0.2
    s1 = torch.rand(c, generator=g) + 0.5
    b1 = torch.randn(c, generator=g) * 0.2
    pk, m1 = pack.pack_conv_weight(w1)
    W1, S1, B1 = pk.to(cuda), s1.to(cuda), b1.to(cuda)
    WD, BD = wd.reshape(c, 9).t().contiguous().to(cuda), bd.to(cuda)
    outs = []
    for fused in ("2", "0"):
        monkeypatch.setenv("DCFA_GHOST", fused)
        xg = cat.to(torch.bfloat16).to(cuda)
        t1 = torch.zeros(n, h, w, c, dtype=torch.bfloat16, device=cuda)
        bufs = [xg, W1, S1, B1, t1, WD, BD]
        common = dict(n_img=n, group_imgs=n, Hi=h, Wi=w, Ho=h, Wo=w, Cin=c, Cout=c)
        dw = abi.new_op(abi.OP_DWCONV, act=act, x=nhwc_view(t1, 4), w=flat_view(5), bias=flat_view(6), y=nhwc_view(xg, 0, 2 * c),
                        n_img=n, group_imgs=n, Hi=h, Wi=w, Cin=c)
        if use_res:
            dw.x2 = nhwc_view(xg, 0, c)
        ops = [abi.new_op(abi.OP_CONV, act=abi.ACT_SILU if act else abi.ACT_NONE, x=nhwc_view(xg, 0, c), w=flat_view(1),
                          scale=flat_view(2), bias=flat_view(3), y=nhwc_view(t1, 4), ksize=1, stride=1, BN=m1["BN"],
                          n_tiles=m1["n_tiles"], k_blocks=m1["k_blocks"], K_real=m1["K_real"], w_gstride=pk.numel(), sb_gstride=c,
                          flags=m1["bk"] | abi.CONV_FLAG_GHOST_HEAD, **common), dw]
        _run(ops, bufs)
        if fused == "2":
            assert float(t1.float().abs().max()) == 0.0, "fused path not taken"
        assert torch.equal(xg[..., :2 * c].float().cpu(), cat[..., :2 * c]), "module wrote outside its output slot"
        outs.append(xg[..., 2 * c:].float().cpu())
    _bf16_close(outs[0], outs[1], "ghost fused vs two kernels")
    x = cat[..., c:2 * c].permute(0, 3, 1, 2)
    t = F.conv2d(x, w1) * s1.view(1, -1, 1, 1) + b1.view(1, -1, 1, 1)
    t = bf16_round(F.silu(t) if act else t)
    t = F.conv2d(t, wd, bd, 1, 1, groups=c)
    t = F.silu(t) if act else t
    if use_res:
        t = t + x
    _bf16_close(outs[0], t.permute(0, 2, 3, 1), "ghost module")


def _cbam_ref(x, fc1, fc2, w7):
    """CBAM restated with plain torch ops (nets/yolo_mul.py:56-102); x NCHW fp32."""
    avg, mx = x.mean((2, 3), keepdim=True), x.amax((2, 3), keepdim=True)
    mlp = lambda v: F.conv2d(F.relu(F.conv2d(v, fc1)), fc2)
    t = x * torch.sigmoid(mlp(avg) + mlp(mx))
    p = torch.cat([t.mean(1, keepdim=True), t.amax(1, keepdim=True)], 1)
    return t * torch.sigmoid(F.conv2d(p, w7, padding=3))


@pytest.mark.parametrize("fused", ["0", "2"])
@pytest.mark.parametrize("b,h,w,c,hidden,groups", [(2, 20, 20, 128, 1, 2), (1, 80, 80, 64, 8, 2), (2, 12, 28, 256, 32, 1),
                                                   (3, 20, 20, 512, 32, 2), (1, 7, 9, 32, 2, 1), (24, 40, 40, 128, 8, 2)])
def test_cbam_chain(cuda, monkeypatch, b, h, w, c, hidden, groups, fused):
    """pool -> mlp -> stats -> apply, two modalities writing adjacent channel slots of one concat buffer; as four
    kernels ("0") and as the one-cluster-per-image fused kernel ("2": forced even where the heuristic would not pick it;
    cluster sizes 8, 4, 2 and 1 occur across the cases)."""
    from dcfa_b200 import abi
    monkeypatch.setenv("DCFA_CBAM_FUSED", fused)
    g = torch.Generator().manual_seed(21)
    n = b * groups
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    fc1 = [torch.randn(hidden, c, 1, 1, generator=g) * 0.2 for _ in range(groups)]
    fc2 = [torch.randn(c, hidden, 1, 1, generator=g) * 0.5 for _ in range(groups)]
    w7 = [torch.randn(1, 2, 7, 7, generator=g) * 0.2 for _ in range(groups)]
    xg = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(cuda)
    parts = 4
    psum = torch.zeros(n, parts, c, device=cuda)
    pmax = torch.zeros(n, parts, c, device=cuda)
    gate = torch.zeros(n, c, device=cuda)
    stats = torch.zeros(n, h, w, 2, device=cuda)
    f1 = torch.stack([t.reshape(hidden, c) for t in fc1]).to(cuda)
    f2 = torch.stack([t.reshape(c, hidden) for t in fc2]).to(cuda)
    k7 = torch.stack([t.reshape(98) for t in w7]).to(cuda)
    slot0 = 16
    ctot = slot0 + groups * c
    y = torch.full((b, h, w, ctot), 5.0, dtype=torch.bfloat16, device=cuda)
    bufs = [xg, psum, pmax, gate, stats, f1, f2, k7, y]
    common = dict(n_img=n, group_imgs=b, Hi=h, Wi=w, Cin=c, hidden=hidden, parts=parts)
    ops = [
        abi.new_op(abi.OP_CBAM_POOL, x=nhwc_view(xg, 0), a0=flat_view(1), a1=flat_view(2), **common),
        abi.new_op(abi.OP_CBAM_MLP, a0=flat_view(1), a1=flat_view(2), a2=flat_view(3), w=flat_view(5), scale=flat_view(6),
                   w_gstride=hidden * c, sb_gstride=hidden * c, **common),
        abi.new_op(abi.OP_CBAM_STATS, x=nhwc_view(xg, 0), a2=flat_view(3), a0=flat_view(4), **common),
        abi.new_op(abi.OP_CBAM_APPLY, x=nhwc_view(xg, 0), a2=flat_view(3), a0=flat_view(4), w=flat_view(7),
                   y=nhwc_view(y, 8, slot0, gi=b, gstride=c), **common),
    ]
    _run(ops, bufs)
    for gg in range(groups):
        ref = _cbam_ref(x[gg * b:(gg + 1) * b], fc1[gg], fc2[gg], w7[gg]).permute(0, 2, 3, 1)
        _bf16_close(y[..., slot0 + gg * c: slot0 + (gg + 1) * c].cpu(), ref, "cbam group %d" % gg)
    assert (y[..., :slot0] == 5.0).all()


@pytest.mark.parametrize("b,h,w,c,hidden,groups", [(2, 20, 20, 256, 16, 2), (1, 12, 12, 64, 4, 1), (1, 40, 40, 256, 16, 2),
                                                   (3, 20, 28, 128, 1, 1)])
def test_sppf_cbam_sequence_fused_vs_separate(cuda, monkeypatch, b, h, w, c, hidden, groups):
    """SPPF_CBAM after cv1 (nets/yolo_mul.py:18-31): x1 = cbam1(t), x_{k+1} = cbam_{k+1}(maxpool5(x_k)), the four results in the
    slots of one concat buffer.  The 19 records run (a) as ONE cluster kernel that keeps each image resident in shared memory
    ("2": forced for these small batches; cluster sizes 2, 1, 8 and 4 occur) and (b) unit by unit ("0"); both against torch."""
    from dcfa_b200 import abi
    g = torch.Generator().manual_seed(77 + c)
    n = b * groups
    t = bf16_round(torch.randn(n, c, h, w, generator=g))
    fc1 = [[torch.randn(hidden, c, 1, 1, generator=g) * 0.2 for _ in range(groups)] for _ in range(4)]
    fc2 = [[torch.randn(c, hidden, 1, 1, generator=g) * 0.5 for _ in range(groups)] for _ in range(4)]
    w7 = [[torch.randn(1, 2, 7, 7, generator=g) * 0.2 for _ in range(groups)] for _ in range(4)]
    tg = t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(cuda)
    parts = 2
    outs = []
    for fused in ("2", "0"):
        monkeypatch.setenv("DCFA_SPPF_FUSED", fused)
        cat = torch.full((n, h, w, 4 * c + 8), 7.0, dtype=torch.bfloat16, device=cuda)     # slots start at channel 8
        pooled = [torch.zeros(n, h, w, c, dtype=torch.bfloat16, device=cuda) for _ in range(3)]
        psum, pmax = torch.zeros(n, parts, c, device=cuda), torch.zeros(n, parts, c, device=cuda)
        gate, stats = torch.zeros(n, c, device=cuda), torch.zeros(n, h, w, 2, device=cuda)
        bufs = [tg, cat, psum, pmax, gate, stats] + pooled
        ops = []
        common = dict(n_img=n, group_imgs=b, Hi=h, Wi=w, Cin=c, hidden=hidden, parts=parts)
        for s in range(4):
            f1 = torch.stack([q.reshape(hidden, c) for q in fc1[s]]).to(cuda)
            f2 = torch.stack([q.reshape(c, hidden) for q in fc2[s]]).to(cuda)
            k7 = torch.stack([q.reshape(98) for q in w7[s]]).to(cuda)
            i0 = len(bufs)
            bufs += [f1, f2, k7]
            if s > 0:
                ops.append(abi.new_op(abi.OP_MAXPOOL5, x=nhwc_view(cat, 1, 8 + (s - 1) * c), y=nhwc_view(pooled[s - 1], 5 + s),
                                      n_img=n, Hi=h, Wi=w, Cin=c, Ho=h, Wo=w, Cout=c))
            xin = nhwc_view(tg, 0) if s == 0 else nhwc_view(pooled[s - 1], 5 + s)
            ops += [
                abi.new_op(abi.OP_CBAM_POOL, x=xin, a0=flat_view(2), a1=flat_view(3), **common),
                abi.new_op(abi.OP_CBAM_MLP, a0=flat_view(2), a1=flat_view(3), a2=flat_view(4), w=flat_view(i0), scale=flat_view(i0 + 1),
                           w_gstride=hidden * c, sb_gstride=hidden * c, **common),
                abi.new_op(abi.OP_CBAM_STATS, x=xin, a2=flat_view(4), a0=flat_view(5), **common),
                abi.new_op(abi.OP_CBAM_APPLY, x=xin, a2=flat_view(4), a0=flat_view(5), w=flat_view(i0 + 2),
                           y=nhwc_view(cat, 1, 8 + s * c), **common),
            ]
        assert len(ops) == 19
        n0 = _launches()
        _run(ops, bufs)
        launched = _launches() - n0
        assert fused == "0" or launched == 1, "fused SPPF_CBAM took %d launches" % launched
        if fused == "2":
            assert all(float(q.float().abs().max()) == 0.0 for q in pooled), "fused path not taken"
        assert (cat[..., :8] == 7.0).all()
        outs.append(cat[..., 8:].float().cpu())
    _bf16_close(outs[0], outs[1], "fused SPPF_CBAM vs separate units")
    x = t
    for s in range(4):
        if s > 0:
            x = F.max_pool2d(x, 5, 1, 2)
        x = bf16_round(torch.cat([_cbam_ref(x[gg * b:(gg + 1) * b], fc1[s][gg], fc2[s][gg], w7[s][gg]) for gg in range(groups)]))
        _bf16_close(outs[0][..., s * c:(s + 1) * c], x.permute(0, 2, 3, 1), "SPPF_CBAM stage %d" % s)


def test_maxpool5_and_upsample(cuda):
    from dcfa_b200 import abi
    g = torch.Generator().manual_seed(3)
    n, h, w, c = 3, 20, 20, 64
    x = bf16_round(torch.randn(n, c, h, w, generator=g))
    x2 = bf16_round(torch.randn(n, c, h, w, generator=g))
    xg = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(cuda)
    x2g = x2.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(cuda)
    y = torch.zeros(n, h, w, c, dtype=torch.bfloat16, device=cuda)
    up = torch.zeros(n, 40, 40, c + 8, dtype=torch.bfloat16, device=cuda)
    ops = [abi.new_op(abi.OP_MAXPOOL5, x=nhwc_view(xg, 0), y=nhwc_view(y, 2), n_img=n, Hi=h, Wi=w, Cin=c),
           abi.new_op(abi.OP_UPSAMPLE, x=nhwc_view(xg, 0), x2=nhwc_view(x2g, 1), y=nhwc_view(up, 3, 8), n_img=n, Hi=h, Wi=w,
                      Ho=40, Wo=40, Cin=c)]
    _run(ops, [xg, x2g, y, up])
    assert torch.equal(y.float().cpu(), F.max_pool2d(x, 5, 1, 2).permute(0, 2, 3, 1))
    ref = F.interpolate(x + x2, size=(40, 40), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    _bf16_close(up[..., 8:].cpu(), ref, "upsample")


def _dfl_ref(maps, nc):
    b = maps[0].shape[0]
    cat = torch.cat([m.reshape(b, 64 + nc, -1) for m in maps], 2)
    box, cls = cat.split((64, nc), 1)
    a = box.shape[-1]
    d = (box.view(b, 4, 16, a).transpose(2, 1).softmax(1) * torch.arange(16.0).view(1, 16, 1, 1)).sum(1)
    return d, cls


@pytest.mark.parametrize("nc", [1, 3])
def test_dfl_and_decode(cuda, nc):
    """DFL (+ level gather) and decode_box against torch restatements; adversarial logits included."""
    import ctypes as C
    from dcfa_b200 import _lib, abi
    g = torch.Generator().manual_seed(17)
    b, h0, w0 = 2, 12, 20
    sizes = [(h0, w0), (h0 // 2, w0 // 2), (h0 // 4, w0 // 4)]
    maps = [torch.randn(b, 64 + nc, hh, ww, generator=g) * 3 for hh, ww in sizes]
    maps[0][0, :16, 0, 0] = 30.0          # uniform large
    maps[0][0, 16:32, 0, 0] = torch.tensor([-30.0] * 15 + [30.0])  # one-hot at the last bin
    maps[1][1, 32:48, 1, 1] = 0.0
    a_tot = sum(hh * ww for hh, ww in sizes)
    mg = [m.to(cuda) for m in maps]
    dbox = torch.zeros(b, 4, a_tot, device=cuda)
    cls = torch.zeros(b, nc, a_tot, device=cuda)
    op = abi.new_op(abi.OP_DFL, a0=flat_view(0), a1=flat_view(1), a2=flat_view(2), y=flat_view(3), x2=flat_view(4), n_img=b,
                    Hi=h0, Wi=w0, nc=nc, A=a_tot)
    _run([op], mg + [dbox, cls])
    d_ref, c_ref = _dfl_ref(maps, nc)
    assert torch.allclose(dbox.cpu(), d_ref, atol=1e-5, rtol=0), (dbox.cpu() - d_ref).abs().max()
    assert torch.equal(cls.cpu(), c_ref)

    # decode_box (utils/utils_bbox.py:49-58)
    anchors, strides = [], []
    for (hh, ww), s in zip(sizes, (8.0, 16.0, 32.0)):
        sy, sx = torch.meshgrid(torch.arange(hh) + 0.5, torch.arange(ww) + 0.5, indexing="ij")
        anchors.append(torch.stack((sx, sy), -1).view(-1, 2))
        strides.append(torch.full((hh * ww, 1), s))
    anc = torch.cat(anchors).t().contiguous()   # (2, A)
    strd = torch.cat(strides).t().contiguous()  # (1, A)
    img_w, img_h = w0 * 8.0, h0 * 8.0
    out = torch.zeros(b, a_tot, 4 + nc, device=cuda)
    ancg, strg = anc.to(cuda), strd.to(cuda)
    _lib.check(_lib.lib.dcfa_decode_box(dbox.data_ptr(), cls.data_ptr(), nc * a_tot, ancg.data_ptr(), a_tot, 1,
                                        strg.data_ptr(), b, a_tot, nc, img_w, img_h, out.data_ptr(),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    d = dbox.cpu()
    lt, rb = d.split(2, 1)
    x1y1, x2y2 = anc[None] - lt, anc[None] + rb
    ref = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strd
    ref = torch.cat((ref, cls.cpu().sigmoid()), 1).permute(0, 2, 1).clone()
    ref[:, :, :4] = ref[:, :, :4] / torch.tensor([img_w, img_h, img_w, img_h])
    assert torch.allclose(out.cpu(), ref, atol=1e-6, rtol=1e-6), (out.cpu() - ref).abs().max()
