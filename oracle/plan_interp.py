"""TEST INFRASTRUCTURE -- a CPU interpreter for the dcfa_op lists the plan compiler emits.

It executes the exact byte-level plan (same parameter blob, same arena offsets, same channel views) with plain
torch CPU ops, rounding activations to bf16 wherever the CUDA kernels store bf16.  It exists so that the host
logic (BN folding, channel permutations, concat slots, BiFPN folding, weight swizzling) can be checked against
the fp32 oracle WITHOUT a GPU, and so that each CUDA kernel can be checked against a second, independent
statement of its contract.  It is never imported by the product path (dcfa-yolo_b200/).
"""
import torch
import torch.nn.functional as F

BK = 64


def _bf16(buf, off_bytes):
    return buf.view(torch.bfloat16)[off_bytes // 2:]


def _f32(buf, off_bytes, numel):
    return buf.view(torch.float32)[off_bytes // 4: off_bytes // 4 + numel]


def _groups(v, n):
    gi = v.gi if v.gi > 0 else n
    return gi, n // gi


def read_nhwc(bufs, v, n, h, w, c):
    """-> float32 [n,h,w,c] gathered through a dcfa_view."""
    base = _bf16(bufs[v.buf], 0)
    gi, ng = _groups(v, n)
    outs = []
    for g in range(ng):
        off = v.off // 2 + g * v.gstride
        outs.append(torch.as_strided(base, (gi, h, w, c), (v.img_stride, w * v.ld, v.ld, 1), off).float())
    return torch.cat(outs)


def write_nhwc(bufs, v, t):
    n, h, w, c = t.shape
    base = _bf16(bufs[v.buf], 0)
    gi, ng = _groups(v, n)
    tb = t.to(torch.bfloat16)
    for g in range(ng):
        off = v.off // 2 + g * v.gstride
        torch.as_strided(base, (gi, h, w, c), (v.img_stride, w * v.ld, v.ld, 1), off).copy_(tb[g * gi:(g + 1) * gi])


def unswizzle_weights(blob, off, groups, w_gstride, bn, n_tiles, k_blocks, bk=0):
    """Inverse of pack.pack_conv_weight -> float32 [groups, n_tiles*bn, k_blocks*tile_k] (tile_k = bk or 64)."""
    tk = bk if bk else BK
    nch = tk // 8
    flat = blob.view(torch.bfloat16)[off // 2: off // 2 + groups * w_gstride].float()
    out = torch.zeros(groups, n_tiles * bn, k_blocks * tk)
    idx = (torch.arange(nch)[None, :] ^ ((torch.arange(bn) * nch // 8) % nch)[:, None])
    for g in range(groups):
        tiles = flat[g * w_gstride: g * w_gstride + n_tiles * k_blocks * bn * tk].view(n_tiles, k_blocks, bn, nch, 8)
        for nt in range(n_tiles):
            for kb in range(k_blocks):
                t = tiles[nt, kb]
                un = torch.gather(t, 1, idx[:, :, None].expand(bn, nch, 8)).reshape(bn, tk)
                out[g, nt * bn:(nt + 1) * bn, kb * tk:(kb + 1) * tk] = un
    return out


def _act(v, act):
    if act == 1:
        return F.relu(v)
    if act == 2:
        return v * torch.sigmoid(v)
    return v


def run_ops(ops, bufs):
    """ops: iterable of abi.Op; bufs: list of CPU uint8 tensors (None for unused slots), modified in place."""
    for op in ops:
        k = op.kind
        n = op.n_img
        G = n // op.group_imgs if op.group_imgs > 0 else 1
        gi = n // G
        if k == 1:  # STEM: bf16 inputs and weights, fp32 accumulate, fp32 scale/bias, ReLU, pool, bf16 store
            c0 = op.Cout
            # weight tiles [G][nblk][3][128 x 16] in the canonical no-swizzle K-major layout (pack.pack_stem); sign-folded
            from dcfa_b200 import pack as _pack
            wraw = bufs[op.w.buf][op.w.off:op.w.off + 2 * G * op.w_gstride].view(torch.bfloat16).view(G, -1)
            sc = _f32(bufs[op.scale.buf], op.scale.off, G * op.sb_gstride).view(G, -1)
            bi = _f32(bufs[op.bias.buf], op.bias.off, G * op.sb_gstride).view(G, -1)
            outs = []
            for g in range(G):
                src = bufs[op.x.buf] if g == 0 else bufs[op.x2.buf]
                if (op.flags & 0x200) and g == 1:   # single uint8 plane, replicated to three channels by the kernel
                    x = src.view(gi, 1, op.Hi, op.Wi).expand(gi, 3, op.Hi, op.Wi).float()
                elif op.flags & 0x100:   # uint8 NHWC pixels, scale already / 255
                    x = src.view(gi, op.Hi, op.Wi, 3).permute(0, 3, 1, 2).float()
                else:
                    x = src.view(torch.float32).view(gi, 3, op.Hi, op.Wi).to(torch.bfloat16).float()
                wt = _pack.unpack_stem(wraw[g], c0)
                y = F.conv2d(x, wt, None, 1, 1) * sc[g, :c0].view(1, -1, 1, 1) + bi[g, :c0].view(1, -1, 1, 1)
                y = F.max_pool2d(F.relu(y), 3, 2, 1)
                outs.append(y.permute(0, 2, 3, 1))
            write_nhwc(bufs, op.y, torch.cat(outs))
        elif k == 2:  # CONV
            x = read_nhwc(bufs, op.x, n, op.Hi, op.Wi, op.Cin).permute(0, 3, 1, 2)
            wall = unswizzle_weights(bufs[op.w.buf], op.w.off, G, op.w_gstride, op.BN, op.n_tiles, op.k_blocks, op.flags & 0xff)
            npad = op.BN * op.n_tiles
            sc = _f32(bufs[op.scale.buf], op.scale.off, G * op.sb_gstride).view(G, -1)[:, :npad]
            bi = _f32(bufs[op.bias.buf], op.bias.off, G * op.sb_gstride).view(G, -1)[:, :npad]
            outs = []
            for g in range(G):
                if op.flags & 0x800:   # pixel-pair k-blocks: [0 | w(dy,0)], [w(dy,1) | w(dy,2)] per kernel row
                    wp = wall[g, :op.Cout]
                    wt = torch.zeros(op.Cout, op.Cin, 3, 3)
                    for dy in range(3):
                        assert float(wp[:, (2 * dy) * 64:(2 * dy) * 64 + 32].abs().max()) == 0.0
                        wt[:, :, dy, 0] = wp[:, (2 * dy) * 64 + 32:(2 * dy) * 64 + 64]
                        wt[:, :, dy, 1] = wp[:, (2 * dy + 1) * 64:(2 * dy + 1) * 64 + 32]
                        wt[:, :, dy, 2] = wp[:, (2 * dy + 1) * 64 + 32:(2 * dy + 1) * 64 + 64]
                else:
                    wt = wall[g, :op.Cout, :op.K_real].view(op.Cout, op.ksize, op.ksize, op.Cin).permute(0, 3, 1, 2)
                y = F.conv2d(x[g * gi:(g + 1) * gi], wt.contiguous(), None, op.stride, op.ksize // 2)
                y = y * sc[g, :op.Cout].view(1, -1, 1, 1) + bi[g, :op.Cout].view(1, -1, 1, 1)
                outs.append(_act(y, op.act) * op.f0)
            y = torch.cat(outs)
            if op.out_mode == 0:
                y = y.permute(0, 2, 3, 1)
                if op.x2.buf >= 0:
                    y = y + read_nhwc(bufs, op.x2, n, op.Ho, op.Wo, op.Cout)
                if op.parts > 0:   # split output: the channels from `parts` on go to the second view
                    write_nhwc(bufs, op.y, y[..., :op.parts])
                    write_nhwc(bufs, op.a0, y[..., op.parts:])
                else:
                    write_nhwc(bufs, op.y, y)
            else:
                dst = bufs[op.y.buf].view(torch.float32).view(n, op.out_ctot, op.Ho, op.Wo)
                dst[:, op.out_coff:op.out_coff + op.Cout] = y
                if op.flags & 0x2000:   # DCFA_CONV_FLAG_DFL: DFL of the box channels + class gather for this level's anchors
                    hw = op.Ho * op.Wo
                    box = y[:, :64].reshape(n, 4, 16, hw)
                    dbox = (torch.softmax(box, 2) * torch.arange(16.0).view(1, 1, 16, 1)).sum(2)
                    bufs[op.a1.buf].view(torch.float32).view(n, 4, op.A)[:, :, op.hidden:op.hidden + hw] = dbox
                    bufs[op.a2.buf].view(torch.float32).view(n, op.nc, op.A)[:, :, op.hidden:op.hidden + hw] = \
                        y[:, 64:64 + op.nc].reshape(n, op.nc, hw)
        elif k == 3:  # DWCONV
            c = op.Cin
            x = read_nhwc(bufs, op.x, n, op.Hi, op.Wi, c).permute(0, 3, 1, 2)
            w = _f32(bufs[op.w.buf], op.w.off, G * 9 * c).view(G, 9, c)
            b = _f32(bufs[op.bias.buf], op.bias.off, G * c).view(G, c)
            outs = []
            for g in range(G):
                wt = w[g].t().reshape(c, 1, 3, 3).contiguous()
                outs.append(_act(F.conv2d(x[g * gi:(g + 1) * gi], wt, b[g], 1, 1, groups=c), op.act))
            y = torch.cat(outs).permute(0, 2, 3, 1)
            if op.x2.buf >= 0:
                y = y + read_nhwc(bufs, op.x2, n, op.Hi, op.Wi, c)
            write_nhwc(bufs, op.y, y)
        elif k == 4:  # CBAM_POOL -> partials
            c, hw, parts = op.Cin, op.Hi * op.Wi, op.parts
            x = read_nhwc(bufs, op.x, n, op.Hi, op.Wi, c).reshape(n, hw, c)
            ppp = -(-hw // parts)
            ps = _f32(bufs[op.a0.buf], op.a0.off, n * parts * c).view(n, parts, c)
            pm = _f32(bufs[op.a1.buf], op.a1.off, n * parts * c).view(n, parts, c)
            for q in range(parts):
                seg = x[:, q * ppp:min(hw, (q + 1) * ppp)]
                if seg.shape[1] == 0:
                    ps[:, q] = 0
                    pm[:, q] = float('-inf')
                else:
                    ps[:, q] = seg.sum(1)
                    pm[:, q] = seg.amax(1)
        elif k == 5:  # CBAM_MLP
            c, hid, parts = op.Cin, op.hidden, op.parts
            ps = _f32(bufs[op.a0.buf], op.a0.off, n * parts * c).view(n, parts, c)
            pm = _f32(bufs[op.a1.buf], op.a1.off, n * parts * c).view(n, parts, c)
            avg, mx = ps.sum(1) / (op.Hi * op.Wi), pm.amax(1)
            fc1 = _f32(bufs[op.w.buf], op.w.off, G * hid * c).view(G, hid, c)
            fc2 = _f32(bufs[op.scale.buf], op.scale.off, G * c * hid).view(G, c, hid)
            gate = _f32(bufs[op.a2.buf], op.a2.off, n * c).view(n, c)
            for g in range(G):
                sl = slice(g * gi, (g + 1) * gi)
                hsum = F.relu(avg[sl] @ fc1[g].t()) + F.relu(mx[sl] @ fc1[g].t())
                gate[sl] = torch.sigmoid(hsum @ fc2[g].t())
        elif k == 6:  # CBAM_STATS
            c = op.Cin
            x = read_nhwc(bufs, op.x, n, op.Hi, op.Wi, c)
            gate = _f32(bufs[op.a2.buf], op.a2.off, n * c).view(n, 1, 1, c)
            t = x * gate
            st = _f32(bufs[op.a0.buf], op.a0.off, n * op.Hi * op.Wi * 2).view(n, op.Hi, op.Wi, 2)
            st[..., 0] = t.mean(-1)
            st[..., 1] = t.amax(-1)
        elif k == 7:  # CBAM_APPLY
            c = op.Cin
            x = read_nhwc(bufs, op.x, n, op.Hi, op.Wi, c)
            gate = _f32(bufs[op.a2.buf], op.a2.off, n * c).view(n, 1, 1, c)
            st = _f32(bufs[op.a0.buf], op.a0.off, n * op.Hi * op.Wi * 2).view(n, op.Hi, op.Wi, 2).permute(0, 3, 1, 2)
            w7 = _f32(bufs[op.w.buf], op.w.off, G * 98).view(G, 1, 2, 7, 7)
            outs = []
            for g in range(G):
                sl = slice(g * gi, (g + 1) * gi)
                s = torch.sigmoid(F.conv2d(st[sl], w7[g], padding=3)).permute(0, 2, 3, 1)
                outs.append(x[sl] * (gate[sl] * s))
            write_nhwc(bufs, op.y, torch.cat(outs))
        elif k == 8:  # MAXPOOL5
            x = read_nhwc(bufs, op.x, n, op.Hi, op.Wi, op.Cin).permute(0, 3, 1, 2)
            write_nhwc(bufs, op.y, F.max_pool2d(x, 5, 1, 2).permute(0, 2, 3, 1))
        elif k == 9:  # UPSAMPLE
            x = read_nhwc(bufs, op.x, n, op.Hi, op.Wi, op.Cin)
            if op.x2.buf >= 0:
                x = x + read_nhwc(bufs, op.x2, n, op.Hi, op.Wi, op.Cin)
            y = F.interpolate(x.permute(0, 3, 1, 2), size=(op.Ho, op.Wo), mode='bilinear', align_corners=True)
            write_nhwc(bufs, op.y, y.permute(0, 2, 3, 1))
        elif k == 10:  # DFL
            no = 64 + op.nc
            hh, ww, maps = op.Hi, op.Wi, []
            for v in (op.a0, op.a1, op.a2):
                maps.append(bufs[v.buf].view(torch.float32)[v.off // 4: v.off // 4 + n * no * hh * ww].view(n, no, hh * ww))
                hh, ww = (hh + 1) // 2, (ww + 1) // 2
            cat = torch.cat(maps, 2)
            box, cls = cat[:, :64], cat[:, 64:]
            a = cat.shape[-1]
            d = (box.reshape(n, 4, 16, a).softmax(2) * torch.arange(16.0).view(1, 1, 16, 1)).sum(2)
            bufs[op.y.buf].view(torch.float32)[:n * 4 * a].view(n, 4, a).copy_(d)
            bufs[op.x2.buf].view(torch.float32)[:n * op.nc * a].view(n, op.nc, a).copy_(cls)
        else:
            raise ValueError("plan_interp: unknown op kind %d" % k)


def run_plan(plan, rgb, nir):
    """Execute a dcfa_b200.plan.Plan on CPU.  rgb, nir: float32 [B,3,H,W] (uint8 [B,H,W,3] for an input_u8 plan).
    -> (dbox, cls, [x0,x1,x2])."""
    from dcfa_b200 import plan as P
    b, no = plan.B, plan.no
    f32 = lambda numel: torch.zeros(numel * 4, dtype=torch.uint8)
    bufs = [None] * P.NUM_BUFS
    bufs[P.BUF_BLOB] = plan.blob_tensor.clone()
    bufs[P.BUF_ARENA] = torch.zeros(plan.arena_bytes + 256, dtype=torch.uint8)
    bufs[P.BUF_RGB] = rgb.contiguous().view(torch.uint8).reshape(-1).clone()
    bufs[P.BUF_NIR] = nir.contiguous().view(torch.uint8).reshape(-1).clone()
    for i, (h, w) in enumerate(plan.level_shapes):
        bufs[P.BUF_X0 + i] = f32(b * no * h * w)
    bufs[P.BUF_DBOX] = f32(b * 4 * plan.A)
    bufs[P.BUF_CLS] = f32(b * plan.nc * plan.A)
    run_ops(plan.ops, bufs)
    x = [bufs[P.BUF_X0 + i].view(torch.float32).view(b, no, h, w) for i, (h, w) in enumerate(plan.level_shapes)]
    return (bufs[P.BUF_DBOX].view(torch.float32).view(b, 4, plan.A), bufs[P.BUF_CLS].view(torch.float32).view(b, plan.nc, plan.A), x)
