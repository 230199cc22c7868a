"""Plan compiler: turns a DCFA-YOLO state_dict into (a) one packed parameter blob and (b) a flat list of
dcfa_op records for lib/libdcfa_b200.so, for a fixed (batch, H, W).

What it folds away at pack time (all are exact restatements of eval-mode arithmetic):
  * BatchNorm -> per-channel scale/bias in the conv epilogue (reference nets/yolo_mul.py:197, :109; eps per site);
  * RepGhost identity-BN branch -> centre tap of the depthwise kernel (nets/repghost.py:117-123, the
    reference's own switch_to_deploy algebra);
  * ShuffleNetV2 channel_shuffle (nets/yolo_mul.py:164-168) -> never executed: tensors stay in physical
    order [x1 | branch2] and every consumer's weights are permuted along their input channels;
  * torch.cat / chunk / split (nets/yolo_mul.py:32,:156,:439; nets/repghost.py:318-320) -> producers write at
    channel offsets of pre-allocated NHWC buffers, consumers read channel sub-views;
  * Concat_BiFPN's normalised scalars (nets/yolo_mul.py:44-51) -> folded into the input-channel columns of
    the consuming 1x1 conv (conv is linear in its input);
  * feat3_rgb + feat3_nir (nets/yolo_mul.py:421) -> summed inside the bilinear upsample kernel, its only consumer;
  * the first convs of the head's box and class branches (nets/yolo_mul.py:389,:391) share their input ->
    one GEMM with N = c2 + c3.
Both modalities run as ONE launch per layer: activations are [2B, ...] (rgb images first), weights carry a
group dimension.

The op list only holds (buffer index, byte offset) references, so it is relocatable; buffer indices are BUF_*.
"""
import math

import torch

from . import abi, pack

BUF_BLOB, BUF_ARENA, BUF_RGB, BUF_NIR, BUF_X0, BUF_X1, BUF_X2, BUF_DBOX, BUF_CLS = range(9)
NUM_BUFS = 9

# nets/yolo_mul.py:331-337
_DEPTH = {'n': 0.33, 's': 0.33, 'm': 0.67, 'l': 1.00, 'x': 1.00}
_WIDTH = {'n': 0.25, 's': 0.50, 'm': 0.75, 'l': 1.00, 'x': 1.25}
_DEEP = {'n': 1.00, 's': 1.00, 'm': 0.75, 'l': 0.50, 'x': 0.50}


def model_dims(phi):
    bc = int(_WIDTH[phi] * 64)
    depth = max(round(_DEPTH[phi] * 3), 1)
    return bc, depth, bc * 4, bc * 8, int(bc * 16 * _DEEP[phi])


def down2(v):
    """Output size of a 3x3 stride-2 pad-1 conv / 3x3 stride-2 pad-1 max-pool."""
    return (v - 1) // 2 + 1


class TRef:
    """A channel sub-view of an NHWC bf16 tensor in the arena."""

    def __init__(self, off, n, h, w, ld, c_off, c, gi=0, gstride=0, perm=None):
        self.off, self.n, self.h, self.w, self.ld, self.c_off, self.c = off, n, h, w, ld, c_off, c
        self.gi, self.gstride = gi, gstride
        self.perm = perm  # logical channel -> physical channel (inside this view); None = identity

    def sub(self, c_off, c, perm=None):
        return TRef(self.off, self.n, self.h, self.w, self.ld, self.c_off + c_off, c, self.gi, self.gstride, perm)

    def view(self, buf=BUF_ARENA):
        return abi.View(buf, self.ld, self.off + 2 * self.c_off, self.h * self.w * self.ld, self.gstride, self.gi, 0)

    def perm_tensor(self):
        return torch.arange(self.c) if self.perm is None else torch.as_tensor(self.perm, dtype=torch.long)


def _flat(buf, off):
    return abi.View(buf, 0, off, 0, 0, 0, 0)


class Plan:
    def __init__(self, sd, phi, num_classes, batch, height, width, input_u8=False, depth_plane=False, reuse_arena=True):
        """input_u8: the two inputs are uint8 NHWC images [B,H,W,3] (pre-`preprocess_input`, utils/utils.py:76-79)
        instead of fp32 NCHW tensors in [0,1]; the stem folds the /255.  depth_plane (with input_u8): the second input
        is the single uint8 plane [B,H,W] that cvtColor would replicate to three channels (utils/utils.py:14-19)."""
        self.phi, self.nc, self.B, self.H, self.W = phi, int(num_classes), int(batch), int(height), int(width)
        self.input_u8 = bool(input_u8)
        self.depth_plane = bool(depth_plane)
        assert self.input_u8 or not self.depth_plane
        self.sd = {k: v.detach().float().cpu() for k, v in sd.items() if torch.is_floating_point(v)}
        self.blob = pack.Blob()
        self.ops = []
        self.op_names = []
        self.arena_bytes = 0
        self._allocs = []    # (bump offset, bytes) of every arena allocation, in program order
        self.conv_flops = 0  # 2*MAC over every convolution, per forward of this batch
        self._build()
        self.arena_bytes_bump = self.arena_bytes
        if reuse_arena:
            self._assign_arena_by_liveness()
        self.blob_tensor = self.blob.finish()
        self.op_array = (abi.Op * len(self.ops))(*self.ops)
        del self.sd

    def info(self):
        """dcfa_plan_info of this plan (what a plan file records)."""
        pi = abi.PlanInfo()
        pi.batch, pi.height, pi.width, pi.num_classes, pi.anchors, pi.no = self.B, self.H, self.W, self.nc, self.A, self.no
        for i, (h, w) in enumerate(self.level_shapes):
            pi.level_hw[i][0], pi.level_hw[i][1] = h, w
        pi.input_u8, pi.depth_plane = int(self.input_u8), int(self.depth_plane)
        return pi

    def save(self, path):
        """Write the plan file dcfa_plan_load reads (include/dcfa_b200.h: dcfa_plan_file_header, the op records, the
        parameter blob): everything a caller without Python needs to run this forward pass."""
        import ctypes as C
        h = abi.PlanFileHeader()
        h.magic = b"DCFAPLN1"
        h.abi_version, h.sizeof_op, h.n_ops, h.nbufs = abi.ABI_VERSION, C.sizeof(abi.Op), len(self.ops), NUM_BUFS
        h.blob_bytes, h.arena_bytes = int(self.blob_tensor.numel()), int(self.arena_bytes)
        h.info = self.info()
        with open(path, "wb") as f:
            f.write(bytes(h))
            f.write(bytes(self.op_array))
            f.write(self.blob_tensor.numpy().tobytes())

    # ------------------------------------------------------------------------------------------ helpers
    def _alloc_bytes(self, nbytes):
        off = (self.arena_bytes + 255) // 256 * 256
        self.arena_bytes = off + nbytes
        self._allocs.append((off, nbytes))
        return off

    _VIEW_FIELDS = ("x", "x2", "y", "w", "scale", "bias", "a0", "a1", "a2")
    # The library may run up to 4 consecutive ops as ONE kernel (a CBAM, a ShuffleNet branch, a RepGhost module); that
    # kernel reads the first op's input while it writes the last op's output, so a tensor stays reserved this many ops
    # past its last reference.
    _FUSION_SPAN = 3

    def _assign_arena_by_liveness(self):
        """Replace the bump offsets by addresses shared between tensors whose lifetimes (first .. last op that
        references them) do not overlap: the arena shrinks ~3x and a consumer's output lands on lines a dead tensor
        left in L2.  Offsets inside the op list are patched in place; the op list stays relocatable."""
        import bisect
        starts = [a[0] for a in self._allocs]
        first = [None] * len(starts)
        last = [None] * len(starts)
        refs = []                                   # (op, field, allocation index)
        for i, op in enumerate(self.ops):
            for f in self._VIEW_FIELDS:
                v = getattr(op, f)
                if v.buf != BUF_ARENA:
                    continue
                a = bisect.bisect_right(starts, v.off) - 1
                assert a >= 0 and v.off < starts[a] + max(self._allocs[a][1], 1), (self.op_names[i], f)
                refs.append((op, f, a))
                first[a] = i if first[a] is None else first[a]
                last[a] = i
        placed = []                                 # (new offset, bytes, first, last)
        new_off = [0] * len(starts)
        total = 0
        order = sorted((a for a in range(len(starts)) if first[a] is not None), key=lambda a: (first[a], -self._allocs[a][1]))
        for a in order:
            nbytes = (self._allocs[a][1] + 255) // 256 * 256
            lo, hi = first[a], last[a] + self._FUSION_SPAN
            busy = sorted((o, o + b) for (o, b, f0, l0) in placed if not (l0 < lo or f0 > hi))
            off = 0
            for (o, e) in busy:                     # first fit among the tensors alive at the same time
                if off + nbytes <= o:
                    break
                off = max(off, e)
            placed.append((off, nbytes, lo, hi))
            new_off[a] = off
            total = max(total, off + nbytes)
        for (op, f, a) in refs:
            v = getattr(op, f)
            v.off = v.off - starts[a] + new_off[a]
            setattr(op, f, v)
        self.arena_bytes = total

    def _tensor(self, n, h, w, c, gi=0):
        off = self._alloc_bytes(n * h * w * c * 2)
        t = TRef(off, n, h, w, c, 0, c)
        return t

    def _emit(self, name, op):
        self.ops.append(op)
        self.op_names.append(name)

    def _bn(self, prefix, eps):
        s = self.sd
        scale = s[prefix + '.weight'] / torch.sqrt(s[prefix + '.running_var'] + eps)
        return scale, s[prefix + '.bias'] - s[prefix + '.running_mean'] * scale

    @staticmethod
    def _permute_in(w, src, col_scale=None):
        """Re-index a conv weight's input channels from logical to the physical order of `src`."""
        if col_scale is not None:
            w = w * col_scale.view(1, -1, 1, 1)
        if src.perm is None:
            return w
        out = torch.empty_like(w)
        out[:, src.perm_tensor()] = w
        return out

    def conv(self, name, src, dst, groups, k, stride, act, res=None, f32_out=None, post_scale=1.0, split=None, macs_per_pixel=None):
        """groups: list (1 or 2 entries) of (weight [Cout,Cin,k,k] logical order, scale [Cout], bias [Cout],
        col_scale or None).  dst: TRef (bf16 NHWC) or None with f32_out = (buf, ctot, coff).
        split = (c, TRef): output channels from c on go to that second tensor instead of dst[..., c:]."""
        packed, meta = [], None
        scs, bis = [], []
        # pixel-pair k-blocks for 3x3 stride-2 convs on dense 32-channel inputs of even width (see pack_conv_weight_pair)
        pair = (k == 3 and stride == 2 and src.c == 32 and src.ld == 32 and src.c_off == 0 and src.w % 2 == 0 and
                res is None and f32_out is None)
        for (w, sc, bi, col) in groups:
            assert w.shape[1] == src.c and w.shape[2] == k, (name, tuple(w.shape), src.c)
            wp = self._permute_in(w, src, col)
            p, meta = pack.pack_conv_weight_pair(wp) if pair else pack.pack_conv_weight(wp)
            packed.append(p)
            npad = meta['BN'] * meta['n_tiles']
            scs.append(pack.pad_channels(sc, npad))
            bis.append(pack.pad_channels(bi, npad))
        w_off = self.blob.add(torch.cat(packed))
        sc_off = self.blob.add(torch.cat(scs))
        bi_off = self.blob.add(torch.cat(bis))
        ho = src.h if stride == 1 else down2(src.h)
        wo = src.w if stride == 1 else down2(src.w)
        g = len(groups)
        op = abi.new_op(abi.OP_CONV, act=act, x=src.view(), w=_flat(BUF_BLOB, w_off), scale=_flat(BUF_BLOB, sc_off),
                        bias=_flat(BUF_BLOB, bi_off), n_img=src.n, group_imgs=src.n // g, Hi=src.h, Wi=src.w, Cin=src.c,
                        Ho=ho, Wo=wo, Cout=meta['Cout'], ksize=k, stride=stride, BN=meta['BN'], n_tiles=meta['n_tiles'],
                        k_blocks=meta['k_blocks'], K_real=meta['K_real'], w_gstride=packed[0].numel(),
                        sb_gstride=meta['BN'] * meta['n_tiles'], f0=post_scale,
                        flags=meta['bk'] | (abi.CONV_FLAG_PAIR if pair else 0))
        if f32_out is None:
            assert dst.n == src.n and dst.h == ho and dst.w == wo and dst.c == meta['Cout'], name
            op.out_mode = abi.OUT_BF16_NHWC
            op.y = dst.view()
            if split is not None:
                c_split, dst2 = split
                assert res is None and dst2.c == meta['Cout'] - c_split and dst2.n == dst.n and dst2.h == ho and dst2.w == wo
                op.parts = c_split
                op.a0 = dst2.view()
        else:
            buf, ctot, coff = f32_out
            op.out_mode = abi.OUT_F32_NCHW
            op.y = abi.View(buf, 0, 0, ctot * ho * wo, 0, 0, 0)
            op.out_ctot, op.out_coff = ctot, coff
        if res is not None:
            op.x2 = res.view()
        # algorithmic FLOPs of the reference's convolutions (a block-diagonal merge does not count its zero blocks)
        self.conv_flops += 2 * src.n * ho * wo * (macs_per_pixel if macs_per_pixel is not None else meta['Cout'] * meta['K_real'])
        self._emit(name, op)

    def dwconv(self, name, src, dst, groups, act, res=None):
        """groups: list of (w [C,1,3,3] with BN already folded, bias [C])."""
        assert src.perm is None
        w_off = self.blob.add(torch.cat([w.reshape(src.c, 9).t().contiguous().reshape(-1) for (w, _) in groups]))
        b_off = self.blob.add(torch.cat([b for (_, b) in groups]))
        op = abi.new_op(abi.OP_DWCONV, act=act, x=src.view(), y=dst.view(), w=_flat(BUF_BLOB, w_off),
                        bias=_flat(BUF_BLOB, b_off), n_img=src.n, group_imgs=src.n // len(groups), Hi=src.h, Wi=src.w,
                        Cin=src.c, Ho=src.h, Wo=src.w, Cout=src.c, ksize=3, stride=1)
        if res is not None:
            op.x2 = res.view()
        self.conv_flops += 2 * src.n * src.h * src.w * src.c * 9
        self._emit(name, op)

    def cbam(self, name, src, dst, prefixes):
        """CBAM over src -> dst (same physical channel order).  prefixes: one state_dict prefix per group."""
        s = self.sd
        c, n, hw = src.c, src.n, src.h * src.w
        perm = src.perm_tensor()
        fc1s, fc2s, w7s = [], [], []
        for p in prefixes:
            fc1 = s[p + '.channelattention.fc1.weight'].reshape(-1, c)   # [hidden, C] logical
            fc2 = s[p + '.channelattention.fc2.weight'].reshape(c, -1)   # [C, hidden] logical
            f1 = torch.empty_like(fc1)
            f1[:, perm] = fc1
            f2 = torch.empty_like(fc2)
            f2[perm, :] = fc2
            fc1s.append(f1.reshape(-1))
            fc2s.append(f2.reshape(-1))
            w7 = s[p + '.spatialattention.conv1.weight']
            assert tuple(w7.shape) == (1, 2, 7, 7), "only kernel_size=7 spatial attention is used by the reference"
            w7s.append(w7.reshape(-1))
        hidden = fc1s[0].numel() // c
        f1_off = self.blob.add(torch.cat(fc1s))
        f2_off = self.blob.add(torch.cat(fc2s))
        w7_off = self.blob.add(torch.cat(w7s))
        parts = min(64, max(1, hw // 128))
        psum = self._alloc_bytes(n * parts * c * 4)
        pmax = self._alloc_bytes(n * parts * c * 4)
        gate = self._alloc_bytes(n * c * 4)
        stats = self._alloc_bytes(n * hw * 2 * 4)
        common = dict(n_img=n, group_imgs=n // len(prefixes), Hi=src.h, Wi=src.w, Cin=c, hidden=hidden, parts=parts)
        A = BUF_ARENA
        self._emit(name + '.pool', abi.new_op(abi.OP_CBAM_POOL, x=src.view(), a0=_flat(A, psum), a1=_flat(A, pmax), **common))
        self._emit(name + '.mlp', abi.new_op(abi.OP_CBAM_MLP, a0=_flat(A, psum), a1=_flat(A, pmax), a2=_flat(A, gate),
                                             w=_flat(BUF_BLOB, f1_off), scale=_flat(BUF_BLOB, f2_off),
                                             w_gstride=hidden * c, sb_gstride=hidden * c, **common))
        self._emit(name + '.stats', abi.new_op(abi.OP_CBAM_STATS, x=src.view(), a2=_flat(A, gate), a0=_flat(A, stats), **common))
        self._emit(name + '.apply', abi.new_op(abi.OP_CBAM_APPLY, x=src.view(), a2=_flat(A, gate), a0=_flat(A, stats),
                                               w=_flat(BUF_BLOB, w7_off), y=dst.view(), **common))
        self.conv_flops += 2 * n * (4 * hidden * c + hw * 98)

    # ------------------------------------------------------------------------------------------ network
    def _conv_group(self, prefix, eps, col_scale=None):
        """(weight, scale, bias, col_scale) of a Conv(conv+bn) block."""
        sc, bi = self._bn(prefix + '.bn', eps)
        return (self.sd[prefix + '.conv.weight'], sc, bi, col_scale)

    def _shuffle_unit(self, name, x, prefixes, x2_src=None):
        """ShuffleNetV2 stride-1 unit (nets/yolo_mul.py:138-168): branch 2 reads x2_src (or, in place, the second half
        of x) and writes the second half of x.  Returns x with the shuffle perm."""
        s = self.sd
        c = x.c
        h = c // 2
        x2 = x.sub(h, h)
        x2_in = x2_src if x2_src is not None else x2
        t1 = self._tensor(x.n, x.h, x.w, h)
        t2 = self._tensor(x.n, x.h, x.w, h)
        g0, gdw, g2 = [], [], []
        for p in prefixes:
            b = p + '.branch2'
            sc, bi = self._bn(b + '.1', 1e-5)
            g0.append((s[b + '.0.weight'], sc, bi, None))
            sc, bi = self._bn(b + '.4', 1e-5)
            gdw.append((s[b + '.3.weight'] * sc.view(-1, 1, 1, 1), s[b + '.3.bias'] * sc + bi))
            sc, bi = self._bn(b + '.6', 1e-5)
            g2.append((s[b + '.5.weight'], sc, bi, None))
        self.conv(name + '.pw1', x2_in, t1, g0, 1, 1, abi.ACT_RELU)
        if x2_src is not None:   # t1, t2 are private and the output does not alias the input: fusable chain
            self.ops[-1].flags |= abi.CONV_FLAG_CHAIN_HEAD
        self.dwconv(name + '.dw', t1, t2, gdw, abi.ACT_NONE)
        self.conv(name + '.pw2', t2, x2, g2, 1, 1, abi.ACT_RELU)
        # logical channel l of the shuffled tensor: even -> x1[l/2], odd -> branch2[(l-1)/2]
        perm = [(l % 2) * h + l // 2 for l in range(c)]
        return TRef(x.off, x.n, x.h, x.w, x.ld, x.c_off, c, x.gi, x.gstride, perm)

    def _repghost_bottleneck(self, name, p, x, dst):
        """RepGhostBottleneck (nets/repghost.py:263-279) with both RepGhostModules in deploy algebra (:117-123)."""
        s = self.sd
        c = x.c
        u1 = self._tensor(x.n, x.h, x.w, c)
        u2 = self._tensor(x.n, x.h, x.w, c)
        v1 = self._tensor(x.n, x.h, x.w, c)

        def module(gp):
            sc, bi = self._bn(gp + '.primary_conv.1', 1e-5)
            pw = (s[gp + '.primary_conv.0.weight'], sc, bi, None)
            s1, b1 = self._bn(gp + '.cheap_operation.1', 1e-5)
            s2, b2 = self._bn(gp + '.fusion_bn.0', 1e-5)
            w = s[gp + '.cheap_operation.0.weight'] * s1.view(-1, 1, 1, 1)
            w = w.clone()
            w[:, 0, 1, 1] += s2
            return pw, (w, b1 + b2)

        pw1, dw1 = module(p + '.ghost1')
        pw2, dw2 = module(p + '.ghost2')
        # u1 / v1 are private and no module output aliases its input: each (1x1 conv, depthwise) pair may run fused
        self.conv(name + '.g1.pw', x, u1, [pw1], 1, 1, abi.ACT_SILU)
        self.ops[-1].flags |= abi.CONV_FLAG_GHOST_HEAD
        self.dwconv(name + '.g1.dw', u1, u2, [dw1], abi.ACT_SILU)
        self.conv(name + '.g2.pw', u2, v1, [pw2], 1, 1, abi.ACT_NONE)
        self.ops[-1].flags |= abi.CONV_FLAG_GHOST_HEAD
        self.dwconv(name + '.g2.dw', v1, dst, [dw2], abi.ACT_NONE, res=x)

    def _c2f(self, name, p, src, dst, depth, col_scale=None):
        """C2f_repghost (nets/repghost.py:308-320): cv1 -> split -> depth x bottleneck -> cv2 over the concat."""
        c = dst.c // 2
        ycat = self._tensor(src.n, src.h, src.w, (2 + depth) * c)
        self.conv(name + '.cv1', src, ycat.sub(0, 2 * c), [self._conv_group(p + '.cv1', 1e-5, col_scale)], 1, 1, abi.ACT_SILU)
        for i in range(depth):
            self._repghost_bottleneck('%s.m%d' % (name, i), '%s.m.%d' % (p, i), ycat.sub((1 + i) * c, c), ycat.sub((2 + i) * c, c))
        self.conv(name + '.cv2', ycat, dst, [self._conv_group(p + '.cv2', 1e-5)], 1, 1, abi.ACT_SILU)

    def _build(self):
        s = self.sd
        B, H, W, nc = self.B, self.H, self.W, self.nc
        bc, depth, c3, c4, c5 = model_dims(self.phi)
        N2 = 2 * B
        mods = ('backbone_rgb', 'backbone_nir')

        # ---- stems (nets/yolo_mul.py:104-115): replicated, sign-folded bf16 weight tile per modality (pack.pack_stem)
        packed, scs, bis, c0pad = [], [], [], 32
        for m in mods:
            sc, bi = self._bn(m + '.stem.conv.1', 1e-5)
            pk, sca, bia, c0pad = pack.pack_stem(s[m + '.stem.conv.0.weight'], sc, bi, u8=self.input_u8)
            packed.append(pk)
            scs.append(sca)
            bis.append(bia)
        w_off = self.blob.add(torch.cat(packed))
        sc_off = self.blob.add(torch.cat(scs))
        b_off = self.blob.add(torch.cat(bis))
        h1, w1 = down2(H), down2(W)
        x = self._tensor(N2, h1, w1, bc)
        self._emit('stem', abi.new_op(abi.OP_STEM, x=_flat(BUF_RGB, 0), x2=_flat(BUF_NIR, 0), w=_flat(BUF_BLOB, w_off),
                                      scale=_flat(BUF_BLOB, sc_off), bias=_flat(BUF_BLOB, b_off), y=x.view(), n_img=N2,
                                      group_imgs=B, Hi=H, Wi=W, Ho=h1, Wo=w1, Cout=bc, Cin=3, ksize=3, stride=1,
                                      BN=c0pad, n_tiles=1, k_blocks=1, K_real=27, w_gstride=packed[0].numel(), sb_gstride=c0pad,
                                      flags=(abi.STEM_FLAG_U8 if self.input_u8 else 0) |
                                      (abi.STEM_FLAG_X2_PLANE if self.depth_plane else 0)))
        self.conv_flops += 2 * N2 * H * W * bc * 27

        # ---- dark2..dark5 (nets/yolo_mul.py:258-277)
        feats = {}
        for stage, cout in (('dark2', 2 * bc), ('dark3', c3), ('dark4', c4), ('dark5', c5)):
            # The stage conv writes its first half straight into the unit's output tensor and its second half (the
            # unit's branch-2 input) into a separate tensor: the fused unit kernel reads halo pixels of its input
            # while neighbouring tiles already write their outputs, so input and output must not alias.
            y = self._tensor(N2, down2(x.h), down2(x.w), cout)
            y2 = self._tensor(N2, down2(x.h), down2(x.w), cout // 2)
            groups = [self._conv_group('%s.%s.0' % (m, stage), 1e-3) for m in mods]
            if (cout // 2) % 16 == 0:
                self.conv(stage + '.0', x, y, groups, 3, 2, abi.ACT_SILU, split=(cout // 2, y2))
            else:   # 16-channel granularity of the split store: narrow models copy through the unfused path
                self.conv(stage + '.0', x, y, groups, 3, 2, abi.ACT_SILU)
                y2 = None
            x = self._shuffle_unit(stage + '.1', y, ['%s.%s.1' % (m, stage) for m in mods], y2)
            feats[stage] = x
        feat1, feat2 = feats['dark3'], feats['dark4']

        # ---- SPPF_CBAM (nets/yolo_mul.py:10-32)
        c_ = c5 // 2
        t = self._tensor(N2, x.h, x.w, c_)
        self.conv('sppf.cv1', x, t, [self._conv_group(m + '.dark5.2.cv1', 1e-3) for m in mods], 1, 1, abi.ACT_SILU)
        cat = self._tensor(N2, x.h, x.w, 4 * c_)
        self.cbam('sppf.cbam1', t, cat.sub(0, c_), [m + '.dark5.2.cbam1' for m in mods])
        for i in (1, 2, 3):
            pooled = self._tensor(N2, x.h, x.w, c_)
            self._emit('sppf.pool%d' % i, abi.new_op(abi.OP_MAXPOOL5, x=cat.sub((i - 1) * c_, c_).view(), y=pooled.view(),
                                                     n_img=N2, Hi=x.h, Wi=x.w, Cin=c_, Ho=x.h, Wo=x.w, Cout=c_))
            self.cbam('sppf.cbam%d' % (i + 1), pooled, cat.sub(i * c_, c_), [m + '.dark5.2.cbam%d' % (i + 1) for m in mods])
        feat3 = self._tensor(N2, x.h, x.w, c5)
        self.conv('sppf.cv2', cat, feat3, [self._conv_group(m + '.dark5.2.cv2', 1e-3) for m in mods], 1, 1, abi.ACT_SILU)

        # ---- fusion: six CBAMs written straight into the BiFPN concat buffers (nets/yolo_mul.py:403-443)
        h3, w3, h4, w4, h5, w5 = feat1.h, feat1.w, feat2.h, feat2.w, feat3.h, feat3.w
        cc1 = self._tensor(B, h4, w4, c5 + 2 * c4)   # [P5_up | feat2_rgb' | feat2_nir']   (:428)
        cc2 = self._tensor(B, h3, w3, c4 + 2 * c3)   # [P4_up | feat1_rgb' | feat1_nir']   (:435)
        cc3 = self._tensor(B, h4, w4, c3 + c4)       # [down(P3) | P4]                     (:439)
        cc4 = self._tensor(B, h5, w5, c4 + 2 * c5)   # [down(P4) | feat3_rgb' | feat3_nir'] (:443)

        def slot2(buf, c_off, c):  # both modalities as channel-adjacent slots of a [B,...] buffer
            return TRef(buf.off, N2, buf.h, buf.w, buf.ld, c_off, c, gi=B, gstride=c)

        self.cbam('cbam_feat1', feat1, slot2(cc2, c4, c3), ['cbam_rgb_feat1', 'cbam_nir_feat1'])
        self.cbam('cbam_feat2', feat2, slot2(cc1, c5, c4), ['cbam_rgb_feat2', 'cbam_nir_feat2'])
        self.cbam('cbam_feat3', feat3, slot2(cc4, c4, c5), ['cbam_rgb_feat3', 'cbam_nir_feat3'])

        wn = s['bi_fpn.w'] / (s['bi_fpn.w'].sum() + 0.0001)   # nets/yolo_mul.py:46

        def bifpn_cols(c_first, c_lat):
            return torch.cat([wn[0].repeat(c_first), wn[1].repeat(c_lat), wn[2].repeat(c_lat)])

        def concat_perm(c_first, lat):  # physical order of [first | lat_rgb | lat_nir] given the lateral's perm
            lp = lat.perm_tensor()
            return torch.cat([torch.arange(c_first), c_first + lp, c_first + lat.c + lp]).tolist()

        def upsample(name, a, b, dst):
            op = abi.new_op(abi.OP_UPSAMPLE, x=a.view(), y=dst.view(), n_img=a.n, Hi=a.h, Wi=a.w, Ho=dst.h, Wo=dst.w,
                            Cin=a.c, Cout=a.c)
            if b is not None:
                op.x2 = b.view()
            self._emit(name, op)

        # P5_up = bilinear(feat3_rgb' + feat3_nir')  (:421,:426)
        upsample('p5_up', cc4.sub(c4, c5), cc4.sub(c4 + c5, c5), cc1.sub(0, c5))
        cc1.perm = concat_perm(c5, feat2)
        self._c2f('up1', 'conv3_for_upsample1', cc1, cc3.sub(c3, c4), depth, bifpn_cols(c5, c4))       # P4 (:430)
        upsample('p4_up', cc3.sub(c3, c4), None, cc2.sub(0, c4))                                       # (:433)
        cc2.perm = concat_perm(c4, feat1)
        p3 = self._tensor(B, h3, w3, c3)
        self._c2f('up2', 'conv3_for_upsample2', cc2, p3, depth, bifpn_cols(c4, c3))                     # P3 (:436)
        self.conv('down1', p3, cc3.sub(0, c3), [self._conv_group('down_sample1', 1e-3)], 3, 2, abi.ACT_SILU)  # (:438)
        p4 = self._tensor(B, h4, w4, c4)
        self._c2f('dn1', 'conv3_for_downsample1', cc3, p4, depth)                                       # (:440)
        self.conv('down2', p4, cc4.sub(0, c4), [self._conv_group('down_sample2', 1e-3)], 3, 2, abi.ACT_SILU)  # (:442)
        p5 = self._tensor(B, h5, w5, c5)
        cc4.perm = concat_perm(c4, feat3)
        self._c2f('dn2', 'conv3_for_downsample2', cc4, p5, depth, bifpn_cols(c4, c5))                   # (:444)

        # ---- decoupled head (nets/yolo_mul.py:387-391, :451-453): fp32 NCHW maps x[i] = [box 64 | cls nc]
        c2h = max(16, c3 // 4, 64)
        c3h = max(c3, nc)
        no = 64 + nc
        self.level_shapes = [(p.h, p.w) for p in (p3, p4, p5)]
        self.A = sum(h * w for h, w in self.level_shapes)
        assert self.level_shapes[1] == (down2(h3), down2(w3)) and self.level_shapes[2] == (down2(h4), down2(w4))
        assert torch.equal(s['dfl.conv.weight'].reshape(-1), torch.arange(16.0)), \
            "dfl.conv.weight must be arange(16) (frozen in the reference, nets/yolo_mul.py:315-317)"
        a_off = 0
        for i, (p, buf) in enumerate(((p3, BUF_X0), (p4, BUF_X1), (p5, BUF_X2))):
            a = self._conv_group('cv2.%d.0' % i, 1e-3)
            b = self._conv_group('cv3.%d.0' % i, 1e-3)
            merged = (torch.cat([a[0], b[0]]), torch.cat([a[1], b[1]]), torch.cat([a[2], b[2]]), None)
            h1t = self._tensor(B, p.h, p.w, c2h + c3h)
            self.conv('head%d.0' % i, p, h1t, [merged], 3, 1, abi.ACT_SILU)
            # second convs of both branches write the two channel slots of ONE tensor, so that the two final biased 1x1
            # convs (:389, :391) run as a single block-diagonal GEMM (K = c2h + c3h, N = 64 + nc) whose epilogue writes
            # x[i] = cat(box, cls) (:453), DFL(box) (:461) and the gathered class logits (:459-460)
            hbc = self._tensor(B, p.h, p.w, c2h + c3h)
            self.conv('head%d.box1' % i, h1t.sub(0, c2h), hbc.sub(0, c2h), [self._conv_group('cv2.%d.1' % i, 1e-3)], 3, 1, abi.ACT_SILU)
            self.conv('head%d.cls1' % i, h1t.sub(c2h, c3h), hbc.sub(c2h, c3h), [self._conv_group('cv3.%d.1' % i, 1e-3)], 3, 1, abi.ACT_SILU)
            wout = torch.zeros(no, c2h + c3h, 1, 1)
            wout[:64, :c2h] = s['cv2.%d.2.weight' % i]
            wout[64:, c2h:] = s['cv3.%d.2.weight' % i]
            bout = torch.cat([s['cv2.%d.2.bias' % i], s['cv3.%d.2.bias' % i]])
            self.conv('head%d.out' % i, hbc, None, [(wout, torch.ones(no), bout, None)], 1, 1, abi.ACT_NONE, f32_out=(buf, no, 0),
                      macs_per_pixel=64 * c2h + nc * c3h)
            op = self.ops[-1]
            if (op.flags & 0xff) != 0:   # TMA packing: fused DFL epilogue; else (odd widths) a separate DFL pass below
                op.flags |= abi.CONV_FLAG_DFL
                op.a1, op.a2 = _flat(BUF_DBOX, 0), _flat(BUF_CLS, 0)
                op.A, op.nc, op.hidden = self.A, nc, a_off
            a_off += p.h * p.w
        if not all(o.flags & abi.CONV_FLAG_DFL for o, n in zip(self.ops, self.op_names) if n.endswith('.out')):
            for o, n in zip(self.ops, self.op_names):
                if n.endswith('.out'):
                    o.flags &= ~abi.CONV_FLAG_DFL
            # ---- DFL (+ the level gather of :459-460) as its own pass
            self._emit('dfl', abi.new_op(abi.OP_DFL, a0=_flat(BUF_X0, 0), a1=_flat(BUF_X1, 0), a2=_flat(BUF_X2, 0),
                                         y=_flat(BUF_DBOX, 0), x2=_flat(BUF_CLS, 0), n_img=B, Hi=h3, Wi=w3, nc=nc, A=self.A))
        self.no = no
