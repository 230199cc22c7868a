"""Weight folding and packing for the CUDA kernels (host side, fp32 math on CPU tensors).

Layouts produced here are exactly what include/dcfa_b200.h documents per op kind."""
import math

import torch

BK = 64  # K elements per pipeline stage of the conv GEMM (one 128-byte swizzle row of bf16)


def fold_bn(bn):
    """Eval-mode BatchNorm2d as per-channel (scale, bias): y = x*scale + bias (nn.BatchNorm2d semantics)."""
    w = bn.weight.detach().float().cpu() if bn.weight is not None else torch.ones(bn.num_features)
    b = bn.bias.detach().float().cpu() if bn.bias is not None else torch.zeros(bn.num_features)
    mean = bn.running_mean.detach().float().cpu()
    var = bn.running_var.detach().float().cpu()
    scale = w / torch.sqrt(var + bn.eps)
    return scale, b - mean * scale


def conv_tiling(cout, max_bn=256):
    """(BN, n_tiles) for a conv with `cout` output channels: UMMA N must be a multiple of 16 in [16, max_bn <= 256]."""
    n_tiles = max(1, math.ceil(cout / max_bn))
    bn = 16 * math.ceil(math.ceil(cout / n_tiles) / 16)
    return bn, n_tiles


def pick_bk(cin):
    """Channels per k-block of the TMA conv path (0: Cin is not a multiple of 16 -> cp.async gather path, flat K)."""
    for bk in (64, 32, 16):
        if cin % bk == 0:
            return bk
    return 0


def swizzle_tile(tile):
    """[rows, bk] bf16 K-major tile (bk = 64 / 32 / 16) -> the swizzled image the UMMA descriptor reads
    (SWIZZLE_128B / 64B / 32B): with n = bk/8 16-byte chunks per row, chunk c of row r is stored at chunk
    position c ^ ((r * n // 8) % n)  (address bits [4,4+log2 n) XOR bits [7,7+log2 n))."""
    rows, bk = tile.shape
    n = bk // 8
    t = tile.reshape(rows, n, 8)
    idx = (torch.arange(n)[None, :] ^ ((torch.arange(rows) * n // 8) % n)[:, None])  # XOR is an involution
    return torch.gather(t, 1, idx[:, :, None].expand(rows, n, 8)).reshape(rows, bk)


def pack_conv_weight_pair(w):
    """3x3 stride-2 conv with Cin = 32 on a dense even-width input (DCFA_CONV_FLAG_PAIR): k-blocks of 64 = one PIXEL PAIR
    (columns 2p, 2p+1: 128 contiguous bytes).  Output column ox reads input columns 2ox-1, 2ox, 2ox+1 = pair (ox-1).odd,
    pair ox.even, pair ox.odd, so per kernel row two k-blocks: [0 | w(dy,0)] on pair ox-1 and [w(dy,1) | w(dy,2)] on pair ox
    -- six 128-byte-row boxes per tile instead of nine 64-byte-row ones (the TMA unit's cost is per box row)."""
    cout, cin, k, _ = w.shape
    assert k == 3 and cin == 32
    bn, n_tiles = conv_tiling(cout)
    full = torch.zeros(n_tiles * bn, 6 * 64, dtype=torch.float32)
    for dy in range(3):
        full[:cout, (2 * dy) * 64 + 32:(2 * dy) * 64 + 64] = w[:, :, dy, 0]
        full[:cout, (2 * dy + 1) * 64:(2 * dy + 1) * 64 + 32] = w[:, :, dy, 1]
        full[:cout, (2 * dy + 1) * 64 + 32:(2 * dy + 1) * 64 + 64] = w[:, :, dy, 2]
    full = full.to(torch.bfloat16)
    tiles = [swizzle_tile(full[nt * bn:(nt + 1) * bn, kb * 64:(kb + 1) * 64]).reshape(-1) for nt in range(n_tiles) for kb in range(6)]
    meta = dict(BN=bn, n_tiles=n_tiles, k_blocks=6, K_real=9 * cin, Cout=cout, Cin=cin, ksize=3, bk=64, pair=True)
    return torch.cat(tiles), meta


def pack_conv_weight(w, bk=None, max_bn=256):
    """w: [Cout, Cin, k, k] fp32 (input channels already in PHYSICAL order) ->
    (packed bf16 [n_tiles*k_blocks*BN*bk], meta dict).  K index = (ky*k + kx)*Cin + ci.
    bk > 0: one k-block per (tap, channel block of bk) -- the TMA path; bk == 0: flat K padded to 64 (gather path)."""
    cout, cin, k, _ = w.shape
    if bk is None:
        bk = pick_bk(cin)
    bn, n_tiles = conv_tiling(cout, max_bn)
    k_real = k * k * cin
    tile_k = bk if bk else BK
    k_blocks = math.ceil(k_real / tile_k)
    wk = w.permute(0, 2, 3, 1).reshape(cout, k_real)
    full = torch.zeros(n_tiles * bn, k_blocks * tile_k, dtype=torch.float32)
    full[:cout, :k_real] = wk
    full = full.to(torch.bfloat16)
    tiles = []
    for nt in range(n_tiles):
        for kb in range(k_blocks):
            tiles.append(swizzle_tile(full[nt * bn:(nt + 1) * bn, kb * tile_k:(kb + 1) * tile_k]).reshape(-1))
    packed = torch.cat(tiles)
    meta = dict(BN=bn, n_tiles=n_tiles, k_blocks=k_blocks, K_real=k_real, Cout=cout, Cin=cin, ksize=k, bk=bk)
    return packed, meta


def pack_stem(w, scale, bias, u8=False):
    """Stem conv [C0,3,3,3] + folded BN -> (weight tiles bf16 [nblk*3*128*16], scale [C0pad], bias [C0pad], C0pad).

    The stem kernel never builds an im2col: its B operand is the bf16 [y][x][4 channel] input patch itself, read through a
    no-swizzle K-major descriptor whose row n is the 16 values starting at pixel 2n (four pixels x 4 channel slots).
    A conv pixel of even column 2n taps pixels 2n..2n+2 (K = kx*4 + ci), one of odd column 2n+1 taps pixels 2n+1..2n+3
    (K = (kx+1)*4 + ci); unused slots are zero.  Both PARITIES share the operand rows, so they are stacked on the 128 MMA
    rows of ONE weight tile per kernel row ky.  Row m = 32*q + 16*e + cl: TMEM lane quarter q (one epilogue warp), column
    parity e, channel lane cl -- the two parities of a channel sit 16 lanes apart in the SAME warp, which swaps its
    column maxima with one shuffle.  Row slot j = 16*q + cl holds channel 64*blk + j % min(C0pad, 64)  (C0pad = 32: two
    replicas of the 32 channels, each pooling other tile rows; C0pad = 128: two blocks of 64 channels, one launch each).
    Tiles are stored in the canonical no-swizzle layout of the tcgen05 descriptor: element (m, k) at
    (m//8)*128 + (k//8)*64 + (m%8)*8 + k%8  (8 rows x 16 bytes core matrices; LBO = 128 B, SBO = 256 B).

    u8=True: the scale carries preprocess_input's 1/255 (utils/utils.py:76-79) -- the kernel multiplies exact integer
    pixels.

    The kernel pools BEFORE applying scale/bias/ReLU, which needs a non-negative scale: channels with a negative BN scale
    get negated weights and |scale| (scale * conv(w) == |scale| * conv(sign * w))."""
    c0 = w.shape[0]
    c0pad = 32 if c0 <= 32 else (64 if c0 <= 64 else 128)
    nblk, cb = max(1, c0pad // 64), min(c0pad, 64)
    sgn = torch.where(scale < 0, -torch.ones_like(scale), torch.ones_like(scale))
    ws = (w * sgn.view(-1, 1, 1, 1)).float()                      # [c, ci, ky, kx]
    tiles = torch.zeros(nblk, 3, 128, 16, dtype=torch.float32)
    rows = torch.arange(128)
    e_of = (rows % 32) // 16
    slot = (rows // 32) * 16 + rows % 16
    for blk in range(nblk):
        src = 64 * blk + slot % cb
        for e in range(2):
            sel = (e_of == e) & (src < c0)
            for ky in range(3):
                for kx in range(3):
                    for ci in range(3):
                        tiles[blk, ky, rows[sel], (kx + e) * 4 + ci] = ws[src[sel], ci, ky, kx]
    t = tiles.reshape(nblk, 3, 16, 8, 2, 8).permute(0, 1, 2, 4, 3, 5)   # [blk, ky, m/8, k/8, m%8, k%8]
    packed = t.reshape(-1).to(torch.bfloat16)
    sc = scale.abs() / 255.0 if u8 else scale.abs()
    return packed, pad_channels(sc, c0pad), pad_channels(bias, c0pad), c0pad


def unpack_stem(packed, c0):
    """Inverse of pack_stem's weight layout (test infrastructure uses it): bf16 [nblk*3*128*16] -> fp32 [c0,3,3,3] (sign-folded)."""
    nblk = packed.numel() // (3 * 128 * 16)
    t = packed.float().reshape(nblk, 3, 16, 2, 8, 8).permute(0, 1, 2, 4, 3, 5).reshape(nblk, 3, 128, 16)
    w = torch.zeros(c0, 3, 3, 3)
    for c in range(c0):
        blk, j = c // 64, c % 64
        m = (j // 16) * 32 + j % 16          # even-parity row of slot j; the odd-parity row is 16 further
        for ky in range(3):
            for kx in range(3):
                for ci in range(3):
                    w[c, ci, ky, kx] = t[blk, ky, m, kx * 4 + ci]
                    assert t[blk, ky, m + 16, (kx + 1) * 4 + ci] == w[c, ci, ky, kx]
    return w


def pad_channels(v, n, fill=0.0):
    out = torch.full((n,), fill, dtype=torch.float32)
    out[:v.numel()] = v.float()
    return out


class Blob:
    """Append-only parameter blob (bytes) with 256-byte aligned entries."""

    def __init__(self):
        self.parts = []
        self.size = 0

    def add(self, t):
        t = t.contiguous()
        raw = t.view(torch.uint8).reshape(-1) if t.dtype != torch.uint8 else t.reshape(-1)
        pad = (-self.size) % 256
        if pad:
            self.parts.append(torch.zeros(pad, dtype=torch.uint8))
            self.size += pad
        off = self.size
        self.parts.append(raw.clone())
        self.size += raw.numel()
        return off

    def finish(self):
        return torch.cat(self.parts) if self.parts else torch.zeros(0, dtype=torch.uint8)
