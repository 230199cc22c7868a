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


def conv_tiling(cout):
    """(BN, n_tiles) for a conv with `cout` output channels: UMMA N must be a multiple of 16 in [16, 256]."""
    n_tiles = max(1, math.ceil(cout / 256))
    bn = 16 * math.ceil(math.ceil(cout / n_tiles) / 16)
    return bn, n_tiles


def swizzle_tile(tile):
    """[rows, 64] bf16 K-major tile -> the 128B-swizzled image the UMMA descriptor (SWIZZLE_128B) reads:
    16-byte chunk c of row r is stored at chunk position c ^ (r % 8)."""
    rows = tile.shape[0]
    t = tile.reshape(rows, 8, 8)
    idx = (torch.arange(8)[None, :] ^ (torch.arange(rows) % 8)[:, None])  # XOR is an involution
    return torch.gather(t, 1, idx[:, :, None].expand(rows, 8, 8)).reshape(rows, 64)


def pack_conv_weight(w):
    """w: [Cout, Cin, k, k] fp32 (input channels already in PHYSICAL order) ->
    (packed bf16 [n_tiles*k_blocks*BN*64], meta dict).  K index = (ky*k + kx)*Cin + ci."""
    cout, cin, k, _ = w.shape
    bn, n_tiles = conv_tiling(cout)
    k_real = k * k * cin
    k_blocks = math.ceil(k_real / BK)
    wk = w.permute(0, 2, 3, 1).reshape(cout, k_real)
    full = torch.zeros(n_tiles * bn, k_blocks * BK, dtype=torch.float32)
    full[:cout, :k_real] = wk
    full = full.to(torch.bfloat16)
    tiles = []
    for nt in range(n_tiles):
        for kb in range(k_blocks):
            tiles.append(swizzle_tile(full[nt * bn:(nt + 1) * bn, kb * BK:(kb + 1) * BK]).reshape(-1))
    packed = torch.cat(tiles)
    meta = dict(BN=bn, n_tiles=n_tiles, k_blocks=k_blocks, K_real=k_real, Cout=cout, Cin=cin, ksize=k)
    return packed, meta


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
