"""Small helpers shared by the op-level tests: build dcfa_view/dcfa_op structs over torch tensors."""
import torch

from dcfa_b200 import abi, pack


def nhwc_view(t, buf, c_off=0, gi=0, gstride=0):
    """t: [N,H,W,Ctot] contiguous tensor living in bufs[buf]; view starts at channel c_off."""
    n, h, w, c = t.shape
    return abi.View(buf, c, c_off * t.element_size(), h * w * c, gstride, gi, 0)


def flat_view(buf, off=0):
    return abi.View(buf, 0, off, 0, 0, 0, 0)


def ptrs(tensors):
    return [t.data_ptr() if t is not None else 0 for t in tensors]


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def conv_op(x, w_list, scale_list, bias_list, y, *, ksize, stride, act, cin, c_off_in=0, c_off_out=0, res=None,
            c_off_res=0, out_mode=abi.OUT_BF16_NHWC, out_ctot=0, out_coff=0, group_imgs=0, post_scale=1.0,
            cout=None, in_gi=0, in_gstride=0, out_gi=0, out_gstride=0, bk=None):
    """Returns (op, bufs).  w_list: per-group [Cout,Cin,k,k] fp32 CPU tensors."""
    packed, metas = [], None
    for w in w_list:
        p, metas = pack.pack_conv_weight(w, bk)
        packed.append(p)
    wg = torch.stack(packed).to(x.device)
    npad = metas["BN"] * metas["n_tiles"]
    sc = torch.stack([pack.pad_channels(s, npad, 0.0) for s in scale_list]).to(x.device)
    bi = torch.stack([pack.pad_channels(b, npad, 0.0) for b in bias_list]).to(x.device)
    n, hi, wi, _ = x.shape
    pad = ksize // 2
    ho = (hi + 2 * pad - ksize) // stride + 1
    wo = (wi + 2 * pad - ksize) // stride + 1
    bufs = [x, wg, sc, bi, y, res]
    op = abi.new_op(abi.OP_CONV, act=act, out_mode=out_mode,
                    x=nhwc_view(x, 0, c_off_in, in_gi, in_gstride), w=flat_view(1), scale=flat_view(2), bias=flat_view(3),
                    n_img=n, group_imgs=group_imgs, Hi=hi, Wi=wi, Cin=cin, Ho=ho, Wo=wo,
                    Cout=cout if cout is not None else metas["Cout"], ksize=ksize, stride=stride,
                    BN=metas["BN"], n_tiles=metas["n_tiles"], k_blocks=metas["k_blocks"], K_real=metas["K_real"],
                    w_gstride=packed[0].numel(), sb_gstride=npad, f0=post_scale, out_ctot=out_ctot, out_coff=out_coff,
                    flags=metas["bk"])
    if out_mode == abi.OUT_BF16_NHWC:
        op.y = nhwc_view(y, 4, c_off_out, out_gi, out_gstride)
    else:
        op.y = abi.View(4, 0, 0, out_ctot * ho * wo, 0, 0, 0)
    if res is not None:
        op.x2 = nhwc_view(res, 5, c_off_res)
    return op, bufs
