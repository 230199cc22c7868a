"""TEST INFRASTRUCTURE -- CPU fp32 restatement of the reference's hot path as plain functions of a state_dict.

Nothing under `dcfa-yolo_b200/` imports this file; it is the checker for tests/, smoke() and bench.py's CPU legs.

Every function cites the reference lines it follows (paths relative to the reference root).  The restatement
is written from the state_dict key names, so it also pins the drop-in's parameter naming.

Generalisation: the shipped reference hard-codes `+128/+64/+256` input channels
(nets/yolo_mul.py:361,364,376) and `(40,40)/(80,80)` interpolate targets (:426,:433), which are only right for
phi='n' at 640x640.  Here those five constants are written as `8*bc`, `4*bc`, `C5`, `feat2.shape[-2:]`,
`feat1.shape[-2:]`; for phi='n'@640 that is the identity.  oracle/make_golden.py applies the same five-constant
patch to the real reference classes to produce the golden vectors for the other configurations.

Pinned by tests/test_oracle_cpu.py against tests/golden/*.npz, which were produced by importing the real
reference from /root/reference (script: oracle/make_golden.py).
"""
import math

import torch
import torch.nn.functional as F

# nets/yolo_mul.py:331-337
DEPTH = {'n': 0.33, 's': 0.33, 'm': 0.67, 'l': 1.00, 'x': 1.00}
WIDTH = {'n': 0.25, 's': 0.50, 'm': 0.75, 'l': 1.00, 'x': 1.25}
DEEP = {'n': 1.00, 's': 1.00, 'm': 0.75, 'l': 0.50, 'x': 0.50}
EPS_YOLO_CONV = 1e-3   # nets/yolo_mul.py:197
EPS_DEFAULT = 1e-5     # torch default: nets/yolo_mul.py:109,141,146,149; nets/repghost.py:82,89,100,298


def dims(phi):
    bc = int(WIDTH[phi] * 64)
    depth = max(round(DEPTH[phi] * 3), 1)
    return bc, depth, bc * 4, bc * 8, int(bc * 16 * DEEP[phi])


def _bn(x, sd, p, eps):
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'], sd[p + '.bias'],
                        False, 0.0, eps)


def _silu(x):
    return x * torch.sigmoid(x)  # nets/yolo_mul.py:183-187


def conv_bn_silu(x, sd, p, stride=1, eps=EPS_YOLO_CONV):
    """Conv: conv(bias=False, pad=k//2) + BN + SiLU.  nets/yolo_mul.py:190-201 (eps 1e-3), nets/repghost.py:291-302 (eps 1e-5)."""
    w = sd[p + '.conv.weight']
    return _silu(_bn(F.conv2d(x, w, None, stride, w.shape[-1] // 2), sd, p + '.bn', eps))


def stem(x, sd, p):
    """Conv_maxpool: nets/yolo_mul.py:104-115."""
    x = F.relu(_bn(F.conv2d(x, sd[p + '.conv.0.weight'], None, 1, 1), sd, p + '.conv.1', EPS_DEFAULT))
    return F.max_pool2d(x, 3, 2, 1)


def shuffle_unit(x, sd, p):
    """ShuffleNetV2 stride 1: nets/yolo_mul.py:138-168."""
    x1, x2 = x.chunk(2, dim=1)
    b = p + '.branch2'
    y = F.relu(_bn(F.conv2d(x2, sd[b + '.0.weight']), sd, b + '.1', EPS_DEFAULT))
    y = _bn(F.conv2d(y, sd[b + '.3.weight'], sd[b + '.3.bias'], 1, 1, groups=y.shape[1]), sd, b + '.4', EPS_DEFAULT)
    y = F.relu(_bn(F.conv2d(y, sd[b + '.5.weight']), sd, b + '.6', EPS_DEFAULT))
    out = torch.cat((x1, y), 1)
    n, c, h, w = out.shape
    return out.view(n, 2, c // 2, h, w).permute(0, 2, 1, 3, 4).contiguous().view(n, c, h, w)


def cbam(x, sd, p):
    """CBAM = channel attention then spatial attention: nets/yolo_mul.py:56-102."""
    fc1, fc2 = sd[p + '.channelattention.fc1.weight'], sd[p + '.channelattention.fc2.weight']
    mlp = lambda v: F.conv2d(F.relu(F.conv2d(v, fc1)), fc2)
    avg = F.adaptive_avg_pool2d(x, 1)
    mx = F.adaptive_max_pool2d(x, 1)
    x = x * torch.sigmoid(mlp(avg) + mlp(mx))
    s = torch.cat([torch.mean(x, dim=1, keepdim=True), torch.max(x, dim=1, keepdim=True)[0]], dim=1)
    w7 = sd[p + '.spatialattention.conv1.weight']
    return x * torch.sigmoid(F.conv2d(s, w7, padding=w7.shape[-1] // 2))


def sppf_cbam(x, sd, p):
    """SPPF_CBAM: nets/yolo_mul.py:10-32."""
    x = cbam(conv_bn_silu(x, sd, p + '.cv1'), sd, p + '.cbam1')
    y1 = cbam(F.max_pool2d(x, 5, 1, 2), sd, p + '.cbam2')
    y2 = cbam(F.max_pool2d(y1, 5, 1, 2), sd, p + '.cbam3')
    y3 = cbam(F.max_pool2d(y2, 5, 1, 2), sd, p + '.cbam4')
    return conv_bn_silu(torch.cat((x, y1, y2, y3), 1), sd, p + '.cv2')


def backbone(x, sd, p):
    """Backbone.forward: nets/yolo_mul.py:252-308."""
    x = stem(x, sd, p + '.stem')
    x = shuffle_unit(conv_bn_silu(x, sd, p + '.dark2.0', 2), sd, p + '.dark2.1')
    feat1 = shuffle_unit(conv_bn_silu(x, sd, p + '.dark3.0', 2), sd, p + '.dark3.1')
    feat2 = shuffle_unit(conv_bn_silu(feat1, sd, p + '.dark4.0', 2), sd, p + '.dark4.1')
    x = shuffle_unit(conv_bn_silu(feat2, sd, p + '.dark5.0', 2), sd, p + '.dark5.1')
    return feat1, feat2, sppf_cbam(x, sd, p + '.dark5.2')


def repghost_module(x, sd, p, act):
    """RepGhostModule (reparam_bn=True, not deployed): nets/repghost.py:70-115."""
    x1 = _bn(F.conv2d(x, sd[p + '.primary_conv.0.weight']), sd, p + '.primary_conv.1', EPS_DEFAULT)
    if act:
        x1 = F.silu(x1)
    x2 = _bn(F.conv2d(x1, sd[p + '.cheap_operation.0.weight'], None, 1, 1, groups=x1.shape[1]), sd,
             p + '.cheap_operation.1', EPS_DEFAULT)
    x2 = x2 + _bn(x1, sd, p + '.fusion_bn.0', EPS_DEFAULT)
    return F.silu(x2) if act else x2


def repghost_bottleneck(x, sd, p):
    """RepGhostBottleneck, stride 1, no SE, identity shortcut: nets/repghost.py:263-279."""
    return repghost_module(repghost_module(x, sd, p + '.ghost1', True), sd, p + '.ghost2', False) + x


def c2f_repghost(x, sd, p, n):
    """C2f_repghost: nets/repghost.py:308-320 (local Conv: BN eps 1e-5)."""
    y = conv_bn_silu(x, sd, p + '.cv1', eps=EPS_DEFAULT)
    c = y.shape[1] // 2
    ys = list(y.split((c, c), 1))
    for i in range(n):
        ys.append(repghost_bottleneck(ys[-1], sd, '%s.m.%d' % (p, i)))
    return conv_bn_silu(torch.cat(ys, 1), sd, p + '.cv2', eps=EPS_DEFAULT)


def bifpn_concat(xs, sd):
    """Concat_BiFPN: nets/yolo_mul.py:36-51."""
    w = sd['bi_fpn.w']
    w = w / (torch.sum(w, dim=0) + 0.0001)
    return torch.cat([w[0] * xs[0], w[1] * xs[1], w[2] * xs[2]], 1)


def head_branch(x, sd, p):
    """cv2[i] / cv3[i]: Conv3x3, Conv3x3, Conv2d 1x1 with bias.  nets/yolo_mul.py:388-391."""
    x = conv_bn_silu(conv_bn_silu(x, sd, p + '.0'), sd, p + '.1')
    return F.conv2d(x, sd[p + '.2.weight'], sd[p + '.2.bias'])


def make_anchors(shapes, strides, offset=0.5):
    """utils/utils_bbox.py:16-28 (always fp32 here)."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        sy, sx = torch.meshgrid(sy, sx, indexing='ij')
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), dtype=torch.float32))
    return torch.cat(pts), torch.cat(st)


def dfl(box, sd):
    """DFL: nets/yolo_mul.py:312-322."""
    b, _, a = box.shape
    return F.conv2d(box.view(b, 4, 16, a).transpose(2, 1).softmax(1), sd['dfl.conv.weight']).view(b, 4, a)


@torch.no_grad()
def yolo_forward(sd, phi, rgb, nir, num_classes):
    """YoloBody.forward: nets/yolo_mul.py:397-462 -> (dbox, cls, x, anchors(2,A), strides(1,A))."""
    sd = {k: v.detach().float().cpu() for k, v in sd.items() if v.dtype.is_floating_point}
    _, depth, _, _, _ = dims(phi)
    f1r, f2r, f3r = backbone(rgb, sd, 'backbone_rgb')
    f1n, f2n, f3n = backbone(nir, sd, 'backbone_nir')
    f1r, f1n = cbam(f1r, sd, 'cbam_rgb_feat1'), cbam(f1n, sd, 'cbam_nir_feat1')
    f2r, f2n = cbam(f2r, sd, 'cbam_rgb_feat2'), cbam(f2n, sd, 'cbam_nir_feat2')
    f3r, f3n = cbam(f3r, sd, 'cbam_rgb_feat3'), cbam(f3n, sd, 'cbam_nir_feat3')
    feat3 = f3r + f3n
    p5_up = F.interpolate(feat3, size=f2r.shape[-2:], mode='bilinear', align_corners=True)
    p4 = c2f_repghost(bifpn_concat([p5_up, f2r, f2n], sd), sd, 'conv3_for_upsample1', depth)
    p4_up = F.interpolate(p4, size=f1r.shape[-2:], mode='bilinear', align_corners=True)
    p3 = c2f_repghost(bifpn_concat([p4_up, f1r, f1n], sd), sd, 'conv3_for_upsample2', depth)
    p4 = c2f_repghost(torch.cat([conv_bn_silu(p3, sd, 'down_sample1', 2), p4], 1), sd, 'conv3_for_downsample1', depth)
    p5 = c2f_repghost(bifpn_concat([conv_bn_silu(p4, sd, 'down_sample2', 2), f3r, f3n], sd), sd,
                      'conv3_for_downsample2', depth)
    x = []
    for i, p in enumerate((p3, p4, p5)):
        x.append(torch.cat((head_branch(p, sd, 'cv2.%d' % i), head_branch(p, sd, 'cv3.%d' % i)), 1))
    anchors, strides = make_anchors([t.shape[-2:] for t in x], (8.0, 16.0, 32.0))
    b = rgb.shape[0]
    no = num_classes + 64
    box, cls = torch.cat([xi.view(b, no, -1) for xi in x], 2).split((64, num_classes), 1)
    return dfl(box, sd), cls, x, anchors.transpose(0, 1), strides.transpose(0, 1)


def dist2bbox_xywh(distance, anchor_points):
    """utils/utils_bbox.py:30-40 with xywh=True, dim=1."""
    lt, rb = torch.split(distance, 2, 1)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    return torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1)


@torch.no_grad()
def decode_box(outputs, input_shape):
    """DecodeBox.decode_box: utils/utils_bbox.py:49-58 -> (B, A, 4+nc) fp32."""
    dbox, cls, _, anchors, strides = outputs
    dbox = dist2bbox_xywh(dbox, anchors.unsqueeze(0)) * strides
    y = torch.cat((dbox, cls.sigmoid()), 1).permute(0, 2, 1).contiguous()
    y[:, :, :4] = y[:, :, :4] / torch.tensor([input_shape[1], input_shape[0], input_shape[1], input_shape[0]],
                                             dtype=torch.float32)
    return y


# --------------------------------------------------------------------------------------------------------------
# deterministic synthetic weights / inputs (numpy MT19937: stable across numpy versions and platforms)
# --------------------------------------------------------------------------------------------------------------
def synth_state_dict(template_sd, seed, mode='stress'):
    """Fill a state_dict (key -> tensor, shapes taken from `template_sd`) deterministically.

    mode 'stress' (SURVEY F4): kaiming-normal conv weights, BN gamma U(0.8,1.2), beta N(0,0.1),
    running_mean N(0,0.1), running_var U(0.75,1.25), conv biases N(0,0.1), bi_fpn.w U(0.5,1.5).
    mode 'default': the reference constructor's init (nets/yolo_training.py:480-498): conv weights N(0,0.02),
    BN gamma N(1,0.02), beta 0, running stats 0/1; conv biases U(-1/sqrt(fan_in), 1/sqrt(fan_in)); bi_fpn.w = 1.
    dfl.conv.weight is arange(16) in both (nets/yolo_mul.py:315-317).
    """
    import numpy as np
    rs = np.random.RandomState(seed)
    out = {}
    gain = 1.0
    if ':' in mode:  # 'stress:0.7' scales the kaiming std (deeper variants need < 1 to keep activations O(1))
        mode, g = mode.split(':')
        gain = float(g)
    for k in sorted(template_sd.keys()):
        v = template_sd[k]
        shape = tuple(v.shape)
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros(shape, dtype=torch.int64)
            continue
        if k == 'dfl.conv.weight':
            out[k] = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
            continue
        stress = mode == 'stress'
        if k == 'bi_fpn.w':
            a = rs.uniform(0.5, 1.5, shape) if stress else np.ones(shape)
        elif k.endswith('running_mean'):
            a = rs.normal(0, 0.1, shape) if stress else np.zeros(shape)
        elif k.endswith('running_var'):
            a = rs.uniform(0.75, 1.25, shape) if stress else np.ones(shape)
        elif len(shape) == 4:  # conv weight
            fan_in = shape[1] * shape[2] * shape[3]
            a = rs.normal(0, gain * math.sqrt(2.0 / fan_in), shape) if stress else rs.normal(0, 0.02, shape)
        elif k.endswith('.weight'):  # BN gamma (1-D)
            a = rs.uniform(0.8, 1.2, shape) if stress else rs.normal(1.0, 0.02, shape)
        elif k.endswith('.bias'):
            is_bn = (k[:-5] + '.running_mean') in template_sd
            if is_bn:
                a = rs.normal(0, 0.1, shape) if stress else np.zeros(shape)
            else:  # conv bias
                wshape = tuple(template_sd[k[:-5] + '.weight'].shape)
                bound = 1.0 / math.sqrt(wshape[1] * wshape[2] * wshape[3])
                a = rs.normal(0, 0.1, shape) if stress else rs.uniform(-bound, bound, shape)
        else:
            raise KeyError('synth_state_dict: unhandled key %s' % k)
        out[k] = torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape).copy())
    return out


def synth_inputs(b, h, w, seed):
    """RGB and depth batches in [0,1): the depth image is one plane replicated x3 (utils/utils.py:14-19)."""
    import numpy as np
    rs = np.random.RandomState(seed)
    rgb = torch.from_numpy(rs.uniform(0, 1, (b, 3, h, w)).astype(np.float32))
    d = torch.from_numpy(rs.uniform(0, 1, (b, 1, h, w)).astype(np.float32))
    return rgb, d.expand(b, 3, h, w).contiguous()
