"""Only the piece of the reference's nets/yolo_training.py that the inference path imports: weights_init
(reference nets/yolo_training.py:480-498, called from YoloBody.__init__ at nets/yolo_mul.py:393-394).
The loss, assigner, EMA and LR schedules are training-only and out of scope."""
import torch.nn as nn

_INITS = {
    'normal': lambda w, gain: nn.init.normal_(w, 0.0, gain),
    'xavier': lambda w, gain: nn.init.xavier_normal_(w, gain=gain),
    'kaiming': lambda w, gain: nn.init.kaiming_normal_(w, a=0, mode='fan_in'),
    'orthogonal': lambda w, gain: nn.init.orthogonal_(w, gain=gain),
}


def weights_init(net, init_type='normal', init_gain=0.02):
    """Every module whose class name contains 'Conv' and that owns a `weight` gets `init_type`; every
    BatchNorm2d gets weight ~ N(1, 0.02), bias = 0.  Same selection rule and distributions as the reference."""
    if init_type not in _INITS:
        raise NotImplementedError('initialization method [%s] is not implemented' % init_type)
    fill = _INITS[init_type]
    for m in net.modules():
        name = type(m).__name__
        if 'Conv' in name and getattr(m, 'weight', None) is not None:
            fill(m.weight.data, init_gain)
        elif 'BatchNorm2d' in name:
            nn.init.normal_(m.weight.data, 1.0, 0.02)
            nn.init.constant_(m.bias.data, 0.0)
    print('initialize network with %s type' % init_type)
