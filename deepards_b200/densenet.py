"""1-D DenseNet backbones on the B200 backend -- drop-in for deepards/models/densenet.py.

Same factory (`densenet18(pretrained=False, progress=True, with_fft=False, only_fft=False, fft_real_only=False)`,
densenet.py:96-98,223-231), attributes (`features`, `avgpool`, `n_out_filters`, `network_name`, `drop_rate`,
`forward_no_pool`, `conv_info()`) and `state_dict` keys (`features.conv0`, `features.norm0`,
`features.denseblockK.denselayerJ.{norm1,conv1,norm2,conv2}`, `features.transitionK.{norm,conv}`,
`features.norm5`; no BatchNorm buffers: track_running_stats=False, densenet.py:107).

`features` is callable and returns the norm5 output (N, 128, 7) through autograd, which is what GradCAM hooks
(gradcam.py:45-47).  The channel concatenation of a dense block (densenet.py:40) never happens as a copy: each
layer's conv writes its 32 new channels straight into the block's channels-last buffer.
"""
import math
from collections import OrderedDict

import torch.nn as nn
import torch.nn.functional as F

from . import autograd as _ag
from .resnet import _Picklable


class _DenseLayer(_Picklable, nn.Module):
    """BN-ReLU-conv1x1-BN-ReLU-conv3(-dropout): parameter container (densenet.py:18-43)."""
    num_layers = 2

    def __init__(self, num_input_features, growth_rate, bn_size, drop_rate):
        super(_DenseLayer, self).__init__()
        mid = bn_size * growth_rate
        self.norm1 = nn.BatchNorm1d(num_input_features, track_running_stats=False)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv1 = nn.Conv1d(num_input_features, mid, kernel_size=1, stride=1, bias=False)
        self.norm2 = nn.BatchNorm1d(mid, track_running_stats=False)
        self.relu2 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv1d(mid, growth_rate, kernel_size=3, stride=1, padding=1, bias=False)
        self.drop_rate = float(drop_rate)

    def conv_info(self):
        return [1, 3], [1, 1], [0, 1]


class _DenseBlock(_Picklable, nn.Module):
    def __init__(self, num_layers, num_input_features, bn_size, growth_rate, drop_rate):
        super(_DenseBlock, self).__init__()
        self.kernel_sizes, self.strides, self.paddings = [], [], []
        self.track_running_stats = False
        for i in range(num_layers):
            layer = _DenseLayer(num_input_features + i * growth_rate, growth_rate, bn_size, drop_rate)
            ks, st, pd = layer.conv_info()
            self.kernel_sizes += ks
            self.strides += st
            self.paddings += pd
            self.add_module('denselayer%d' % (i + 1), layer)
        self.num_layers = _DenseLayer.num_layers * num_layers

    def conv_info(self):
        return self.kernel_sizes, self.strides, self.paddings


class _Transition(_Picklable, nn.Module):
    """BN-ReLU-conv1x1-AvgPool(2,2): parameter container (densenet.py:68-80)."""
    num_layers = 1

    def __init__(self, num_input_features, num_output_features):
        super(_Transition, self).__init__()
        self.norm = nn.BatchNorm1d(num_input_features, track_running_stats=False)
        self.relu = nn.ReLU(inplace=True)
        self.conv = nn.Conv1d(num_input_features, num_output_features, kernel_size=1, stride=1, bias=False)
        self.pool = nn.AvgPool1d(kernel_size=2, stride=2)

    def conv_info(self):
        return [1, 2], [1, 2], [0, 0]


class _Features(_Picklable, nn.Module):
    """The `features` stack.  Calling it runs conv0 ... norm5 on the B200 plan and returns (N, C, 7)."""

    def __init__(self):
        super(_Features, self).__init__()
        self.network_name = 'densenet'
        self.precision = None

    def _drop_key(self):
        return tuple(float(m.drop_rate) for m in self.modules() if isinstance(m, _DenseLayer) and m.drop_rate > 0)

    def forward(self, x):
        if x.dim() != 3 or x.shape[1] != 1 or x.shape[2] != 224:
            raise RuntimeError("deepards_b200 DenseNet expects (N, 1, 224), got %s" % (tuple(x.shape),))
        return _ag.run_plan(self, self, None, x, x.shape[0], "features", self._drop_key())


class DenseNet(_Picklable, nn.Module):
    def __init__(self, growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4, drop_rate=0.2,
                 num_classes=1000, with_fft=False, only_fft=False, fft_real_only=False):
        super(DenseNet, self).__init__()
        if with_fft or only_fft:
            raise NotImplementedError("with_fft / only_fft inputs are not implemented on the B200 backend")
        self.kernel_sizes, self.strides, self.paddings = [7, 3], [2, 2], [3, 1]
        self.n_layers = 0
        self.inplanes = num_init_features
        self.drop_rate = drop_rate
        feats = _Features()
        feats.add_module('conv0', nn.Conv1d(1, num_init_features, kernel_size=7, stride=2, padding=3, bias=False))
        feats.add_module('norm0', nn.BatchNorm1d(num_init_features, track_running_stats=False))
        feats.add_module('relu0', nn.ReLU(inplace=True))
        feats.add_module('pool0', nn.MaxPool1d(kernel_size=3, stride=2, padding=1))
        nf = num_init_features
        for i, n_layers in enumerate(block_config):
            block = _DenseBlock(n_layers, nf, bn_size, growth_rate, drop_rate)
            self._note(block)
            feats.add_module('denseblock%d' % (i + 1), block)
            nf += n_layers * growth_rate
            if i != len(block_config) - 1:
                trans = _Transition(nf, nf // 2)
                self._note(trans)
                self.n_layers += trans.num_layers  # the reference counts transitions twice (densenet.py:143)
                feats.add_module('transition%d' % (i + 1), trans)
                nf //= 2
        feats.add_module('norm5', nn.BatchNorm1d(nf, track_running_stats=False))
        self.features = feats
        for m in self.modules():
            if isinstance(m, nn.Conv1d):
                m.weight.data.normal_(0, math.sqrt(2.0 / (m.kernel_size[0] * m.out_channels)))
            elif isinstance(m, nn.BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.n_out_filters = nf
        self.avgpool = nn.AvgPool1d(7, stride=1)
        self.network_name = 'densenet'
        self.precision = None

    def _note(self, obj):
        ks, st, pd = obj.conv_info()
        self.n_layers += obj.num_layers
        self.kernel_sizes += ks
        self.strides += st
        self.paddings += pd

    def conv_info(self):
        return self.kernel_sizes, self.strides, self.paddings

    def __setattr__(self, name, value):
        # keep `features` in step when the owner changes precision / name on the backbone
        super(DenseNet, self).__setattr__(name, value)
        if name in ('precision', 'network_name') and 'features' in self._modules:
            object.__setattr__(self._modules['features'], name, value)

    def forward(self, x):
        """x: (N, 1, 224) -> (N, n_out_filters): relu(features) -> AvgPool1d(7) -> flatten (densenet.py:179-189),
        as ONE plan (norm5 + ReLU + pooling fused into the tail kernels)."""
        if x.dim() != 3 or x.shape[1] != 1 or x.shape[2] != 224:
            raise RuntimeError("deepards_b200 DenseNet expects (N, 1, 224), got %s" % (tuple(x.shape),))
        return _ag.run_plan(self, self, None, x, x.shape[0], "backbone", self.features._drop_key())

    def forward_no_pool(self, x):
        """relu(features(x)) (densenet.py:191-193), used by ProtoPNet."""
        return F.relu(self.features(x))


def _densenet(arch, growth_rate, block_config, num_init_features, pretrained, progress, **kwargs):
    if pretrained:
        raise NotImplementedError("pretrained ImageNet weights do not exist for the 1-D networks")
    model = DenseNet(growth_rate, block_config, num_init_features, **kwargs)
    model.network_name = arch
    return model


def densenet18(pretrained=False, progress=True, **kwargs):
    """1-D DenseNet-18 = growth 32, blocks (2,2,2,2), 64 stem features (densenet.py:223-231)."""
    return _densenet('densenet18', 32, (2, 2, 2, 2), 64, pretrained, progress, **kwargs)


def densenet121(pretrained=False, progress=True, **kwargs):
    """1-D DenseNet-121 (densenet.py:234-242)."""
    return _densenet('densenet121', 32, (6, 12, 24, 16), 64, pretrained, progress, **kwargs)
