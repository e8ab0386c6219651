"""``nerv.models`` subset: conv/deconv block builders used by the SAVi CNNs.

Semantics inferred from reference call sites savi.py:231-239 (encoder, k=5,
stride 1|2 must map 64->64 / 128->64) and savi.py:269-284 (decoder, stride-2
blocks must double the size).
"""
from torch import nn


def _norm2d(norm, channels):
    if norm in ('', None):
        return nn.Identity()
    if norm == 'bn':
        return nn.BatchNorm2d(channels)
    if norm == 'in':
        return nn.InstanceNorm2d(channels)
    if norm == 'gn':
        return nn.GroupNorm(max(1, channels // 16), channels)
    raise ValueError(f'unknown norm {norm!r}')


def _act(act):
    if act in ('', None):
        return nn.Identity()
    table = {'relu': nn.ReLU, 'leakyrelu': nn.LeakyReLU, 'tanh': nn.Tanh,
             'gelu': nn.GELU, 'sigmoid': nn.Sigmoid}
    return table[act.lower()]()


def conv_norm_act(in_channels, out_channels, kernel_size, stride=1,
                  dilation=1, groups=1, norm='bn', act='relu', dim='2d'):
    assert dim == '2d'
    conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride,
                     padding=kernel_size // 2, dilation=dilation,
                     groups=groups, bias=norm in ('', None))
    return nn.Sequential(conv, _norm2d(norm, out_channels), _act(act))


def deconv_norm_act(in_channels, out_channels, kernel_size, stride=1,
                    dilation=1, groups=1, norm='bn', act='relu', dim='2d'):
    assert dim == '2d'
    deconv = nn.ConvTranspose2d(
        in_channels, out_channels, kernel_size, stride=stride,
        padding=kernel_size // 2, output_padding=stride - 1,
        dilation=dilation, groups=groups, bias=norm in ('', None))
    return nn.Sequential(deconv, _norm2d(norm, out_channels), _act(act))


def deconv_out_shape(in_size, stride, padding, kernel_size, out_padding,
                     dilation=1):
    return (in_size - 1) * stride - 2 * padding + \
        dilation * (kernel_size - 1) + out_padding + 1
