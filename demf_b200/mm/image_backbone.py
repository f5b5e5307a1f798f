"""The frozen image feature extractor in front of the Deformable-DETR encoder: mmdet 2.14 `ResNet`
and `ChannelMapper`, as the reference builds them from configs/deformdetr/imvotenet_image.py:3-20
(inherited by configs/demf/demf_votenet.py:1-5) in demf/modeling/detectors/demfnet.py:42-46 and runs
them under `torch.no_grad()` in `extract_img_feat` (demfnet.py:124-132).

Same registry names, constructor arguments and state-dict keys as upstream (`conv1`, `bn1`,
`layerN.M.{conv1,bn1,conv2,bn2,conv3,bn3,downsample.0,downsample.1}`; `convs.N.{conv,gn}`,
`extra_convs.N.{conv,gn}`), so released checkpoints load. The branch is frozen and its arithmetic is
plain dense convolution: the convolutions are library kernels (cuDNN, channels-last when on CUDA; in frozen
inference every conv -> BN -> (add) -> ReLU group is ONE fused cuDNN convolution with the BatchNorm folded into
its weight and bias) -- SURVEY.md section 8(f) row 2 names them "cuDNN"; the sm_100a kernels of this repository
start at the encoder's deformable attention.
"""
import torch
import torch.nn as nn

from .bricks import BaseModule, ConvModule, build_norm_layer
from .registry import BACKBONES, NECKS


def _fold_bn(conv, bn):
    """(W', b') with bn(conv(x)) == conv(x; W') + b' for a BatchNorm in eval mode; cached on the modules until one
    of the five source tensors changes (the branch is frozen: computed once)."""
    src = (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple((t.data_ptr(), t._version) for t in src)
    hit = conv.__dict__.get('_demf_folded')
    if hit is None or hit[0] != key:
        with torch.no_grad():
            scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
            w = (conv.weight * scale.view(-1, 1, 1, 1)).contiguous(memory_format=torch.channels_last)
            b = (bn.bias - bn.running_mean * scale).contiguous()
        hit = (key, w, b)
        conv.__dict__['_demf_folded'] = hit
    return hit[1], hit[2]


def _frozen_inference(x, *bns):
    """The fused library path applies: CUDA, no autograd, every BatchNorm using its running statistics."""
    return x.is_cuda and not torch.is_grad_enabled() and all(
        isinstance(bn, nn.modules.batchnorm._BatchNorm) and not bn.training and bn.track_running_stats for bn in bns)


def _conv_bias_relu(x, conv, bn, residual=None):
    """relu(bn(conv(x)) [+ residual]) as ONE cuDNN fused convolution (BatchNorm folded into weight and bias):
    the activation tensor is written once instead of conv -> BN -> (add) -> ReLU, each a full pass."""
    w, b = _fold_bn(conv, bn)
    if residual is None:
        return torch.cudnn_convolution_relu(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    return torch.cudnn_convolution_add_relu(x, w, residual, 1.0, b, conv.stride, conv.padding, conv.dilation,
                                            conv.groups)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style='pytorch',
                 norm_cfg=dict(type='BN')):
        super().__init__()
        assert style in ('pytorch', 'caffe')
        # 'pytorch': the stride sits on the 3x3 convolution, 'caffe': on the first 1x1
        s1, s2 = (1, stride) if style == 'pytorch' else (stride, 1)
        self.conv1 = nn.Conv2d(inplanes, planes, 1, stride=s1, bias=False)
        self.add_module('bn1', build_norm_layer(norm_cfg, planes)[1])
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=s2, padding=dilation, dilation=dilation,
                               bias=False)
        self.add_module('bn2', build_norm_layer(norm_cfg, planes)[1])
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.add_module('bn3', build_norm_layer(norm_cfg, planes * self.expansion)[1])
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    fused_inference = True

    def forward(self, x):
        if self.fused_inference and _frozen_inference(x, self.bn1, self.bn2, self.bn3) and (
                self.downsample is None or _frozen_inference(x, self.downsample[1])):
            if self.downsample is None:
                identity = x
            else:
                wd, bd = _fold_bn(self.downsample[0], self.downsample[1])
                identity = nn.functional.conv2d(x, wd, bd, self.downsample[0].stride)
            out = _conv_bias_relu(x, self.conv1, self.bn1)
            out = _conv_bias_relu(out, self.conv2, self.bn2)
            return _conv_bias_relu(out, self.conv3, self.bn3, residual=identity)
        identity = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        out += identity
        return self.relu(out)


def _res_layer(inplanes, planes, blocks, stride, style, norm_cfg):
    downsample = None
    if stride != 1 or inplanes != planes * Bottleneck.expansion:
        downsample = nn.Sequential(
            nn.Conv2d(inplanes, planes * Bottleneck.expansion, 1, stride=stride, bias=False),
            build_norm_layer(norm_cfg, planes * Bottleneck.expansion)[1])
    layers = [Bottleneck(inplanes, planes, stride, 1, downsample, style, norm_cfg)]
    inplanes = planes * Bottleneck.expansion
    for _ in range(1, blocks):
        layers.append(Bottleneck(inplanes, planes, 1, 1, None, style, norm_cfg))
    return nn.Sequential(*layers)


@BACKBONES.register_module()
class ResNet(BaseModule):
    """mmdet ResNet (bottleneck depths): forward(img (B,3,H,W)) -> tuple of the `out_indices` stages."""

    arch_settings = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}

    def __init__(self, depth, in_channels=3, stem_channels=None, base_channels=64, num_stages=4,
                 strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1), out_indices=(0, 1, 2, 3),
                 style='pytorch', deep_stem=False, avg_down=False, frozen_stages=-1, conv_cfg=None,
                 norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, dcn=None,
                 stage_with_dcn=(False, False, False, False), plugins=None, with_cp=False,
                 zero_init_residual=True, pretrained=None, init_cfg=None):
        super().__init__(init_cfg)
        if depth not in self.arch_settings:
            raise KeyError(f'invalid depth {depth} for resnet (bottleneck depths 50/101/152 are built here)')
        if deep_stem or avg_down or dcn is not None or plugins is not None or conv_cfg is not None \
                or any(d != 1 for d in dilations):
            raise NotImplementedError('only the plain ResNet of configs/deformdetr/imvotenet_image.py:3-12')
        assert 1 <= num_stages <= 4 and max(out_indices) < num_stages
        self.depth, self.num_stages = depth, num_stages
        self.out_indices, self.frozen_stages = tuple(out_indices), frozen_stages
        self.norm_eval, self.zero_init_residual = norm_eval, zero_init_residual
        stem_channels = stem_channels or base_channels
        self.conv1 = nn.Conv2d(in_channels, stem_channels, 7, stride=2, padding=3, bias=False)
        self.add_module('bn1', build_norm_layer(norm_cfg, stem_channels)[1])
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.res_layers = []
        inplanes = stem_channels
        for i, blocks in enumerate(self.arch_settings[depth][:num_stages]):
            planes = base_channels * 2 ** i
            self.add_module(f'layer{i + 1}', _res_layer(inplanes, planes, blocks, strides[i], style, norm_cfg))
            inplanes = planes * Bottleneck.expansion
            self.res_layers.append(f'layer{i + 1}')
        self.feat_dim = inplanes
        self._freeze_stages()

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            for m in (self.conv1, self.bn1):
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            m = getattr(self, f'layer{i}')
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, (nn.modules.batchnorm._BatchNorm, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        if self.zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.constant_(m.bn3.weight, 0)
        self._is_init = True

    def forward(self, x):
        if x.is_cuda and x.dim() == 4:   # NHWC is what the library's tensor-core convolutions want
            x = x.contiguous(memory_format=torch.channels_last)
        if Bottleneck.fused_inference and _frozen_inference(x, self.bn1):
            x = self.maxpool(_conv_bias_relu(x, self.conv1, self.bn1))
        else:
            x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        outs = []
        for i, name in enumerate(self.res_layers):
            x = getattr(self, name)(x)
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
        return self


@NECKS.register_module()
class ChannelMapper(BaseModule):
    """mmdet ChannelMapper: one k x k ConvModule per input level to `out_channels`, plus stride-2 3x3
    ConvModules producing `num_outs - len(in_channels)` coarser levels (the first from the LAST INPUT,
    the following ones from the previous extra output)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type='ReLU'), num_outs=None,
                 init_cfg=dict(type='Xavier', layer='Conv2d', distribution='uniform')):
        super().__init__(init_cfg)
        assert isinstance(in_channels, (list, tuple))
        self.extra_convs = None
        if num_outs is None:
            num_outs = len(in_channels)
        self.convs = nn.ModuleList()
        for c in in_channels:
            self.convs.append(ConvModule(c, out_channels, kernel_size, padding=(kernel_size - 1) // 2,
                                         conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg))
        if num_outs > len(in_channels):
            self.extra_convs = nn.ModuleList()
            for i in range(len(in_channels), num_outs):
                c = in_channels[-1] if i == len(in_channels) else out_channels
                self.extra_convs.append(ConvModule(c, out_channels, 3, stride=2, padding=1,
                                                   conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg))

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        self._is_init = True

    def forward(self, inputs):
        assert len(inputs) == len(self.convs)
        outs = [self.convs[i](inputs[i]) for i in range(len(inputs))]
        if self.extra_convs:
            for i, conv in enumerate(self.extra_convs):
                outs.append(conv(inputs[-1] if i == 0 else outs[-1]))
        return tuple(outs)
