"""ORACLE (test infrastructure only -- never imported by the product path).

CPU fp32 restatement of the reference's DeepLabv3+ (ResNet-50 v1c) 'val' forward against a reference-format state_dict:
ResNet deep stem (src/models/backbones/seg/resnet.py:67-80) + torchvision Bottleneck layers (restated in fcos_oracle),
Deeplabv3PlusHead.forward (src/models/heads/seg/deeplabv3plus_head.py:56-68) with ASPP / image pool
(deeplabv3_head.py:15-75), DepthwiseSeparableConvModule (bricks/depthwise_separable_conv_module.py:96-99),
BaseSegHead.classify (base_seg_head.py:32-37; Dropout2d is the identity in eval mode) and the segmentor's
bilinear upsample + argmax (segmentors/encoder_decoder.py:131-133).  Pinned by tools/make_golden_deeplab.py.
NOTE: the reference really runs ResNet-50 at stride 32 here (its output_stride=8 branch is a no-op for resnet50).
"""
import torch
import torch.nn.functional as F

from .fcos_oracle import RESNET50_BLOCKS, bottleneck

EPS = 1e-5


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'], sd[p + '.bias'], False, 0.0, EPS)


def conv_module(x, sd, p, stride=1, pad=0, dil=1, groups=1):
    """ConvModule: conv (no bias) -> BN -> ReLU (src/models/bricks/conv_module.py:201-214)."""
    return F.relu(_bn(F.conv2d(x, sd[p + '.conv.weight'], None, stride, pad, dil, groups), sd, p + '.bn'))


def ds_conv(x, sd, p, dil=1):
    """DepthwiseSeparableConvModule 3x3 (padding == dilation)."""
    x = conv_module(x, sd, p + '.depthwise_conv', 1, dil, dil, groups=x.shape[1])
    return conv_module(x, sd, p + '.pointwise_conv')


def resnet50v1c(x, sd, prefix='backbone.', out_stages=(1, 4)):
    s = prefix + 'stem.'
    x = F.relu(_bn(F.conv2d(x, sd[s + '0.weight'], None, 2, 1), sd, s + '1'))
    x = F.relu(_bn(F.conv2d(x, sd[s + '3.weight'], None, 1, 1), sd, s + '4'))
    x = F.relu(_bn(F.conv2d(x, sd[s + '6.weight'], None, 1, 1), sd, s + '7'))
    x = F.max_pool2d(x, 3, 2, 1)
    outs = []
    for li, nb in enumerate(RESNET50_BLOCKS, start=1):
        for bi in range(nb):
            x = bottleneck(x, sd, f'{prefix}layer{li}.{bi}', 2 if (bi == 0 and li > 1) else 1)
        if li in out_stages:
            outs.append(x)
    return outs


def head(feats, sd, prefix='head.', dilations=(1, 12, 24, 36)):
    low, x = feats
    outs = [F.interpolate(conv_module(F.adaptive_avg_pool2d(x, 1), sd, prefix + 'proj.1'), size=x.shape[2:], mode='bilinear', align_corners=False)]
    for i, d in enumerate(dilations):
        outs.append(conv_module(x, sd, f'{prefix}aspp.{i}') if d == 1 else ds_conv(x, sd, f'{prefix}aspp.{i}', d))
    o = conv_module(torch.cat(outs, 1), sd, prefix + 'reduce', 1, 1)
    lo = conv_module(low, sd, prefix + 'low_proj')
    o = F.interpolate(o, size=lo.shape[2:], mode='bilinear', align_corners=False)
    o = torch.cat([o, lo], 1)
    o = ds_conv(o, sd, prefix + 'fuse.0')
    o = ds_conv(o, sd, prefix + 'fuse.1')
    return F.conv2d(o, sd[prefix + 'cls_seg.weight'], sd[prefix + 'cls_seg.bias'])


def forward(x, sd, out_hw=None):
    """Returns (feats, logits [B,19,H/4,W/4], labels int64 [B,H,W])."""
    with torch.no_grad():
        feats = resnet50v1c(x, sd)
        logits = head(feats, sd)
        up = F.interpolate(logits, size=out_hw or x.shape[2:], mode='bilinear', align_corners=False)
        return feats, logits, torch.argmax(up, dim=1)
