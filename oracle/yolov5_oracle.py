"""ORACLE (test infrastructure only -- never imported by the product path).

CPU fp32 restatement of the reference's YOLOv5-s inference forward, written against a reference-format
``state_dict`` with plain torch functional ops.  It does not import /root/reference, so it travels to the GPU
box.  Pinned against the reference's own modules by tools/make_golden.py -> tests/golden/*.npz and
tests/test_oracle_golden.py (the reference has no tests/KATs for this path, SURVEY.md §4, so the pin is
"outputs of the reference itself run in the build container").

Each function cites the reference code it restates (paths relative to /root/reference).
"""

import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # every BatchNorm2d on the path has eps forced to 1e-3:
#   src/models/backbones/det/yolov5_csp_darknet.py:99-101, src/models/necks/yolov5_neck.py:37-39

ANCHORS = [[[1.25000, 1.62500], [2.00000, 3.75000], [4.12500, 2.87500]],
           [[1.87500, 3.81250], [3.87500, 2.81250], [3.68750, 7.43750]],
           [[3.62500, 2.81250], [4.87500, 6.18750], [11.65625, 10.18750]]]  # src/models/yolov5.py:157-159
STRIDES = [8.0, 16.0, 32.0]  # src/models/detects/yolov5_detect.py:13


def conv_bn_silu(x, sd, prefix, stride=1, pad=0, act=True):
    """ConvModule.forward (src/models/bricks/conv_module.py:201-214, order conv->norm->act, conv bias off with
    norm :108-110) and Conv.forward (src/models/modules/yolo11_modules.py:27-39): SiLU(BN_eval(conv(x)))."""
    y = F.conv2d(x, sd[prefix + '.conv.weight'], None, stride, pad)
    y = F.batch_norm(y, sd[prefix + '.bn.running_mean'], sd[prefix + '.bn.running_var'], sd[prefix + '.bn.weight'],
                     sd[prefix + '.bn.bias'], False, 0.0, BN_EPS)
    return F.silu(y) if act else y


def csp_layer(x, sd, p, n, shortcut, names=('conv1', 'conv2', 'conv3', 'conv1', 'conv2')):
    """CSPLayer.forward (src/models/modules/yolo_modules.py:135-140) / C3.forward (yolo11_modules.py:216-217):
    conv3(cat(m(conv1(x)), conv2(x))); bottleneck = 1x1 -> 3x3 (+x) (yolo_modules.py:95-104, yolo11_modules.py:182-183)."""
    c1, c2, c3, b1, b2 = names
    x1 = conv_bn_silu(x, sd, f'{p}.{c1}')
    x2 = conv_bn_silu(x, sd, f'{p}.{c2}')
    for i in range(n):
        t = conv_bn_silu(x1, sd, f'{p}.m.{i}.{b1}')
        t = conv_bn_silu(t, sd, f'{p}.m.{i}.{b2}', 1, 1)
        x1 = t + x1 if shortcut else t
    return conv_bn_silu(torch.cat((x1, x2), 1), sd, f'{p}.{c3}')


def sppf(x, sd, p):
    """SPPF.forward int-kernel path (src/models/modules/yolo_modules.py:185-194): 1x1, 3 chained maxpool5/s1/p2, cat, 1x1."""
    x = conv_bn_silu(x, sd, f'{p}.conv1')
    y1 = F.max_pool2d(x, 5, 1, 2)
    y2 = F.max_pool2d(y1, 5, 1, 2)
    y3 = F.max_pool2d(y2, 5, 1, 2)
    return conv_bn_silu(torch.cat([x, y1, y2, y3], 1), sd, f'{p}.conv2')


def backbone(x, sd, prefix='backbone.', num_blocks=(1, 2, 3, 1)):
    """YOLOv5CSPDarknet.forward (src/models/backbones/det/yolov5_csp_darknet.py:83-91): stem k6 s2 p2 (:36-45);
    stage i = ConvModule k3 s2 p1 + CSPLayer(n, shortcut except stage 4 :69) [+ SPPF(5) in stage 4 :72-80]."""
    x = conv_bn_silu(x, sd, prefix + 'stem', 2, 2)
    outs = []
    for i in range(4):
        p = f'{prefix}stage{i + 1}'
        x = conv_bn_silu(x, sd, p + '.0', 2, 1)
        x = csp_layer(x, sd, p + '.1', num_blocks[i], shortcut=(i != 3))
        if i == 3:
            x = sppf(x, sd, p + '.2')
        if i >= 1:
            outs.append(x)
    return outs


_NECK_NAMES = ('cv1', 'cv2', 'cv3', 'cv1', 'cv2')


def neck(feats, sd, prefix='neck.'):
    """YOLOv5Neck.forward (src/models/necks/yolov5_neck.py:42-51) with UpsamplingModule / DownsamplingModule
    (src/models/modules/yolo11_modules.py:388-408): lateral 1x1 -> nearest x2 -> cat -> C3(n=1, no shortcut);
    3x3 s2 -> cat(lateral) -> C3."""
    x3, x4, x5 = feats

    def up(x, y, p):
        xc = conv_bn_silu(x, sd, f'{prefix}{p}.conv')
        u = F.interpolate(xc, scale_factor=2, mode='nearest')
        return csp_layer(torch.cat([u, y], 1), sd, f'{prefix}{p}.fuse', 1, False, _NECK_NAMES), xc

    def down(x, y, p):
        d = conv_bn_silu(x, sd, f'{prefix}{p}.down', 2, 1)
        return csp_layer(torch.cat([d, y], 1), sd, f'{prefix}{p}.fuse', 1, False, _NECK_NAMES)

    x4_up, x4_t = up(x5, x4, 'up_1')
    x3_up, x3_t = up(x4_up, x3, 'up_2')
    x4_down = down(x3_up, x3_t, 'down_1')
    x5_down = down(x4_down, x4_t, 'down_2')
    return [x3_up, x4_down, x5_down]


def detect(feats, sd, prefix='detect.', nc=80):
    """YOLOv5Detect.forward eval branch (src/models/detects/yolov5_detect.py:39-57) and _make_grid (:60-65).
    Returns (z [B, sum(3*ny*nx), 85], [raw_i [B,3,ny,nx,85]])."""
    no, na = nc + 5, 3
    z, raws = [], []
    for i, x in enumerate(feats):
        r = F.conv2d(x, sd[f'{prefix}m.{i}.weight'], sd[f'{prefix}m.{i}.bias'])
        bs, _, ny, nx = r.shape
        r = r.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        raws.append(r)
        yv, xv = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing='ij')
        grid = torch.stack((xv, yv), 2).expand((1, na, ny, nx, 2)).float()
        anchor_grid = (torch.tensor(ANCHORS[i]).float() * STRIDES[i]).view((1, na, 1, 1, 2)).expand((1, na, ny, nx, 2)).float()
        y = r.sigmoid()
        y[..., 0:2] = (y[..., 0:2] * 2 - 0.5 + grid) * STRIDES[i]
        y[..., 2:4] = (y[..., 2:4] * 2) ** 2 * anchor_grid
        z.append(y.view(bs, -1, no))
    return torch.cat(z, 1), raws


def forward(x, sd):
    """out, train_out = detect(neck(backbone(imgs)))  (src/models/yolov5.py:256)."""
    with torch.no_grad():
        return detect(neck(backbone(x, sd), sd), sd)

def input_transform(frames_u8_hwc, mean, std, reverse_channels=True):
    """ToTensor + Normalize of the reference (src/data/transforms/det_transforms.py:85-99 and :102-109): HWC -> CHW, channel
    reversal (BGR -> RGB), float32 / 255, then torchvision F.normalize = (x - mean[c]) / std[c] in fp32.
    frames: uint8 [B,H,W,3] (numpy or torch) -> float32 torch tensor [B,3,H,W].  Pinned by tests/golden/input_transform.npz."""
    f = torch.as_tensor(frames_u8_hwc)
    x = f.permute(0, 3, 1, 2)
    if reverse_channels:
        x = x.flip(1)
    x = x.contiguous().to(torch.float32).div_(255.0)
    m = torch.as_tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    sd = torch.as_tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return x.sub_(m).div_(sd)


def rel_err(a, b):
    """Parity metric of SURVEY.md §8(d): max|a-b| / (max|b| + 1e-12)."""
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-12))
