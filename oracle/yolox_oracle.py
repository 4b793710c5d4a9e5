"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the reference's YOLOX-s inference path (the composite that runs):
  backbone      CSPDarknet.forward        src/models/backbones/det/csp_darknet.py:83-91 (Focus yolo_modules.py:29-37,
                                          CSPLayer :135-140, SPPF tuple-kernel path :185-194)
  neck          YOLOXNeck.forward         src/models/necks/yolox_neck.py:49-74
  head          YOLOXHead.forward         src/models/heads/yolox_head.py:74-88 (1x1 stems with padding=1, :35)
  post_process  yolox_post_process        src/models/yolox.py:18-68, incl. the third-party torchvision.ops.batched_nms (:64;
                torchvision is not vendored, README.md:55 names 0.7.0): numel > 4000 on CPU tensors -> _batched_nms_vanilla
                (per-class nms, result ordered by score), else _batched_nms_coordinate_trick (boxes + class * (max + 1)); both
                restated with oracle.nms_oracle.greedy_nms (= torchvision.ops.nms CPU semantics).

Pinned against the reference's own modules / functions by tools/make_golden_yolox.py (tests/golden/yolox_*.npz).
Documented deviation: _batched_nms_vanilla orders its result with an UNSTABLE sort (scores[keep].sort(descending=True)); the
oracle uses a stable one (ties: lower index first).  Inputs without exactly equal kept scores behave identically."""
import numpy as np
import torch
import torch.nn.functional as F

from .nms_oracle import greedy_nms
from .yolov5_oracle import conv_bn_silu, csp_layer

STRIDES = [8, 16, 32]  # src/models/yolox.py:93
BN_EPS = 1e-3


def conv_bn_swish(x, sd, prefix, stride=1, pad=0):
    """ConvModule with act_cfg=dict(type='Swish') (the backbone's default, csp_darknet.py:28): bricks/swish.py computes
    x * sigmoid(x), which differs from F.silu in the last bit."""
    y = F.conv2d(x, sd[prefix + '.conv.weight'], None, stride, pad)
    y = F.batch_norm(y, sd[prefix + '.bn.running_mean'], sd[prefix + '.bn.running_var'], sd[prefix + '.bn.weight'],
                     sd[prefix + '.bn.bias'], False, 0.0, BN_EPS)
    return y * torch.sigmoid(y)


def csp_layer_swish(x, sd, p, n, shortcut):
    """yolo_modules.CSPLayer.forward (:135-140) with Swish ConvModules; bottleneck = 1x1 -> 3x3 (+x) (:95-104)."""
    x1 = conv_bn_swish(x, sd, f'{p}.conv1')
    x2 = conv_bn_swish(x, sd, f'{p}.conv2')
    for i in range(n):
        t = conv_bn_swish(conv_bn_swish(x1, sd, f'{p}.m.{i}.conv1'), sd, f'{p}.m.{i}.conv2', 1, 1)
        x1 = t + x1 if shortcut else t
    return conv_bn_swish(torch.cat((x1, x2), 1), sd, f'{p}.conv3')


def focus(x, sd, p):
    tl, tr = x[..., ::2, ::2], x[..., ::2, 1::2]
    bl, br = x[..., 1::2, ::2], x[..., 1::2, 1::2]
    return conv_bn_swish(torch.cat((tl, bl, tr, br), 1), sd, p + '.conv', 1, 1)


def sppf_parallel(x, sd, p, ks=(5, 9, 13)):
    x = conv_bn_swish(x, sd, f'{p}.conv1')
    return conv_bn_swish(torch.cat([x] + [F.max_pool2d(x, k, 1, k // 2) for k in ks], 1), sd, f'{p}.conv2')


def backbone(x, sd, prefix='backbone.', layers=(1, 3, 3, 1)):
    x = focus(x, sd, prefix + 'stem')
    outs = []
    for i in range(4):
        p = f'{prefix}stage{i + 1}'
        x = conv_bn_swish(x, sd, p + '.0', 2, 1)
        if i == 3:
            x = sppf_parallel(x, sd, p + '.1')
            x = csp_layer_swish(x, sd, p + '.2', layers[i], shortcut=False)
        else:
            x = csp_layer_swish(x, sd, p + '.1', layers[i], shortcut=True)
        if i >= 1:
            outs.append(x)
    return outs


def neck(feats, sd, prefix='neck.'):
    x2, x1, x0 = feats
    up = lambda t: F.interpolate(t, scale_factor=2, mode='nearest')
    fpn_out0 = conv_bn_silu(x0, sd, prefix + 'lateral_conv0')
    f_out0 = csp_layer(torch.cat([up(fpn_out0), x1], 1), sd, prefix + 'C3_p4', 1, False)
    fpn_out1 = conv_bn_silu(f_out0, sd, prefix + 'reduce_conv1')
    pan_out2 = csp_layer(torch.cat([up(fpn_out1), x2], 1), sd, prefix + 'C3_p3', 1, False)
    p_out1 = conv_bn_silu(pan_out2, sd, prefix + 'bu_conv2', 2, 1)
    pan_out1 = csp_layer(torch.cat([p_out1, fpn_out1], 1), sd, prefix + 'C3_n3', 1, False)
    p_out0 = conv_bn_silu(pan_out1, sd, prefix + 'bu_conv1', 2, 1)
    pan_out0 = csp_layer(torch.cat([p_out0, fpn_out0], 1), sd, prefix + 'C3_n4', 1, False)
    return [pan_out2, pan_out1, pan_out0]


def head(feats, sd, prefix='head.'):
    outs = []
    for k, x in enumerate(feats):
        xx = conv_bn_silu(x, sd, f'{prefix}stems.{k}', 1, 1)  # 1x1 conv with padding=1: the map grows by two pixels
        cf = conv_bn_silu(conv_bn_silu(xx, sd, f'{prefix}cls_convs.{k}.0', 1, 1), sd, f'{prefix}cls_convs.{k}.1', 1, 1)
        rf = conv_bn_silu(conv_bn_silu(xx, sd, f'{prefix}reg_convs.{k}.0', 1, 1), sd, f'{prefix}reg_convs.{k}.1', 1, 1)
        cls_o = F.conv2d(cf, sd[f'{prefix}cls_preds.{k}.weight'], sd[f'{prefix}cls_preds.{k}.bias'])
        reg_o = F.conv2d(rf, sd[f'{prefix}reg_preds.{k}.weight'], sd[f'{prefix}reg_preds.{k}.bias'])
        obj_o = F.conv2d(rf, sd[f'{prefix}obj_preds.{k}.weight'], sd[f'{prefix}obj_preds.{k}.bias'])
        outs.append(torch.cat([reg_o, obj_o, cls_o], 1))
    return outs


def forward(x, sd):
    with torch.no_grad():
        return head(neck(backbone(x, sd), sd), sd)


def decode(outputs, strides=STRIDES, num_classes=80):
    """yolox.py:19-51 -> float32 [B, A, 5 + nc] with corner boxes, sigmoid obj / classes."""
    with torch.no_grad():
        grids, ss = [], []
        for o, s in zip(outputs, strides):
            h, w = o.shape[-2:]
            yv, xv = torch.meshgrid([torch.arange(h), torch.arange(w)], indexing='ij')
            grids.append(torch.stack((xv, yv), 2).view(1, -1, 2))
            ss.append(torch.full((1, h * w, 1), s))
        out = torch.cat([o.flatten(start_dim=2) for o in outputs], dim=2).permute(0, 2, 1).clone()
        g = torch.cat(grids, 1).type(out.dtype)
        s = torch.cat(ss, 1).type(out.dtype)
        out[..., 0:2] = (out[..., 0:2] + g) * s
        out[..., 2:4] = torch.exp(out[..., 2:4]) * s
        out[..., 4:5] = torch.sigmoid(out[..., 4:5])
        out[..., 5:5 + num_classes] = torch.sigmoid(out[..., 5:5 + num_classes])
        box = out.new(out.shape)
        box[:, :, 0] = out[:, :, 0] - out[:, :, 2] / 2
        box[:, :, 1] = out[:, :, 1] - out[:, :, 3] / 2
        box[:, :, 2] = out[:, :, 0] + out[:, :, 2] / 2
        box[:, :, 3] = out[:, :, 1] + out[:, :, 3] / 2
        out[:, :, :4] = box[:, :, :4]
        return out


def records(decoded, num_classes=80):
    """[B, A, 5+nc] decoded tensor -> the 8-float candidate records of the C ABI: (x1,y1,x2,y2,obj,class_conf,class_pred,obj*class_conf)."""
    d = decoded.numpy() if isinstance(decoded, torch.Tensor) else np.asarray(decoded)
    cls = d[:, :, 5:5 + num_classes]
    cp = cls.argmax(2)  # first maximum, like torch.max
    cc = np.take_along_axis(cls, cp[..., None], 2)[..., 0]
    rec = np.zeros(d.shape[:2] + (8,), np.float32)
    rec[..., :5] = d[..., :5]
    rec[..., 5] = cc
    rec[..., 6] = cp.astype(np.float32)
    rec[..., 7] = d[..., 4] * cc
    return rec


def batched_nms(boxes, scores, idxs, iou_thr, vanilla_above=1000):
    """torchvision.ops.batched_nms on CPU tensors (boxes.py): numel > 4000 <=> more than 1000 boxes -> vanilla."""
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    boxes = boxes.astype(np.float32)
    scores = scores.astype(np.float32)
    if n > vanilla_above:
        keep_mask = np.zeros(n, bool)
        for c in np.unique(idxs):
            cur = np.where(idxs == c)[0]
            keep_mask[cur[greedy_nms(boxes[cur], scores[cur], iou_thr)]] = True
        keep = np.where(keep_mask)[0]
        return keep[np.argsort(-scores[keep], kind='stable')]
    max_coordinate = boxes.max()
    offsets = idxs.astype(np.float32) * (max_coordinate + np.float32(1))
    return np.asarray(greedy_nms(boxes + offsets[:, None], scores, iou_thr), np.int64)


def nms_records(rec, conf_thre=0.01, nms_thre=0.65, vanilla_above=1000):
    """yolox.py:54-67 on candidate records [A, 8] of ONE image -> (rows [n,7], kept location indices [n])."""
    rec = np.asarray(rec, np.float32)
    mask = rec[:, 7] >= np.float32(conf_thre)
    loc = np.where(mask)[0]
    det = rec[loc]
    if det.shape[0] == 0:
        return np.zeros((0, 7), np.float32), np.zeros((0,), np.int64)
    keep = batched_nms(det[:, :4], det[:, 4] * det[:, 5], det[:, 6], nms_thre, vanilla_above)
    return det[keep, :7], loc[keep]


def post_process(outputs, strides=STRIDES, num_classes=80, conf_thre=0.01, nms_thre=0.65):
    """yolox_post_process: list of raw head outputs -> list (per image) of rows [n,7] (None-equivalent: empty array)."""
    rec = records(decode(outputs, strides, num_classes), num_classes)
    return [nms_records(r, conf_thre, nms_thre) for r in rec]


def make_stress_records(A=8972, nc=80, regime='typical', seed=2, img=640):
    """Candidate records for NMS tests: clustered boxes with unique scores.  'few' (< 1000 pass -> coordinate trick),
    'typical' (> 1000 pass -> per-class NMS), 'all' (every location passes)."""
    rng = np.random.default_rng(seed)
    G = int(rng.integers(20, 120))
    gxy = rng.uniform(40, img - 40, size=(G, 2))
    gwh = rng.uniform(16, 200, size=(G, 2))
    gcls = rng.integers(0, nc, size=G)
    owner = rng.integers(0, G, size=A)
    cxy = gxy[owner] + rng.normal(0, 6, size=(A, 2))
    wh = (gwh[owner] * (1 + rng.normal(0, 0.1, size=(A, 2)))).clip(2, None)
    frac = {'few': 0.05, 'typical': 0.4, 'all': 1.1}[regime]
    on = rng.random(A) < frac
    obj = np.where(on, rng.beta(4, 2, size=A), rng.beta(1, 60, size=A) * 0.05).astype(np.float32)
    cc = rng.beta(5, 2, size=A).astype(np.float32)
    cls = np.where(rng.random(A) < 0.9, gcls[owner], rng.integers(0, nc, size=A))
    obj = np.clip(obj + (np.arange(A) % 991).astype(np.float32) * np.float32(2.0 ** -20), 0, 1)  # unique-ish scores
    rec = np.zeros((A, 8), np.float32)
    rec[:, 0:2] = (cxy - wh / 2).astype(np.float32)
    rec[:, 2:4] = (cxy + wh / 2).astype(np.float32)
    rec[:, 4], rec[:, 5], rec[:, 6] = obj, cc, cls.astype(np.float32)
    rec[:, 7] = rec[:, 4] * rec[:, 5]
    return rec


def canonical_rows(rows):
    """Rows ordered by score (obj * class_conf) descending with ties broken by the row contents: the reference's final ordering of
    equal scores is implementation defined (unstable sort), so fixtures are compared in this canonical order."""
    rows = np.asarray(rows, np.float32)
    if rows.shape[0] == 0:
        return rows
    sc = rows[:, 4] * rows[:, 5]
    order = np.lexsort((rows[:, 3], rows[:, 2], rows[:, 1], rows[:, 0], -sc))
    return rows[order]
