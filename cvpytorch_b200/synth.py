"""Deterministic synthetic YOLOv5-s weights in the reference's ``state_dict`` format (workload generator).

There is no network, so there are no trained checkpoints: tests and bench.py use random-init weights of the
named architecture.  Naive random BatchNorm statistics collapse the features (logits become
bias-dominated and a low-precision kernel *appears* to pass), so the BN running statistics come from a
calibration pass of the REFERENCE model (tools/make_golden.py, run once in the build container) stored in
tests/golden/yolov5s_calib.npz (~40 KB):  conv weights ~ N(0, 1/fan_in) from a per-key seeded generator,
BN gamma ~ U(0.5,1.5), beta ~ N(0,0.2), running_mean/var := batch statistics of randn(4,3,640,640) (seed 7)
through the reference in train mode with momentum 1.0, Detect conv weights scaled so std(logit-bias)=1.5 and
biases set to the reference's prior (src/models/detects/yolov5_detect.py:29-36).
The same state_dict loads into the reference modules and into the drop-in modules.
"""
import hashlib
import math
import os

import numpy as np
import torch

CALIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'yolov5s_calib.npz')


def _gen(key):
    seed = int.from_bytes(hashlib.sha256(key.encode()).digest()[:4], 'little')
    return torch.Generator().manual_seed(seed)


def base_state_dict(template):
    """template: {key: tensor} with the reference's key names/shapes.  Returns un-calibrated synthetic values:
    conv weights ~ N(0, 1/fan_in), norm-layer gamma ~ U(0.5,1.5) / beta ~ N(0,0.2), conv biases 0, running stats (0,1)."""
    out = {}
    for k, v in template.items():
        g = _gen(k)
        stem = k.rsplit('.', 1)[0]
        is_norm = (stem + '.weight') in template and template[stem + '.weight'].dim() == 1
        if k.endswith('num_batches_tracked'):
            out[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith('anchors'):
            out[k] = v.clone().float()
        elif k.endswith('.scale'):
            out[k] = torch.ones(v.shape)
        elif k.endswith('running_mean'):
            out[k] = torch.zeros(v.shape)
        elif k.endswith('running_var'):
            out[k] = torch.ones(v.shape)
        elif k.endswith('.weight') and v.dim() == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            out[k] = torch.randn(v.shape, generator=g) / math.sqrt(fan_in)
        elif k.endswith('.weight') and v.dim() == 1:
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith('.bias') and is_norm:
            out[k] = torch.randn(v.shape, generator=g) * 0.2
        elif k.endswith('.bias'):
            out[k] = torch.zeros(v.shape)
        else:
            raise KeyError(f'unexpected key {k}')
    return out


def detect_prior_bias(num_classes=80, na=3, strides=(8.0, 16.0, 32.0)):
    """Bias prior of YOLOv5Detect.init_weight (src/models/detects/yolov5_detect.py:29-36) on a zero base."""
    outs = []
    for s in strides:
        b = torch.zeros(na, num_classes + 5)
        b[:, 4] += math.log(8 / (640 / s) ** 2)
        b[:, 5:] += math.log(0.6 / (num_classes - 0.999999))
        outs.append(b.view(-1))
    return outs


def apply_calibration(sd, calib):
    """calib: mapping with '<bn prefix>.running_mean/var' arrays and 'detect_scale' [3]."""
    for k in list(sd.keys()):
        if k.endswith('running_mean') or k.endswith('running_var'):
            sd[k] = torch.from_numpy(np.asarray(calib[k])).float().clone()
    scale = np.asarray(calib['detect_scale'])
    pri = detect_prior_bias()
    for i in range(3):
        sd[f'detect.m.{i}.weight'] = sd[f'detect.m.{i}.weight'] * float(scale[i])
        sd[f'detect.m.{i}.bias'] = pri[i].clone()
    return sd


def template_state_dict():
    """Key/shape template from the drop-in modules (identical to the reference's, tests/test_host_logic.py)."""
    from . import models as M
    bb = M.build_backbone({'name': 'YOLOv5CSPDarknet', 'subtype': 'yolov5_s', 'out_stages': [2, 3, 4]})
    nk = M.build_neck({'name': 'YOLOv5Neck', 'in_channels': [256, 512, 1024], 'out_channels': [256, 512, 1024],
                       'depth_mul': 0.33, 'width_mul': 0.5})
    dt = M.build_detect({'name': 'YOLOv5Detect', 'in_channels': [256, 512, 1024], 'depth_mul': 0.33, 'width_mul': 0.5,
                         'anchors': M.YOLOv5.anchors, 'num_classes': 80})
    t = {}
    for p, m in (('backbone.', bb), ('neck.', nk), ('detect.', dt)):
        for k, v in m.state_dict().items():
            t[p + k] = v
    return t


def yolov5s_state_dict(calibrated=True):
    sd = base_state_dict(template_state_dict())
    if calibrated:
        if not os.path.exists(CALIB_PATH):
            raise FileNotFoundError(f'{CALIB_PATH} missing: run tools/make_golden.py in the build container')
        sd = apply_calibration(sd, np.load(CALIB_PATH))
    return sd


def split_prefix(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


YOLOV5S_CFG = {'TYPE': 'yolov5_s',
               'BACKBONE': {'name': 'YOLOv5Backbone', 'out_stages': [2, 3, 4], 'output_stride': 32, 'pretrained': False},
               'NECK': {'name': 'YOLOv5Neck', 'in_channels': [256, 512, 1024], 'out_channels': [256, 512, 1024]},
               'DETECT': {'name': 'YOLOv5Detect', 'in_channels': [256, 512, 1024]},
               'LOSS': {'name': 'YOLOv5Loss', 'hyp_box': 0.05, 'hyp_obj': 1.0, 'hyp_cls': 0.5}}  # conf/coco_yolov5_s.yml:65-71


def build_yolov5s(calibrated=True, device=None):
    """Drop-in YOLOv5 model (the conf/coco_yolov5_s.yml configuration) with the synthetic weights loaded."""
    from . import models as M
    dictionary = [{f'c{i}': 1.0} for i in range(80)]
    model = M.YOLOv5(dictionary=dictionary, model_cfg=dict(YOLOV5S_CFG))
    model.load_state_dict(yolov5s_state_dict(calibrated), strict=True)
    model.eval()
    if device is not None:
        model.to(device)
    return model


# ================================================================================================= FCOS-R50
FCOS_CALIB_PATH = os.path.join(os.path.dirname(CALIB_PATH), 'fcos_calib.npz')
FCOS_CFG = {'BACKBONE': {'name': 'ResNet', 'subtype': 'resnet50', 'out_stages': [2, 3, 4], 'output_stride': 32, 'pretrained': True},
            'NECK': {'name': 'FCOSFPN', 'in_channels': [512, 1024, 2048], 'out_channels': 256},
            'HEAD': {'name': 'FCOSHead', 'in_channel': 256, 'GN': True, 'cnt_on_reg': True, 'prior': 0.01},
            'LOSS': {'name': 'FCOSLoss', 'strides': [8, 16, 32, 64, 128]},
            'DETECT': {'name': 'FCOSDetect', 'score_threshold': 0.05, 'nms_iou_threshold': 0.6, 'max_detection_boxes_num': 1000,
                       'strides': [8, 16, 32, 64, 128]}}  # conf/coco_fcos.yml:58-64


def fcos_template_state_dict(num_classes=80):
    from . import fcos_models as FM
    bb = FM.build_backbone({**FCOS_CFG['BACKBONE'], 'pretrained': False})
    nk = FM.build_neck(FCOS_CFG['NECK'])
    hd = FM.build_head({**FCOS_CFG['HEAD'], 'num_classes': num_classes})
    t = {}
    for p, m in (('backbone.', bb), ('neck.', nk), ('head.', hd)):
        for k, v in m.state_dict().items():
            t[p + k] = v
    return t


def fcos_apply_calibration(sd, calib):
    for k in list(sd.keys()):
        if k.endswith('running_mean') or k.endswith('running_var'):
            sd[k] = torch.from_numpy(np.asarray(calib[k])).float().clone()
    sc = np.asarray(calib['head_scale'])  # (cls, cnt, reg)
    sd['head.cls_logits.weight'] = sd['head.cls_logits.weight'] * float(sc[0])
    sd['head.cnt_logits.weight'] = sd['head.cnt_logits.weight'] * float(sc[1])
    sd['head.reg_pred.weight'] = sd['head.reg_pred.weight'] * float(sc[2])
    sd['head.cls_logits.bias'] = torch.full_like(sd['head.cls_logits.bias'], -math.log((1 - 0.01) / 0.01))  # fcos_head.py:61
    sd['head.reg_pred.bias'] = torch.full_like(sd['head.reg_pred.bias'], 2.0)  # ltrb ~ exp(2 +- 1) px so neighbouring boxes overlap
    return sd


def fcos_state_dict(calibrated=True):
    sd = base_state_dict(fcos_template_state_dict())
    if calibrated:
        if not os.path.exists(FCOS_CALIB_PATH):
            raise FileNotFoundError(f'{FCOS_CALIB_PATH} missing: run tools/make_golden.py in the build container')
        sd = fcos_apply_calibration(sd, np.load(FCOS_CALIB_PATH))
    return sd


def build_fcos(calibrated=True):
    from . import fcos_models as FM
    dictionary = [{f'c{i}': 1.0} for i in range(80)]
    model = FM.FCOS(dictionary=dictionary, model_cfg=dict(FCOS_CFG))
    model.load_state_dict(fcos_state_dict(calibrated), strict=True)
    model.eval()
    return model


# ================================================================================================= DeepLabv3+ R50v1c
DEEPLAB_CALIB_PATH = os.path.join(os.path.dirname(CALIB_PATH), 'deeplab_calib.npz')
DEEPLAB_CFG = {'BACKBONE': {'name': 'ResNet', 'subtype': 'resnet50v1c', 'out_stages': [1, 4], 'output_stride': 8, 'pretrained': True},
               'NECK': None, 'AUX_HEAD': None,
               'HEAD': {'name': 'Deeplabv3PlusHead', 'num_classes': 19, 'in_channels': 2048, 'channels': 512, 'dilations': [1, 12, 24, 36],
                        'low_in_channels': 256, 'low_channels': 48},
               'LOSS': {'name': 'CrossEntropyLoss2d'}}  # conf/seg/deeplabv3plus/cityscapes_deeplabv3plus_r50.yml:57-62


def deeplab_template_state_dict():
    from . import seg_models as SM
    bb = SM.build_backbone({**DEEPLAB_CFG['BACKBONE'], 'pretrained': False})
    hd = SM.build_head(DEEPLAB_CFG['HEAD'])
    t = {}
    for p, m in (('backbone.', bb), ('head.', hd)):
        for k, v in m.state_dict().items():
            t[p + k] = v
    return t


def deeplab_apply_calibration(sd, calib):
    for k in list(sd.keys()):
        if k.endswith('running_mean') or k.endswith('running_var'):
            sd[k] = torch.from_numpy(np.asarray(calib[k])).float().clone()
    sd['head.cls_seg.weight'] = sd['head.cls_seg.weight'] * float(np.asarray(calib['cls_scale']))
    g = _gen('head.cls_seg.bias.synth')
    sd['head.cls_seg.bias'] = torch.randn(sd['head.cls_seg.bias'].shape, generator=g) * 0.5
    return sd


def deeplab_state_dict(calibrated=True):
    sd = base_state_dict(deeplab_template_state_dict())
    if calibrated:
        if not os.path.exists(DEEPLAB_CALIB_PATH):
            raise FileNotFoundError(f'{DEEPLAB_CALIB_PATH} missing: run tools/make_golden_deeplab.py in the build container')
        sd = deeplab_apply_calibration(sd, np.load(DEEPLAB_CALIB_PATH))
    return sd


def build_deeplab(calibrated=True):
    from . import seg_models as SM
    dictionary = [{f'c{i}': 1.0} for i in range(19)]
    model = SM.EncoderDecoder(dictionary=dictionary, model_cfg=dict(DEEPLAB_CFG))
    model.load_state_dict(deeplab_state_dict(calibrated), strict=True)
    model.eval()
    return model


# ================================================================================================= YOLOX-s
YOLOX_CALIB_PATH = os.path.join(os.path.dirname(CALIB_PATH), 'yolox_calib.npz')
YOLOX_CFG = {'TYPE': 'yolox_s',
             'BACKBONE': {'name': 'CSPDarknet', 'subtype': 'yolox_s', 'out_stages': [2, 3, 4], 'output_stride': 32, 'pretrained': False},
             'NECK': {'name': 'YOLOXNeck', 'channels': [256, 512, 1024]},
             'HEAD': {'name': 'YOLOXHead', 'in_channels': [256, 512, 1024]}}


def yolox_template_state_dict(num_classes=80):
    from . import yolox_models as XM
    bb = XM.build_backbone(YOLOX_CFG['BACKBONE'])
    nk = XM.build_neck({**YOLOX_CFG['NECK'], 'depth_mul': 0.33, 'width_mul': 0.5})
    hd = XM.build_head({**YOLOX_CFG['HEAD'], 'depth_mul': 0.33, 'width_mul': 0.5, 'num_classes': num_classes})
    t = {}
    for p, m in (('backbone.', bb), ('neck.', nk), ('head.', hd)):
        for k, v in m.state_dict().items():
            t[p + k] = v
    return t


def yolox_apply_calibration(sd, calib):
    """BN statistics of the calibration pass + predictor scales / biases chosen so that a 640x640 noise image yields a few
    thousand candidates with overlapping boxes (the un-calibrated head emits none: sigmoid(-4.6)^2 << 0.01)."""
    for k in list(sd.keys()):
        if k.endswith('running_mean') or k.endswith('running_var'):
            sd[k] = torch.from_numpy(np.asarray(calib[k])).float().clone()
    sc = np.asarray(calib['pred_scale'])  # [3 levels][cls, reg, obj]
    for i in range(3):
        sd[f'head.cls_preds.{i}.weight'] = sd[f'head.cls_preds.{i}.weight'] * float(sc[i, 0])
        sd[f'head.reg_preds.{i}.weight'] = sd[f'head.reg_preds.{i}.weight'] * float(sc[i, 1])
        sd[f'head.obj_preds.{i}.weight'] = sd[f'head.obj_preds.{i}.weight'] * float(sc[i, 2])
        sd[f'head.cls_preds.{i}.bias'] = torch.full_like(sd[f'head.cls_preds.{i}.bias'], -2.0)
        sd[f'head.obj_preds.{i}.bias'] = torch.full_like(sd[f'head.obj_preds.{i}.bias'], -2.0)
        sd[f'head.reg_preds.{i}.bias'] = torch.tensor([0.0, 0.0, 1.6, 1.6])  # wh ~ 5 strides: neighbouring boxes overlap
    return sd


def yolox_state_dict(calibrated=True):
    sd = base_state_dict(yolox_template_state_dict())
    if calibrated:
        if not os.path.exists(YOLOX_CALIB_PATH):
            raise FileNotFoundError(f'{YOLOX_CALIB_PATH} missing: run tools/make_golden_yolox.py in the build container')
        sd = yolox_apply_calibration(sd, np.load(YOLOX_CALIB_PATH))
    return sd


def build_yolox(calibrated=True):
    from . import yolox_models as XM
    dictionary = [{f'c{i}': 1.0} for i in range(80)]
    model = XM.YOLOX(dictionary=dictionary, model_cfg=dict(YOLOX_CFG))
    model.load_state_dict(yolox_state_dict(calibrated), strict=True)
    model.eval()
    return model
