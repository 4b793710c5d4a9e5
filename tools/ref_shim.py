"""Import shim for the *reference* (shanglianlm0525/CvPytorch @ /root/reference).

Only used inside the build container to (a) pin the oracle against the reference's own
code and (b) generate the golden fixtures under tests/golden/.  Never imported by the
product, the GPU tests or bench.py (the GPU box has no /root/reference).
Recipe follows SURVEY.md Appendix A; nothing under /root/reference is modified.
"""
import importlib
import math
import sys

REF = '/root/reference'


def install():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import numpy
    numpy.math = math
    table = {
        'torchvision.models.convnext': ['_MODELS_URLS'], 'torchvision.models.efficientnet': ['model_urls'],
        'torchvision.models.mnasnet': ['_MODEL_URLS'], 'torchvision.models.mobilenetv2': ['model_urls'],
        'torchvision.models.mobilenetv3': ['model_urls'], 'torchvision.models.regnet': ['model_urls'],
        'torchvision.models.resnet': ['model_urls'], 'torchvision.models.shufflenetv2': ['model_urls'],
        'torchvision.models.squeezenet': ['model_urls'], 'torchvision.models.vision_transformer': ['model_urls'],
        'torchvision.models.vgg': ['model_urls'], 'torchvision.models.densenet': ['model_urls']}
    for mod, attrs in table.items():
        m = importlib.import_module(mod)
        for a in attrs:
            if not hasattr(m, a):
                setattr(m, a, {})


def build_yolov5s():
    """Returns (backbone, neck, detect, non_max_suppression) of the reference, composed by hand
    exactly as SURVEY.md §3.5 row 1 describes (the YAML route does not build)."""
    install()
    from src.models.backbones import build_backbone
    from src.models.detects import build_detect
    from src.models.yolov5 import YOLOv5, non_max_suppression
    oldneck = importlib.import_module('src.models.necks.yolov5_neck')
    bb = build_backbone({'name': 'YOLOv5CSPDarknet', 'subtype': 'yolov5_s', 'out_stages': [2, 3, 4]})
    nk = oldneck.YOLOv5Neck([256, 512, 1024], [256, 512, 1024], depth_mul=0.33, width_mul=0.5)
    dt = build_detect({'name': 'YOLOv5Detect', 'in_channels': [256, 512, 1024], 'depth_mul': 0.33,
                       'width_mul': 0.5, 'anchors': YOLOv5.anchors, 'num_classes': 80})
    return bb, nk, dt, non_max_suppression
