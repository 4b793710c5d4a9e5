"""Generates tests/golden/input_transform.npz from the reference's own ToTensor + Normalize (run in the build container only).

The reference classes (src/data/transforms/det_transforms.py:80-109) are imported by file path with ``pycocotools`` stubbed
(absent here, only used by ConvertCocoPolysToMask).  Frames are seeded uint8 HWC images that contain every byte value."""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/src/data/transforms/det_transforms.py'


def main():
    sys.dont_write_bytecode = True
    pc = types.ModuleType('pycocotools')
    pc.mask = types.ModuleType('pycocotools.mask')
    sys.modules.setdefault('pycocotools', pc)
    sys.modules.setdefault('pycocotools.mask', pc.mask)
    spec = importlib.util.spec_from_file_location('ref_det_transforms', REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mean, std = [0.406, 0.456, 0.485], [0.225, 0.224, 0.229]  # conf/coco_yolov5_s.yml:59
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, size=(2, 32, 48, 3), dtype=np.uint8)
    frames[0, 0, :, :].reshape(-1)[:144] = np.arange(144) % 256   # every byte value appears in every channel position
    frames[0, 1, :, :].reshape(-1)[:144] = (np.arange(144) + 112) % 256
    outs = []
    for f in frames:
        sample = {'image': f.copy(), 'target': {'boxes': np.zeros((1, 4), np.float32)}}
        sample = mod.ToTensor()(sample)
        sample = mod.Normalize(mean=mean, std=std)(sample)
        outs.append(sample['image'].numpy())
    out = np.stack(outs)
    assert out.shape == (2, 3, 32, 48) and out.dtype == np.float32
    np.savez_compressed(os.path.join(ROOT, 'tests/golden/input_transform.npz'), frames=frames, tensor=out, mean=np.float32(mean), std=np.float32(std))
    print('wrote tests/golden/input_transform.npz', out.shape, float(out.min()), float(out.max()))


if __name__ == '__main__':
    main()
