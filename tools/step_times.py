"""In-sequence step time (CUDA events, 10 steps after 3 warm-ups, device-resident synthetic input) of the secondary configurations:
  python tools/step_times.py [fcos] [deeplab] [yolox]
FCOS R50 800x800 bs32 (BASELINE.json configs[4]), DeepLabv3+ R50v1c 1024x2048 bs16 (configs[2]), YOLOX-s 640x640 bs64 (inference
half of configs[3])."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cvpytorch_b200 import synth  # noqa: E402


def timed(fn, n=10, w=3):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


which = sys.argv[1:] or ['fcos', 'deeplab', 'yolox']
torch.manual_seed(1029)
if 'yolox' in which:
    m = synth.build_yolox(True)
    x = torch.randn(64, 3, 640, 640).cuda()
    ms = timed(lambda: m.predict(x))
    det, cnt = m.predict(x)
    print(f'YOLOX-s 640x640 bs64 (forward + decode + batched_nms): {ms:.2f} ms/step -> {64 / ms * 1e3:.1f} img/s; kept/img mean {float(cnt.float().mean()):.0f}')
    del m, x
    torch.cuda.empty_cache()
if 'fcos' in which:
    m = synth.build_fcos(True)
    x = torch.randn(32, 3, 800, 800).cuda()
    ms = timed(lambda: m.predict(x))
    out = m.predict(x)
    print(f'FCOS R50 800x800 bs32: {ms:.2f} ms/step -> {32 / ms * 1e3:.1f} img/s; kept/img mean {float(out[-1].float().mean()):.0f}')
    del m, x, out
    torch.cuda.empty_cache()
if 'deeplab' in which:
    m = synth.build_deeplab(True)
    x = torch.randn(16, 3, 1024, 2048).cuda()
    ms = timed(lambda: m.predict(x))
    lab = m.predict(x)
    print(f'DeepLabv3+ R50v1c 1024x2048 bs16: {ms:.2f} ms/step -> {16 / ms * 1e3:.1f} img/s; labels {tuple(lab.shape)} {lab.dtype}; '
          f'mem {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB')
