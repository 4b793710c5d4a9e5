"""Runs the training conv kernels of one layer shape a few times (for ncu captures):  python tools/run_train_kernels.py cin cout k H W B"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cvpytorch_b200 import train as T  # noqa: E402

cin, cout, k, H, W, B = map(int, sys.argv[1:7])
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, cin, generator=g).cuda().to(torch.bfloat16)
dy = torch.randn(B, H, W, cout, generator=g).cuda().to(torch.bfloat16)
wf, wb = T.pack_weights((torch.randn(cout, cin, k, k, generator=g) * 0.05).cuda())
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for name, fn in (('forward', lambda: T.conv(x, wf, cout, k)), ('backward-data', lambda: T.conv(dy, wb, cin, k)), ('backward-weight', lambda: T.conv_wgrad(x, dy, k))):
    ts = []
    for _ in range(3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    fl = 2.0 * B * H * W * cin * cout * k * k
    print(f'{name:16s} {cin}->{cout} k{k} {H}x{W} B{B}: {min(ts):.4f} ms  {fl / min(ts) / 1e9:.1f} TF/s  {2.0 * B * H * W * (cin + cout) / min(ts) / 1e6:.0f} GB/s')
