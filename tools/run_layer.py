"""Runs one fused conv layer a few times (for ncu captures):  python tools/run_layer.py cin cout k stride H W B [reps] [block_n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cvpytorch_b200 import ops  # noqa: E402

cin, cout, k, s, H, W, B = map(int, sys.argv[1:8])
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 3
bn = int(sys.argv[9]) if len(sys.argv) > 9 else 0
g = torch.Generator().manual_seed(0)
tin = ops.SplitTensor(B, H, W, cin)
tin.data.normal_(0, 1)
tin.data[1].mul_(2 ** -11)
w = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64) / (cin * k * k) ** 0.5
wp, bp = ops.pack_conv_weights(w, torch.zeros(cout, dtype=torch.float64))
Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
out = ops.SplitTensor(B, Ho, Wo, cout)
plan = ops.ConvPlan(tin.view(), out.view(), wp, bp, k, s, k // 2, 1, 'silu', block_n=bn)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
byts = 4.0 * (B * H * W * cin + B * Ho * Wo * cout)
print(f'conv {cin}->{cout} k{k} s{s} {H}x{W} B{B}: {min(ts):.4f} ms  {byts / min(ts) / 1e6:.0f} GB/s  {2.0 * B * Ho * Wo * cout * cin * k * k / min(ts) / 1e9:.1f} TF/s')
