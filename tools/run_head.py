"""Runs the fused detect head (1x1 conv + YOLOv5 decode epilogue, CVB_OUT_YOLO) of one level a few times (for ncu captures):
  python tools/run_head.py [cin=128] [n=80] [B=64]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cvpytorch_b200 import ops  # noqa: E402

cin = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else 80
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
na, nc = 3, 80
no = nc + 5
g = torch.Generator().manual_seed(cin)
tin = ops.SplitTensor(B, n, n, cin)
tin.data.normal_(0, 1)
tin.data[1].mul_(2 ** -11)
w = (torch.randn(na * no, cin, 1, 1, generator=g) * (1.0 / cin ** 0.5)).double()
b = (torch.randn(na * no, generator=g) * 0.5 - 4.0).double()
A = na * n * n
z = torch.zeros(B, A, no, device='cuda')
ws = ops.NmsWorkspace(B, A, nc)
wy, by = ops.pack_yolo_head_weights(w, b, na, no)
y = ops.yolo_decode_desc(na, no, torch.tensor([[10., 13.], [16., 30.], [33., 23.]]), 8.0, z, A, 0, ws, 0.001, True)
plan = ops.ConvPlan(tin.view(), ops.CvbView(z.data_ptr(), B, n, n, na * 128, na * 128, 0), wy, by, 1, 1, 0, 1, None, yolo=y)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ts = []
for _ in range(3):
    flush.zero_()
    ops.nms_reset(ws)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
byts = 4.0 * B * n * n * cin + 4.0 * B * A * no
print(f'fused head {cin}->{na}x{no} {n}x{n} B{B}: {min(ts):.4f} ms  {byts / min(ts) / 1e6:.0f} GB/s (input split16 + z fp32)')
