"""Fused detect head (CVB_OUT_YOLO) vs fp32-out conv + cvb_yolo_decode: time per level and per-role cycle shares of the fused kernel.
  python tools/head_profile.py [B]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['CVB_DIAG_LIB'] = '1'  # the diagnostics build (csrc/build.sh diag)
import torch  # noqa: E402

from cvpytorch_b200 import _lib, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
na, nc = 3, 80
no = nc + 5


def timeit(fn, reps=5):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for cin, n, stride in ((128, 80, 8.0), (256, 40, 16.0), (512, 20, 32.0)):
    g = torch.Generator().manual_seed(cin)
    tin = ops.SplitTensor(B, n, n, cin)
    tin.data.normal_(0, 1)
    tin.data[1].mul_(2 ** -11)
    w = (torch.randn(na * no, cin, 1, 1, generator=g) * (1.0 / cin ** 0.5)).double()
    b = (torch.randn(na * no, generator=g) * 0.5 - 4.0).double()   # few scores above conf, like the calibrated model
    anchors_px = torch.tensor([[10., 13.], [16., 30.], [33., 23.]])
    A = na * n * n
    z = torch.zeros(B, A, no, device='cuda')
    ws = ops.NmsWorkspace(B, A, nc)
    raw = ops.F32Tensor(B, n, n, 256)
    wp, bp = ops.pack_conv_weights(w, b)
    pc = ops.ConvPlan(tin.view(), raw.view(0, na * no), wp, bp, 1, 1, 0, 1, None)
    apx = anchors_px.cuda().contiguous()
    t_conv = timeit(pc.run)
    t_dec = timeit(lambda: ops.yolo_decode(raw.view(0, na * no), na, no, apx, stride, z, A, 0, None, ws, 0.001, True))
    print(f'level {n}x{n} cin {cin} B {B}: conv(fp32 raw) {t_conv:.4f} ms + decode {t_dec:.4f} ms = {t_conv + t_dec:.4f} ms')
    wy, by = ops.pack_yolo_head_weights(w, b, na, no)
    for dbg in (0, 32, 64, 96, 4 | 96):
        os.environ['CVB_DBG'] = str(dbg)
        y = ops.yolo_decode_desc(na, no, anchors_px, stride, z, A, 0, ws, 0.001, True)
        pf = ops.ConvPlan(tin.view(), ops.CvbView(z.data_ptr(), B, n, n, na * 128, na * 128, 0), wy, by, 1, 1, 0, 1, None, yolo=y)
        grid = ctypes.c_int32(0)
        buf = torch.zeros(1024 * 16, dtype=torch.int64, device='cuda')
        _lib.check(_lib.lib().cvb_conv_plan_set_profile(pf.handle, buf.data_ptr(), ctypes.byref(grid)), 'set_profile')
        t = timeit(pf.run)
        c = buf[:grid.value * 16].view(grid.value, 16).double().mean(0).tolist()
        tot = max(c[4], 1.0)
        print(f'  fused dbg={dbg:3d} {t:.4f} ms grid {grid.value} | cycles/CTA {c[4]:9.0f} | producer waitA {c[1] / max(c[0], 1):5.1%} | MMA wait acc {c[5] / tot:5.1%} act {c[6] / tot:5.1%} '
              f'issue {(c[4] - c[5] - c[6] - c[7]) / tot:5.1%} | epilogue (thread 0) wait acc {c[9] / max(c[8], 1):5.1%} stage-free {c[10] / max(c[8], 1):5.1%} convert {c[11] / max(c[8], 1):5.1%} '
              f'barrier {c[12] / max(c[8], 1):5.1%} copy+rowmax {c[13] / max(c[8], 1):5.1%}', flush=True)
    os.environ.pop('CVB_DBG', None)
