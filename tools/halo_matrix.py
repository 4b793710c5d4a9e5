"""A/B matrix of the conv loader modes on single layers (GPU box).  One line per configuration: time, plan."""
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cvpytorch_b200 import ops  # noqa: E402

LAYERS = [(64, 64, 3, 1, 80, 80, 64), (32, 32, 3, 1, 160, 160, 64), (128, 128, 3, 1, 40, 40, 64), (32, 64, 3, 2, 320, 320, 64),
          (64, 128, 3, 2, 160, 160, 64), (256, 256, 3, 1, 20, 20, 64)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def bench(cin, cout, k, s, H, W, B, env, reps=7):
    for kk in ('CVB_HALO', 'CVB_HALO_BK', 'CVB_HALO_CTAS', 'CVB_HALO_RES'):
        os.environ.pop(kk, None)
    os.environ.update(env)
    g = torch.Generator().manual_seed(0)
    tin = ops.SplitTensor(B, H, W, cin)
    tin.data.normal_(0, 1)
    tin.data[1].mul_(2 ** -11)
    w = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64) / (cin * k * k) ** 0.5
    wp, bp = ops.pack_conv_weights(w, torch.zeros(cout, dtype=torch.float64))
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    out = ops.SplitTensor(B, Ho, Wo, cout)
    try:
        plan = ops.ConvPlan(tin.view(), out.view(), wp, bp, k, s, k // 2, 1, 'silu')
    except Exception as e:  # infeasible combination
        return None, str(e)[:60]
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2], ''


os.environ['CVB_PLAN_DEBUG'] = '1'
for L in LAYERS:
    t0, _ = bench(*L, {'CVB_HALO': '0'})
    print(f'layer {L}: classic {t0:.4f} ms', flush=True)
    for mode in (2, 1):
        for bk in (64, 32):
            for ctas in (1, 2):
                for res in (1, 0):
                    env = {'CVB_HALO': str(mode), 'CVB_HALO_BK': str(bk), 'CVB_HALO_CTAS': str(ctas), 'CVB_HALO_RES': str(res)}
                    t, err = bench(*L, env)
                    print(f'   halo={mode} bk={bk} ctas={ctas} res={res}: ' + (f'{t:.4f} ms  ({t / t0:.2f}x)' if t else f'-- {err}'), flush=True)
