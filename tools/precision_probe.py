"""Measures the accumulation error of the tcgen05 conv kernel against an exact (float64) convolution of the SAME
hi+lo operands, as a function of the reduction length K -- separates tensor-core accumulation behaviour
(round-toward-zero in the fp32 adder) from the fp16-pair representation error."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from cvpytorch_b200 import ops  # noqa: E402


def probe(cin, k, cout=128, B=2, H=32, W=32, positive=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    if positive:
        x, w = x.abs(), w.abs()
    tin = ops.SplitTensor(B, H, W, cin)
    ops.nchw_to_split(x.cuda(), tin.view())
    wp, bp = ops.pack_conv_weights(w.double(), torch.zeros(cout, dtype=torch.float64))
    out = ops.F32Tensor(B, H, W, cout)
    plan = ops.ConvPlan(tin.view(), out.view(), wp, bp, k, 1, k // 2, 1, None)
    plan.run()
    torch.cuda.synchronize()
    y = out.data.permute(0, 3, 1, 2).double()
    xr = (tin.data[0].double() + tin.data[1].double()).permute(0, 3, 1, 2)             # exact hi+lo operands
    wr = (wp[0].double() + wp[1].double())[:cout].reshape(cout, k, k, cin).permute(0, 3, 1, 2)
    y64 = F.conv2d(xr, wr, None, 1, k // 2)
    yx = F.conv2d(x.cuda().double(), w.cuda().double(), None, 1, k // 2)               # exact conv of the un-split operands
    d = y - y64
    rel_max = float(d.abs().max() / y64.abs().max())
    rel_rms = float(d.pow(2).mean().sqrt() / y64.pow(2).mean().sqrt())
    bias_mag = float((y.abs() - y64.abs()).mean() / y64.abs().mean())
    rep = float((y64 - yx).abs().max() / yx.abs().max())
    n_mma = cin * k * k // 16 * 3
    print(f'cin {cin:5d} k{k} K={cin * k * k:5d} mma/chain {n_mma:4d} positive={int(positive)}: accum err max {rel_max:.2e} rms {rel_rms:.2e} '
          f'|y| bias {bias_mag:+.2e}   (representation err of the fp16 pairs: {rep:.2e})')


if __name__ == '__main__':
    for positive in (False, True):
        for cin, k in ((64, 1), (256, 1), (1024, 1), (64, 3), (256, 3)):
            probe(cin, k, positive=positive)
