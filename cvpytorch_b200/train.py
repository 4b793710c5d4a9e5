"""Training-step drop-ins for the YOLOX C3 block (BASELINE.json configs[3]): forward AND backward on the B200 kernels.

Mirrors (paths relative to /root/reference), with the same constructor arguments and state_dict keys:
  BaseConv     src/models/modules/yolox_modules.py:35-55    nn.Conv2d(bias=False) -> nn.BatchNorm2d -> SiLU, training mode (batch statistics)
  Bottleneck   src/models/modules/yolox_modules.py:79-96    1x1 -> 3x3 (+ shortcut)
  CSPLayer     src/models/modules/yolox_modules.py:99-129   conv1 | conv2 -> bottlenecks -> cat -> conv3

What runs where: every convolution (forward, backward-data, backward-weight) is a bf16 tcgen05 implicit GEMM and every BatchNorm / SiLU pass
(statistics, apply, backward reductions, backward apply) is a kernel of libcvb200.so (csrc/train_kernels.cu).  torch.autograd only chains the
per-layer Functions and adds the gradients of the shortcut / concat fan-outs (tensor plumbing); the loss and the optimiser stay in PyTorch,
as in the reference trainer (trainer.py:177-207).  Precision: bf16 activations / gradients, fp32 accumulation, fp32 master weights and fp32
weight gradients (BASELINE.json configs[3] asks for a bf16 step).  No CPU fallback: CPU tensors raise.

Inside a block activations are NHWC bf16; `CSPLayer.forward` accepts / returns the reference's NCHW tensors (layout change at the boundary).
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _check_cuda(t, what):
    if not t.is_cuda:
        raise _lib.CvbError(f'{what}: CUDA tensor required (the B200 training path has no CPU fallback)')


def pack_weights(w):
    """fp32 [cout,cin,k,k] -> (bf16 [cout,k*k,cin] forward / wgrad layout, bf16 [cin,k*k,cout] backward-data operand)."""
    cout, cin, k, _ = w.shape
    wf = torch.empty((cout, k * k, cin), dtype=torch.bfloat16, device=w.device)
    wb = torch.empty((cin, k * k, cout), dtype=torch.bfloat16, device=w.device)
    _lib.check(_lib.lib().cvb_train_pack_weights(_p(w.detach().float().contiguous()), cout, cin, k, _p(wf), _p(wb), _stream()), 'cvb_train_pack_weights')
    return wf, wb


def conv(x, w_packed, cout, k, y_prev=None, stat_prev=None, stride=1):
    """x NHWC bf16 [B,H,W,cin] -> [B,Ho,Wo,cout] (pad k//2).  (y_prev, stat_prev): SiLU' epilogue of the producing layer (backward-data only)."""
    B, H, W, cin = x.shape
    Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    out = torch.empty((B, Ho, Wo, cout), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().cvb_train_conv(_p(x), B, H, W, cin, _p(w_packed), cout, k, stride, _p(out), _p(y_prev), _p(stat_prev), _stream()), 'cvb_train_conv')
    return out


def conv_dgrad_s2(dy, w_bwd, cin, H, W, y_prev=None, stat_prev=None):
    """Backward-data of the 3x3 / stride 2 / pad 1 convolution: dy [B,Ho,Wo,cout] -> dx [B,H,W,cin] (four parity sub-convolutions)."""
    B, Ho, Wo, cout = dy.shape
    dx = torch.empty((B, H, W, cin), dtype=torch.bfloat16, device=dy.device)
    _lib.check(_lib.lib().cvb_train_conv_dgrad_s2(_p(dy), B, Ho, Wo, cout, _p(w_bwd), cin, H, W, _p(dx), _p(y_prev), _p(stat_prev), _stream()),
               'cvb_train_conv_dgrad_s2')
    return dx


def conv_wgrad(x, dy, k, stride=1):
    """-> fp32 [cout, cin, k, k] (the nn.Conv2d layout)."""
    B, H, W, cin = x.shape
    cout = dy.shape[3]
    dw = torch.zeros((cout, k * k, cin), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().cvb_train_conv_wgrad(_p(x), _p(dy), B, H, W, cin, cout, k, stride, _p(dw), _stream()), 'cvb_train_conv_wgrad')
    return dw.view(cout, k, k, cin).permute(0, 3, 1, 2)


class _ConvBnSiLU(torch.autograd.Function):
    """a = silu(bn_train(conv(x, W))) with the backward the reference obtains from torch.autograd, on the B200 kernels."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, running_mean, running_var, eps, momentum, k, stride=1):
        _check_cuda(x, 'BaseConv (B200 training)')
        B, H, W, cin = x.shape
        cout = weight.shape[0]
        L = _lib.lib()
        wf, wb = pack_weights(weight)
        y = conv(x, wf, cout, k, stride=stride)
        npix = B * y.shape[1] * y.shape[2]
        stat = torch.empty((4, cout), dtype=torch.float32, device=x.device)
        scratch = torch.empty((2, cout), dtype=torch.float32, device=x.device)
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        _lib.check(L.cvb_train_bn_stats(_p(y), npix, cout, _p(g32), _p(b32), float(eps), float(momentum), _p(running_mean), _p(running_var),
                                        _p(scratch), _p(stat), _stream()), 'cvb_train_bn_stats')
        a = torch.empty_like(y)
        _lib.check(L.cvb_train_bn_silu_fwd(_p(y), npix, cout, _p(stat), _p(a), _stream()), 'cvb_train_bn_silu_fwd')
        ctx.save_for_backward(x, y, stat, g32, wb)
        ctx.k = k
        ctx.stride = stride
        ctx.need_dx = x.requires_grad
        return a

    @staticmethod
    def backward(ctx, da):
        x, y, stat, g32, wb = ctx.saved_tensors
        k = ctx.k
        B, H, W, cin = x.shape
        cout = y.shape[3]
        L = _lib.lib()
        da = da.contiguous()
        sums = torch.empty((2, cout), dtype=torch.float32, device=x.device)
        dy = torch.empty_like(y)
        _lib.check(L.cvb_train_bn_silu_bwd(_p(da), 0, _p(y), B * y.shape[1] * y.shape[2], cout, _p(stat), _p(g32), _p(sums), _p(dy), _stream()),
                   'cvb_train_bn_silu_bwd')
        dw = conv_wgrad(x, dy, k, ctx.stride)
        dx = None
        if ctx.need_dx:
            dx = conv(dy, wb, cin, k) if ctx.stride == 1 else conv_dgrad_s2(dy, wb, cin, H, W)
        return dx, dw, sums[1].clone(), sums[0].clone(), None, None, None, None, None, None


class _BottleneckChain(torch.autograd.Function):
    """Bottleneck body  t = BaseConv1x1(x);  u = BaseConv3x3(t)  as ONE autograd node, so that the backward-data convolution of the 3x3 applies
    the 1x1 layer's SiLU' in its epilogue (dz_t leaves the tensor core already multiplied; d(loss)/dt is never written)."""

    @staticmethod
    def forward(ctx, x, w1, g1, b1, rm1, rv1, w2, g2, b2, rm2, rv2, eps, momentum):
        _check_cuda(x, 'Bottleneck (B200 training)')
        B, H, W, cin = x.shape
        L = _lib.lib()
        npix = B * H * W

        def layer(inp, w, g, b, rm, rv, k):
            cout = w.shape[0]
            wf, wb = pack_weights(w)
            y = conv(inp, wf, cout, k)
            stat = torch.empty((4, cout), dtype=torch.float32, device=inp.device)
            scratch = torch.empty((2, cout), dtype=torch.float32, device=inp.device)
            g32, b32 = g.detach().float().contiguous(), b.detach().float().contiguous()
            _lib.check(L.cvb_train_bn_stats(_p(y), npix, cout, _p(g32), _p(b32), float(eps), float(momentum), _p(rm), _p(rv), _p(scratch), _p(stat), _stream()),
                       'cvb_train_bn_stats')
            a = torch.empty_like(y)
            _lib.check(L.cvb_train_bn_silu_fwd(_p(y), npix, cout, _p(stat), _p(a), _stream()), 'cvb_train_bn_silu_fwd')
            return y, stat, g32, wb, a

        y1, st1, g1f, wb1, t = layer(x, w1, g1, b1, rm1, rv1, 1)
        y2, st2, g2f, wb2, u = layer(t, w2, g2, b2, rm2, rv2, 3)
        ctx.save_for_backward(x, y1, st1, g1f, wb1, t, y2, st2, g2f, wb2)
        ctx.need_dx = x.requires_grad
        return u

    @staticmethod
    def backward(ctx, du):
        x, y1, st1, g1f, wb1, t, y2, st2, g2f, wb2 = ctx.saved_tensors
        B, H, W, cin = x.shape
        c1, c2 = y1.shape[3], y2.shape[3]
        npix = B * H * W
        L = _lib.lib()
        du = du.contiguous()
        sums2 = torch.empty((2, c2), dtype=torch.float32, device=x.device)
        dy2 = torch.empty_like(y2)
        _lib.check(L.cvb_train_bn_silu_bwd(_p(du), 0, _p(y2), npix, c2, _p(st2), _p(g2f), _p(sums2), _p(dy2), _stream()), 'cvb_train_bn_silu_bwd')
        dw2 = conv_wgrad(t, dy2, 3)
        dz1 = conv(dy2, wb2, c1, 3, y_prev=y1, stat_prev=st1)  # backward-data of the 3x3 with the 1x1 layer's SiLU' in the epilogue
        sums1 = torch.empty((2, c1), dtype=torch.float32, device=x.device)
        dy1 = torch.empty_like(y1)
        _lib.check(L.cvb_train_bn_silu_bwd(_p(dz1), 1, _p(y1), npix, c1, _p(st1), _p(g1f), _p(sums1), _p(dy1), _stream()), 'cvb_train_bn_silu_bwd')
        dw1 = conv_wgrad(x, dy1, 1)
        dx = conv(dy1, wb1, cin, 1) if ctx.need_dx else None
        return (dx, dw1, sums1[1].clone(), sums1[0].clone(), None, None, dw2, sums2[1].clone(), sums2[0].clone(), None, None, None, None)


class BaseConv(nn.Module):
    """src/models/modules/yolox_modules.py:35-55 (act='silu', groups=1; stride 1, or stride 2 with ksize 3 = the downsampling conv in front of
    every CSPLayer of the YOLOX backbone).  forward: NHWC bf16 in -> NHWC bf16 out."""

    def __init__(self, in_channels, out_channels, ksize, stride, groups=1, bias=False, act='silu'):
        super().__init__()
        if not (stride == 1 or (stride == 2 and ksize == 3)) or groups != 1 or bias or act != 'silu' or ksize not in (1, 3):
            raise NotImplementedError('B200 training BaseConv: ksize in {1,3}, stride 1 (or 2 with ksize 3), groups 1, no bias, SiLU')
        if in_channels % 64 or out_channels % 64:
            raise NotImplementedError('B200 training BaseConv: channel counts must be multiples of 64')
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=ksize, stride=stride, padding=(ksize - 1) // 2, groups=groups, bias=bias)
        self.bn = nn.BatchNorm2d(out_channels)
        self.act = nn.SiLU(inplace=True)  # (parameter-free; kept for module-tree parity with the reference)
        self.ksize = ksize
        self.stride = stride

    def forward(self, x):
        if not self.training:
            raise RuntimeError('cvpytorch_b200.train.BaseConv is the TRAINING drop-in (batch statistics); use the inference modules for eval()')
        bn = self.bn
        bn.num_batches_tracked += 1
        mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
        return _ConvBnSiLU.apply(x, self.conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, mom, self.ksize, self.stride)

    def forward_nchw(self, x):
        """Reference interface (NCHW in / out) for a stand-alone BaseConv, e.g. the stride-2 conv of a `dark` stage."""
        _check_cuda(x, 'BaseConv (B200 training)')
        return self.forward(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)).permute(0, 3, 1, 2).to(x.dtype)


class Bottleneck(nn.Module):
    """src/models/modules/yolox_modules.py:79-96 (depthwise=False)."""

    def __init__(self, in_channels, out_channels, shortcut=True, expansion=0.5, depthwise=False, act='silu'):
        super().__init__()
        if depthwise:
            raise NotImplementedError('depthwise bottlenecks are not on the B200 training path')
        hidden_channels = int(out_channels * expansion)
        self.conv1 = BaseConv(in_channels, hidden_channels, 1, stride=1, act=act)
        self.conv2 = BaseConv(hidden_channels, out_channels, 3, stride=1, act=act)
        self.use_add = shortcut and in_channels == out_channels

    def forward(self, x):
        if not self.training:
            raise RuntimeError('training drop-in: call .train()')
        c1, c2 = self.conv1, self.conv2
        for bn in (c1.bn, c2.bn):
            bn.num_batches_tracked += 1
        mom = c1.bn.momentum if c1.bn.momentum is not None else 1.0 / float(c1.bn.num_batches_tracked)
        y = _BottleneckChain.apply(x, c1.conv.weight, c1.bn.weight, c1.bn.bias, c1.bn.running_mean, c1.bn.running_var,
                                   c2.conv.weight, c2.bn.weight, c2.bn.bias, c2.bn.running_mean, c2.bn.running_var, c1.bn.eps, mom)
        if self.use_add:
            y = y + x
        return y


class CSPLayer(nn.Module):
    """src/models/modules/yolox_modules.py:99-129.  forward(x NCHW) -> NCHW (reference interface); `forward_nhwc` keeps the internal layout."""

    def __init__(self, in_channels, out_channels, n=1, shortcut=True, expansion=0.5, depthwise=False, act='silu'):
        super().__init__()
        hidden_channels = int(out_channels * expansion)
        self.conv1 = BaseConv(in_channels, hidden_channels, 1, stride=1, act=act)
        self.conv2 = BaseConv(in_channels, hidden_channels, 1, stride=1, act=act)
        self.conv3 = BaseConv(2 * hidden_channels, out_channels, 1, stride=1, act=act)
        self.m = nn.Sequential(*[Bottleneck(hidden_channels, hidden_channels, shortcut, 1.0, depthwise, act=act) for _ in range(n)])

    def forward_nhwc(self, x):
        x_1 = self.conv1(x)
        x_2 = self.conv2(x)
        x_1 = self.m(x_1)
        return self.conv3(torch.cat((x_1, x_2), dim=3))

    def forward(self, x):
        _check_cuda(x, 'CSPLayer (B200 training)')
        xh = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        return self.forward_nhwc(xh).permute(0, 3, 1, 2).to(x.dtype)
