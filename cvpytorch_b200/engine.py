"""Graph builder / executor for the fused B200 forward path.

A drop-in module "emits" its layers into a GraphBuilder: every activation is a channel-slice view (Val) of a
caller-owned split-NHWC buffer, every conv becomes a CvbConvPlan (TMA descriptors encoded once), and the
concats / upsamples of the reference graph disappear into buffer aliasing:

  * torch.cat((m(conv1(x)), conv2(x)))  (CSPLayer/C3)  -> conv1||conv2 run as ONE GEMM writing one buffer whose
    first half the bottleneck chain then updates in place; conv3 simply reads the whole buffer.
  * cat([up(x_conv), y]) -> fuse.cv1/cv2  (UpsamplingModule)  -> conv1x1(up(x)) == up(conv1x1(x)): a quarter-size
    fp32 partial GEMM on x_conv, added (nearest-upsampled) in the epilogue of the GEMM over y.
  * cat([down(x), lateral])  (DownsamplingModule)  -> both producers write into channel slices of one buffer.
  * SPPF cat([x, y1, y2, y3]) -> conv1 and the pooling kernel write the four slices of one buffer.

The executor is a flat list of steps; runs of conv plans are submitted with a single C call, and the whole
list can be captured into a CUDA graph (no step synchronises or allocates).
"""
import torch

from . import ops


class Val:
    """A channel-slice view of a split-NHWC activation buffer."""

    def __init__(self, tensor, c0=0, c=None):
        self.tensor = tensor
        self.c0 = c0
        self.c = tensor.C - c0 if c is None else c

    @property
    def B(self):
        return self.tensor.B

    @property
    def H(self):
        return self.tensor.H

    @property
    def W(self):
        return self.tensor.W

    def view(self):
        return self.tensor.view(self.c0, self.c)

    def slice(self, c0, c):
        assert c0 + c <= self.c
        return Val(self.tensor, self.c0 + c0, c)


class GraphBuilder:
    def __init__(self, batch, device='cuda'):
        self.B = batch
        self.device = device
        self.steps = []      # ('conv', ConvPlan) | ('fn', callable)
        self.buffers = []    # keep-alive
        self.n_convs = 0
        self.flops = 0       # 2*M*N*K of the reference graph (algorithmic, before the hi/lo x3)
        self._graph = None
        self.layer_log = []  # (name, cin, cout, k, s, H, W) for reports

    # ---------------------------------------------------------------- buffers
    def new_act(self, H, W, C):
        t = ops.SplitTensor(self.B, H, W, C, device=self.device)
        self.buffers.append(t)
        return Val(t)

    def new_f32(self, H, W, C):
        t = ops.F32Tensor(self.B, H, W, C, device=self.device)
        self.buffers.append(t)
        return t

    # ---------------------------------------------------------------- ops
    def conv(self, x, w64, b64, k, stride=1, pad=0, act='silu', out=None, residual=None, up_partial=None,
             f32_out=None, dilation=1, name='', block_n=0, w_window=0, residual_before_act=False, residual_scale=1.0):
        """x: Val.  w64: [O,I,k,k] float64 (BN already folded), b64: [O] float64.
        out: Val (split16) or None (allocate).  f32_out: F32Tensor -> plain fp32 output."""
        O, I = w64.shape[0], w64.shape[1]
        assert I == x.c, (name, I, x.c)
        Ho = (x.H + 2 * pad - dilation * (k - 1) - 1) // stride + 1
        Wo = (x.W + 2 * pad - dilation * (k - 1) - 1) // stride + 1
        if w_window:  # x is the zero-padded row-window layout (see include/cvb200.h)
            Wo = x.W - (w_window - 1)
            Ho = out.H if out is not None else x.H
            wp, bp = ops.pack_conv_weights(ops.window_weights(w64, w_window), b64, device=self.device)
        else:
            wp, bp = ops.pack_conv_weights(w64, b64, device=self.device)
        if f32_out is not None:
            out_view = f32_out.view(0, O)
            ret = f32_out
        else:
            if out is None:
                out = self.new_act(Ho, Wo, O)
            assert out.c == O and out.H == Ho and out.W == Wo, (name, out.c, O, out.H, Ho)
            out_view = out.view()
            ret = out
        plan = ops.ConvPlan(x.view(), out_view, wp, bp, k, stride, pad, dilation, act,
                            residual=residual.view() if residual is not None else None,
                            up_partial=up_partial.view(0, O) if up_partial is not None else None, block_n=block_n,
                            w_window=w_window, residual_before_act=1 if residual_before_act else 0, residual_scale=residual_scale)
        self.steps.append(('conv', plan))
        self.n_convs += 1
        self.flops += 2 * self.B * Ho * Wo * O * I * k * k
        self.layer_log.append((name, I, O, k, stride, Ho, Wo))
        return ret

    def conv_yolo_head(self, x, w64, b64, na, no, anchors_px, stride, z, z_rows, z_off, nms_ws=None, conf_thres=0.001, multi_label=True, name=''):
        """Detect-head 1x1 conv with the YOLOv5 decode as its epilogue (CVB_OUT_YOLO): writes this level's rows of z (and the NMS
        histogram / per-row best scores) straight from the accumulator; no raw [B,ny,nx,na*no] tensor exists."""
        O, I = w64.shape[0], w64.shape[1]
        assert I == x.c and O == na * no, (name, I, x.c, O)
        wp, bp = ops.pack_yolo_head_weights(w64, b64, na, no, device=self.device)
        out_view = ops.CvbView(z.data_ptr(), self.B, x.H, x.W, na * 128, na * 128, 0)
        y = ops.yolo_decode_desc(na, no, anchors_px, stride, z, z_rows, z_off, nms_ws, conf_thres, multi_label)
        plan = ops.ConvPlan(x.view(), out_view, wp, bp, 1, 1, 0, 1, None, yolo=y, keepalive=(z,))
        self.steps.append(('conv', plan))
        self.n_convs += 1
        self.flops += 2 * self.B * x.H * x.W * O * I
        self.layer_log.append((name, I, O, 1, 1, x.H, x.W))

    def fn(self, f):
        self.steps.append(('fn', f))

    # ---------------------------------------------------------------- execution
    def run(self):
        i, n = 0, len(self.steps)
        while i < n:
            kind, obj = self.steps[i]
            if kind == 'conv':
                j = i
                while j < n and self.steps[j][0] == 'conv':
                    j += 1
                ops.run_plans([s[1] for s in self.steps[i:j]])
                i = j
            else:
                obj()
                i += 1

    def capture(self, warmup=2, tail=None):
        """Captures run() into a CUDA graph (kernels are launched on torch's current stream by the C ABI).
        tail: optional callable appended to the captured step (e.g. the NCCL all-gather of the results: SURVEY.md 5 wants the
        collective inside the graph, not an eager launch after every replay)."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.run()
                if tail is not None:
                    tail()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run()
            if tail is not None:
                tail()
        self._graph = g
        return g

    def replay(self):
        if self._graph is None:
            self.run()
        else:
            self._graph.replay()
