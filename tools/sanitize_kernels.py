"""One small launch of every kernel family (every conv_tc_kernel<BLOCK_N, BLOCK_K, OUT> instantiation the planner can select, both
activation loaders, residual / up-partial epilogues, the aux kernels, decode, the three NMS pipelines), for
  compute-sanitizer --tool memcheck  python tools/sanitize_kernels.py
  compute-sanitizer --tool racecheck python tools/sanitize_kernels.py
Shapes are tiny (the sanitizer serialises everything).  Results are compared loosely against torch only to make sure the launches ran."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from cvpytorch_b200 import models as M  # noqa: E402
from cvpytorch_b200 import ops  # noqa: E402
from oracle import nms_oracle as NO  # noqa: E402
from oracle import yolox_oracle as XO  # noqa: E402
import test_conv_gpu as T  # noqa: E402

torch.backends.cudnn.allow_tf32 = False  # the torch fp32 reference must not run on TF32 tensor cores
torch.backends.cuda.matmul.allow_tf32 = False

n = 0


def conv(*a, **k):
    global n
    err = T._run_conv(*a, **k)
    if err >= 2e-5:
        print('NUMERIC MISMATCH', a, k, err)
    n += 1


# (cin, cout, k, s, p, B, H, W): every (BLOCK_N, BLOCK_K) pair through cin / cout / block_n choices
for cin in (16, 32, 64):           # BLOCK_K 16 / 32 / 64
    for bn in (32, 64, 128, 256):
        conv(cin, max(bn, 32), 1, 1, 0, 1, 16, 16, block_n=bn)
        conv(cin, max(bn, 32), 1, 1, 0, 1, 16, 16, block_n=bn, act=None, f32_out=True)
for halo in (-1, 1, 2):
    conv(32, 32, 3, 1, 1, 1, 16, 24, halo=halo)
    conv(64, 64, 3, 1, 1, 1, 16, 24, halo=halo, residual=True)
    conv(64, 128, 3, 2, 1, 1, 16, 24, halo=halo)
    conv(128, 128, 3, 1, 1, 1, 16, 16, halo=halo, no_resident=1)
conv(128, 128, 1, 1, 0, 1, 16, 16, up=True)
conv(64, 64, 1, 1, 0, 2, 20, 20, residual=True)
print('conv launches ok:', n)

# aux kernels
x = torch.randn(1, 3, 64, 64, device='cuda')
t = ops.SplitTensor(1, 32, 35, 16)
ops.stem_s2d(x, t.view())
p = ops.SplitTensor(1, 20, 20, 256)
ops.nchw_to_split(torch.randn(1, 64, 20, 20, device='cuda'), p.view(0, 64))
ops.sppf_pool(p.view(0, 64), p.view(64, 64), p.view(128, 64), p.view(192, 64))
ops.split_to_nchw(p.view())

# YOLOv5 decode + NMS (both phases) on a small stress tensor
pred = torch.from_numpy(NO.make_stress_prediction(1, A=25200, regime='typical', seed=2)).cuda()
M.non_max_suppression(pred, 0.001, 0.6, multi_label=True)
pred = torch.from_numpy(NO.make_stress_prediction(1, A=25200, regime='capped', seed=3)).cuda()
M.non_max_suppression(pred, 0.001, 0.6, multi_label=True)
torch.cuda.synchronize()
print('yolo nms ok')

# YOLOX post-process
rec = XO.make_stress_records(regime='typical', seed=2)
ws = ops.YoloxWorkspace(1, rec.shape[0])
ops.yolox_nms(ws, 0.01, 0.65, cand=torch.from_numpy(rec[None]).cuda().contiguous())
torch.cuda.synchronize()
print('yolox nms ok')

# FCOS NMS
B, N = 1, 4096
g = torch.Generator().manual_seed(0)
sc = torch.rand(B, N, generator=g).cuda()
cl = torch.randint(1, 81, (B, N), generator=g, dtype=torch.int32).cuda()
xy = torch.rand(B, N, 2, generator=g) * 700
wh = torch.rand(B, N, 2, generator=g) * 90 + 4
bx = torch.cat([xy, xy + wh], 2).cuda().contiguous()
fw = ops.FcosWorkspace(B, N)
ops.fcos_nms(fw, 0.05, 0.6, scores=sc, classes=cl, boxes=bx)
torch.cuda.synchronize()
print('fcos nms ok')

# whole model at a tiny size (every graph step incl. the fused decode histogram path)
from cvpytorch_b200 import synth  # noqa: E402
m = synth.build_yolov5s(True)
m.predict(torch.randn(1, 3, 64, 64, device='cuda'))
torch.cuda.synchronize()
print('model ok')

# fused detect head: 1x1 conv with the YOLOv5 decode as its epilogue (CVB_OUT_YOLO; bulk-store and generic copy-out paths)
for (Bh, ny, nx, cinh, na, nc, zo) in ((1, 12, 20, 64, 3, 80, 4), (1, 7, 13, 32, 3, 80, 3), (1, 9, 16, 96, 2, 3, 8)):
    no = nc + 5
    tin = ops.SplitTensor(Bh, ny, nx, cinh)
    ops.nchw_to_split(torch.randn(Bh, cinh, ny, nx, device='cuda'), tin.view())
    wh_ = (torch.randn(na * no, cinh, 1, 1) * (2.5 / cinh ** 0.5)).double()
    bh_ = (torch.randn(na * no) * 0.5).double()
    A = na * ny * nx + zo + 7
    wsh = ops.NmsWorkspace(Bh, A, nc)
    ops.nms_reset(wsh)
    zh = torch.zeros(Bh, A, no, device='cuda')
    wy, by = ops.pack_yolo_head_weights(wh_, bh_, na, no)
    yd = ops.yolo_decode_desc(na, no, torch.tensor([[10., 13.], [16., 30.], [33., 23.]][:na]), 8.0, zh, A, zo, wsh, 0.05, True)
    ops.ConvPlan(tin.view(), ops.CvbView(zh.data_ptr(), Bh, ny, nx, na * 128, na * 128, 0), wy, by, 1, 1, 0, 1, None, yolo=yd).run()
    ops.yolo_nms(zh, wsh, 0.05, 0.6, True, hist_ready=True)
torch.cuda.synchronize()
print('fused head ok')

# training kernels: conv forward / backward-data (+ SiLU' epilogue) / backward-weight, BatchNorm + SiLU passes, one C3 step
from cvpytorch_b200 import train as TR  # noqa: E402
for (Bt, Ht, Wt, ci, co, kk) in ((1, 8, 16, 64, 64, 1), (2, 9, 7, 128, 64, 3), (1, 8, 8, 64, 128, 3)):
    xt = torch.randn(Bt, Ht, Wt, ci, device='cuda').to(torch.bfloat16)
    dyt = torch.randn(Bt, Ht, Wt, co, device='cuda').to(torch.bfloat16)
    wf_, wb_ = TR.pack_weights(torch.randn(co, ci, kk, kk, device='cuda') * 0.05)
    TR.conv(xt, wf_, co, kk)
    TR.conv(dyt, wb_, ci, kk)
    TR.conv_wgrad(xt, dyt, kk)
for (Bt, Ht, Wt, ci, co) in ((1, 9, 14, 64, 64), (2, 8, 8, 64, 128)):  # stride 2: parity maps, four backward-data sub-convolutions
    xt = torch.randn(Bt, Ht, Wt, ci, device='cuda').to(torch.bfloat16)
    wf_, wb_ = TR.pack_weights(torch.randn(co, ci, 3, 3, device='cuda') * 0.05)
    yt = TR.conv(xt, wf_, co, 3, stride=2)
    TR.conv_dgrad_s2(yt, wb_, ci, Ht, Wt)
    TR.conv_wgrad(xt, yt, 3, stride=2)
mt = TR.CSPLayer(128, 128, n=1).cuda().train()
xt = torch.randn(1, 128, 8, 8, device='cuda', requires_grad=True)
mt(xt).sum().backward()
torch.cuda.synchronize()
print('training kernels ok')
