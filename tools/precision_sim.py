"""Test-infrastructure tool (not on the product path): simulates cheaper operand formats on the CPU oracle to show why the
conv kernel keeps fp16 (hi, lo) pairs for BOTH operands.  Rounds conv inputs and/or weights of every layer of the calibrated
YOLOv5-s oracle to fp16 / bf16 and reports the end-to-end error at 1x3x640x640 against the fp32 oracle:
  fp16     activations and weights rounded to fp16 (what a single tensor-core pass computes)  -> logits 1.4e-2
  a16w32   only activations rounded                                                          -> 1.2e-2
  a32w16   only weights rounded                                                              -> 9.8e-3
  bf16     both rounded to bf16                                                              -> 1.4e-1
All are 10-100x outside the 1e-3 budget of BASELINE.json (the calibrated random network amplifies per-layer rounding ~28x),
whereas the three-product fp16 split measures 2e-5 on the GPU (profiles/r01/e2e_error_vs_reference_golden.txt)."""
import sys
sys.path.insert(0, '/root/repo')
import torch, torch.nn.functional as F
from oracle import yolov5_oracle as O
from cvpytorch_b200 import synth
torch.set_num_threads(8)
sd = synth.yolov5s_state_dict() if hasattr(synth,'yolov5s_state_dict') else None
torch.manual_seed(1029); x = torch.randn(1,3,640,640)
with torch.no_grad():
    ref = O.forward(x, sd)
z_ref = ref[0] if isinstance(ref,(tuple,list)) else ref
orig = F.conv2d
def mk(mode):
    def conv(x, w, b=None, *a, **k):
        if mode=='fp16':
            x = x.half().float(); w = w.half().float()
        elif mode=='bf16':
            x = x.bfloat16().float(); w = w.bfloat16().float()
        elif mode=='a16w32':   # activations fp16 only, weights hi+lo
            x = x.half().float()
        elif mode=='a32w16':
            w = w.half().float()
        return orig(x, w, b, *a, **k)
    return conv
for mode in ['fp16','a16w32','a32w16','bf16']:
    F.conv2d = mk(mode)
    with torch.no_grad():
        out = O.forward(x, sd)
    z = out[0] if isinstance(out,(tuple,list)) else out
    raw_r = ref[1]; raw = out[1]
    e = float((z-z_ref).abs().max()/z_ref.abs().max())
    er = max(float((a-b).abs().max()/b.abs().max()) for a,b in zip(raw, raw_r))
    # score part error in absolute terms
    es = float((z[...,4:]-z_ref[...,4:]).abs().max())
    print(mode, 'decoded z rel', f'{e:.2e}', 'raw logits rel', f'{er:.2e}', 'max abs score err', f'{es:.2e}')
F.conv2d = orig
