"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the reference's YOLOX C3 block IN TRAINING MODE, forward and backward:
  BaseConv     src/models/modules/yolox_modules.py:35-55     act(bn(conv(x))), nn.Conv2d(bias=False), nn.BatchNorm2d with batch statistics, SiLU
  Bottleneck   src/models/modules/yolox_modules.py:79-96     conv2(conv1(x)) (+ x)
  CSPLayer     src/models/modules/yolox_modules.py:99-129    conv3(cat(m(conv1(x)), conv2(x)))
  backward     trainer.py:177-207                            loss.backward() through torch.autograd
written functionally over a state_dict with the reference's key names (F.conv2d / F.batch_norm(training=True) / F.silu) -- the backward is
torch.autograd's, exactly as in the reference.  Pinned against the reference's own modules by tools/make_golden_train.py ->
tests/golden/c3_train.npz (tests/test_oracle_golden.py::test_c3_train_oracle_reproduces_reference_fixture: forward, input gradient, every
parameter gradient and the running statistics, <= 1e-5)."""
import torch
import torch.nn.functional as F

BN_EPS = 1e-3        # src/models/yolox.py init: every BatchNorm2d gets eps = 1e-3, momentum = 0.03
BN_MOMENTUM = 0.03


def base_conv(x, sd, prefix, ksize, stride=1):
    """sd: name -> tensor (weights with requires_grad for the backward; running statistics are updated in place)."""
    y = F.conv2d(x, sd[prefix + '.conv.weight'], None, stride, (ksize - 1) // 2)
    y = F.batch_norm(y, sd[prefix + '.bn.running_mean'], sd[prefix + '.bn.running_var'], sd[prefix + '.bn.weight'], sd[prefix + '.bn.bias'], True,
                     BN_MOMENTUM, BN_EPS)
    return F.silu(y)


def bottleneck(x, sd, prefix, shortcut=True):
    y = base_conv(base_conv(x, sd, prefix + '.conv1', 1), sd, prefix + '.conv2', 3)
    return y + x if shortcut else y


def csp_layer(x, sd, n, prefix='', shortcut=True):
    p = prefix + '.' if prefix else ''
    x1 = base_conv(x, sd, p + 'conv1', 1)
    x2 = base_conv(x, sd, p + 'conv2', 1)
    for i in range(n):
        x1 = bottleneck(x1, sd, f'{p}m.{i}', shortcut)
    return base_conv(torch.cat((x1, x2), 1), sd, p + 'conv3', 1)


def train_step(x, G, sd_np, n):
    """One forward + backward with loss = sum(out * G).  sd_np: name -> numpy array.  Returns (y, dx, {param: grad}, {running stat: value})."""
    sd = {}
    for k, v in sd_np.items():
        t = torch.as_tensor(v).clone()
        if t.dtype.is_floating_point and 'running_' not in k:
            t.requires_grad_(True)
        sd[k] = t
    x = x.clone().requires_grad_(True)
    y = csp_layer(x, sd, n)
    (y * G).sum().backward()
    grads = {k: v.grad for k, v in sd.items() if v.requires_grad}
    stats = {k: v for k, v in sd.items() if 'running_' in k}
    return y.detach(), x.grad, grads, stats


def synthetic_state(cin, cout, n, seed=0):
    """Seeded parameters with the reference block's key names (bench.py's CPU leg; the GPU arm loads the same tensors)."""
    g = torch.Generator().manual_seed(seed)
    hid = cout // 2
    sd = {}

    def conv(prefix, ci, co, k):
        sd[prefix + '.conv.weight'] = (torch.randn(co, ci, k, k, generator=g) * (1.4 / (ci * k * k) ** 0.5)).numpy()
        sd[prefix + '.bn.weight'] = (torch.rand(co, generator=g) + 0.5).numpy()
        sd[prefix + '.bn.bias'] = (torch.randn(co, generator=g) * 0.3).numpy()
        sd[prefix + '.bn.running_mean'] = torch.zeros(co).numpy()
        sd[prefix + '.bn.running_var'] = torch.ones(co).numpy()
        sd[prefix + '.bn.num_batches_tracked'] = torch.zeros((), dtype=torch.long).numpy()
    conv('conv1', cin, hid, 1)
    conv('conv2', cin, hid, 1)
    conv('conv3', 2 * hid, cout, 1)
    for i in range(n):
        conv(f'm.{i}.conv1', hid, hid, 1)
        conv(f'm.{i}.conv2', hid, hid, 3)
    return sd


def dark_stage(x, sd, n):
    """One `dark` stage of the YOLOX CSPDarknet (src/models/backbones/det/csp_darknet.py: nn.Sequential(BaseConv(c, 2c, 3, 2), CSPLayer(2c, 2c, n))):
    keys '0.*' = the stride-2 BaseConv, '1.*' = the CSPLayer."""
    return csp_layer(base_conv(x, sd, '0', 3, 2), sd, n, prefix='1')
