#!/usr/bin/env python
"""bench.py -- images/sec of the YOLOv5-s 640x640 bs64 forward hot path (backbone+neck+detect+decode+batched NMS).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                      (CPU arm: the reference's algorithm on the host cores)

One JSON line on rank 0.  `value`: inputs resident in HBM (device time, CUDA events, max over ranks).  `e2e`: the
same metric through the public pipeline API with pinned HOST buffers (H2D of every step's frames and D2H of every
step's detections inside the timed region).  `roofline`: the conv stack (the dominant kernel family), algorithmic
FLOPs / measured duration against the measured cuBLAS bf16 peak.  `cpu_baseline`: the oracle port on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'images/sec YOLOv5-s 640x640 bs64 forward (backbone+neck+detect+decode+NMS)'
ALG_GFLOP_PER_IMG = 16.43359375  # 1051.75 GFLOP / 64 (2*M*N*K over the 60 convs)


def env_int(k, d):
    return int(os.environ.get(k, d))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu}', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for l in self.lines:
            p = [x.strip() for x in l.split(',')]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(max(mx)) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arms
def synthetic_frames(batch, seed=1029):
    torch.manual_seed(seed)  # the trainer's own seed (trainer.py:55)
    return torch.randn(batch, 3, 640, 640)


def usable_cores():
    """Host threads this process may really use: affinity mask, capped by the cgroup CPU quota if one is set."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    for path in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
        try:
            txt = open(path).read().split()
            if path.endswith('cpu.max'):
                if txt[0] != 'max':
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    n = min(n, max(1, q // int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())))
        except Exception:
            pass
    return max(1, n)


def pick_cpu_threads(probe):
    """All the host threads the CPU arm can use *productively*: oversubscribed intra-op pools are slower than fewer
    threads, so time a small probe at n, n/2, n/4, ... and keep the fastest (reported as `cores`)."""
    n = usable_cores()
    cands = sorted({max(1, n >> k) for k in range(0, 4)} | {min(n, 32), min(n, 16)}, reverse=True)
    best, best_t = cands[0], float('inf')
    for c in cands:
        torch.set_num_threads(c)
        probe()
        t0 = time.perf_counter()
        probe()
        dt = time.perf_counter() - t0
        if dt < best_t * 0.95:
            best, best_t = c, dt
    return best


def cpu_reference_throughput(steps, warmup, batch=8):
    """The reference's algorithm (oracle port: torch CPU fp32 conv/BN/SiLU graph + numpy NMS) on all host cores."""
    from cvpytorch_b200 import synth
    from oracle import nms_oracle as NO
    from oracle import yolov5_oracle as YO
    sd = synth.yolov5s_state_dict(True)
    x = synthetic_frames(batch)
    cores = pick_cpu_threads(lambda: YO.forward(x[:2], sd))
    torch.set_num_threads(cores)

    def step():
        z, _ = YO.forward(x, sd)
        NO.non_max_suppression(z.numpy(), 0.001, 0.6, multi_label=True)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, cores, dt / steps, f'{steps} steps x bs{batch} 640x640 forward+decode+NMS, torch {torch.__version__} fp32, {cores} threads'


def run_reference_arm(args):
    rank = env_int('RANK', 0)
    if rank != 0:
        return
    steps = max(1, args.steps)
    v, cores, spt, sample = cpu_reference_throughput(steps, max(1, min(args.warmup, 2)))
    line = {'impl': 'reference', 'metric': METRIC, 'value': round(v, 3), 'unit': 'images/sec', 'n_gpus': args.gpus, 'steps': steps,
            'warmup': max(1, min(args.warmup, 2)), 'ms_per_step': round(spt * 1e3, 3), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'YOLOv5-s 640x640 forward+decode+NMS (conf/coco_yolov5_s.yml), CPU sample bs8 per step', 'parallelism': 'cpu'},
            'cpu_baseline': {'value': round(v, 3), 'unit': 'images/sec', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': round(v, 3), 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200_arm(args):
    import torch.distributed as dist
    from cvpytorch_b200 import _lib, synth
    from cvpytorch_b200.runtime import InferencePipeline
    world = env_int('WORLD_SIZE', 1)
    rank = env_int('RANK', 0)
    local = env_int('LOCAL_RANK', 0)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch
    if args.scaling == 'strong':
        # fixed global batch (args.batch images in total), contiguous shards per rank (global B=64 -> 8 img/GPU at W=8)
        if args.batch % world != 0:
            raise SystemExit('--scaling strong needs --batch divisible by the number of GPUs')
        B = args.batch // world
    model = synth.build_yolov5s(calibrated=True)
    K, W = args.steps, max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        return ms

    # ------------------------------------------------------------- device-resident arm ("value")
    G = model.build_graph(B, 640, 640, dev)
    g, ws = G['g'], G['ws']
    x_dev = synthetic_frames(B, seed=1029 + rank).to(dev)
    G['holder']['x'] = x_dev
    M = ws.max_det
    gathered = torch.empty((world * B, M * 7 + 1), dtype=torch.float32, device=dev)
    from cvpytorch_b200 import dist as cdist

    conv_segments = []  # (start_idx, end_idx) of consecutive conv steps -> event pairs
    steps_list = g.steps

    def run_step(events=None):
        """One pass of the hot path.  events: list to append (start, end) CUDA event pairs around the conv segments."""
        i, n = 0, len(steps_list)
        from cvpytorch_b200 import ops
        while i < n:
            kind, obj = steps_list[i]
            if kind == 'conv':
                j = i
                while j < n and steps_list[j][0] == 'conv':
                    j += 1
                if events is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                ops.run_plans([s[1] for s in steps_list[i:j]])
                if events is not None:
                    e1.record()
                    events.append((e0, e1))
                i = j
            else:
                obj()
                i += 1
        if world > 1:
            dist.all_gather_into_tensor(gathered, cdist.pack_detections(ws.det, ws.det_idx, ws.det_count))

    if args.graph:
        g.capture()

    def do_step(events=None):
        if args.graph:
            g.replay()
            if world > 1:
                dist.all_gather_into_tensor(gathered, cdist.pack_detections(ws.det, ws.det_idx, ws.det_count))
        else:
            run_step(events)

    for _ in range(W):
        do_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    ev = []
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for _ in range(K):
        do_step(ev if not args.graph else None)
    t_end.record()
    barrier()
    total_ms = max_over_ranks(t_start.elapsed_time(t_end))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / K
    value = world * B * K / (total_ms / 1e3)
    launches = _lib.launch_count() - n0
    conv_ms = None
    if ev:
        conv_ms = sum(a.elapsed_time(b) for a, b in ev) / K
    if args.graph:
        # graph replays do not pass through the C ABI: one eager pass (outside the timed region) counts the kernels of a step
        # and times the conv segments with CUDA events for the roofline
        K2 = min(K, 10)
        n1 = _lib.launch_count()
        ev2 = []
        for _ in range(K2):
            run_step(ev2)
        torch.cuda.synchronize()
        launches = (_lib.launch_count() - n1) // K2 * K
        conv_ms = sum(a.elapsed_time(b) for a, b in ev2) / K2
    overflow = int(ws.status[0].item())

    # ------------------------------------------------------------- end-to-end arm ("e2e"): pinned host in, host out
    def time_pipeline(pipe, hosts):
        for i in range(W):
            pipe.result(pipe.submit(hosts[i % 2]))
        barrier()
        smp = ClockSampler(local)
        if rank == 0:
            smp.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        slots = []
        for i in range(K):
            slots.append(pipe.submit(hosts[i % 2]))
            if i >= 1:
                pipe.result(slots[i - 1])  # host reads the previous step's detections while this one runs
        last = pipe.result(slots[-1])
        torch.cuda.current_stream().wait_event(pipe.ev_host[slots[-1]])  # chain the pipeline's last D2H into the timing stream
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        return ms, last, (smp.stop() if rank == 0 else None)

    pipe = InferencePipeline(model, B, 640, 640, dev, depth=2, use_cuda_graph=True)
    hosts = [synthetic_frames(B, seed=2000 + rank * 16 + i).pin_memory() for i in range(2)]
    e2e_ms, last, e2e_clk = time_pipeline(pipe, hosts)
    e2e_value = world * B * K / (e2e_ms / 1e3)
    kept_mean = float(last[2].float().mean())
    h2d_f32, d2h_f32 = pipe.h2d_bytes, pipe.d2h_bytes
    del pipe, hosts

    # same end-to-end loop fed with camera-side uint8 HWC frames (ToTensor + Normalize fused into the stem loader)
    pipe8 = InferencePipeline(model, B, 640, 640, dev, depth=2, use_cuda_graph=True, uint8_frames=True)
    def frames_u8(seed):
        # uint8 frames whose transformed values follow the same N(0,1) statistics as the fp32 arm (quantised to 1/255, clipped to [0,1]):
        # the NMS workload is data dependent, uniform byte noise would time a different candidate regime
        z = synthetic_frames(B, seed=seed)
        nm = model.input_norm
        m = torch.tensor(nm['mean'], dtype=torch.float32).view(1, 3, 1, 1)
        sd = torch.tensor(nm['std'], dtype=torch.float32).view(1, 3, 1, 1)
        u = ((z * sd + m) * 255.0).round_().clamp_(0, 255).to(torch.uint8)   # tensor (RGB, CHW) order
        return u.flip(1).permute(0, 2, 3, 1).contiguous().pin_memory()      # camera order: HWC, BGR

    hosts8 = [frames_u8(3000 + rank * 16 + i) for i in range(2)]
    e2e8_ms, _, e2e8_clk = time_pipeline(pipe8, hosts8)
    e2e8_value = world * B * K / (e2e8_ms / 1e3)
    h2d8, d2h8 = pipe8.h2d_bytes, pipe8.d2h_bytes

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak_tf = float(peaks.get('bf16_tflops_sustained', 1400.0))
        peak_src = 'measured bf16_tflops_sustained (MEASURED_PEAKS.json)' if peaks else 'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'
        roof = None
        if conv_ms:
            ach = ALG_GFLOP_PER_IMG * B / conv_ms  # GFLOP / ms == TFLOP/s
            # DRAM bytes of the conv launches of one bs64 step, from the committed ncu launch list (tools/launch_summary.py)
            traffic = None
            try:
                tj = json.load(open(os.path.join(ROOT, 'profiles', 'r01', 'conv_traffic.json')))
                if int(tj.get('batch', 0)) == B:
                    traffic = int(tj['conv_dram_bytes_per_step'])
            except Exception:
                pass
            # per-layer conv roofline: sum_i max(3 * flops_i / P_tensor, bytes_i / BW_hbm) with the stored 4 B/element
            hbm = float(peaks.get('hbm_gbs', 6500.0))
            bound_ms = 0.0
            for (_n, cin, cout, k, s_, Ho, Wo) in g.layer_log:
                fl = 2.0 * B * Ho * Wo * cout * cin * k * k
                by = 4.0 * (B * (Ho * s_) * (Wo * s_) * cin + B * Ho * Wo * cout) + 4.0 * cout * cin * k * k
                bound_ms += max(3 * fl / (peak_tf * 1e12), by / (hbm * 1e9)) * 1e3
            roof = {'bound': 'tensor', 'achieved': round(ach, 2), 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': round(ach / peak_tf, 4),
                    'traffic': traffic, 'kernel': 'conv_tc_kernel<*> (all %d fused conv launches of one step)' % g.n_convs,
                    'conv_ms_per_step': round(conv_ms, 4), 'peak_source': peak_src,
                    'algorithmic_bytes_per_step': int(7.80e9 * B / 64),
                    'per_layer_roofline': {'bound_ms': round(bound_ms, 4), 'frac': round(bound_ms / conv_ms, 4),
                                           'definition': 'sum over the conv layers of max(3*flops/P_tensor, bytes/BW_hbm): three fp16 MMA products per '
                                                         'fp32 product, activations stored as fp16 hi+lo (4 B/element); P = %.1f TF/s, BW = %.1f GB/s' % (peak_tf, hbm)},
                    'note': 'algorithmic FLOPs 2*M*N*K of the fp32 reference graph (1051.75 GFLOP/bs64); the kernel issues 3 fp16 MMA products per fp32 '
                            'product (hi/lo split, fp32-equivalent accuracy), so frac <= 1/3 by construction; traffic = dram read+write of the conv '
                            'launches of one step (ncu, profiles/r01); per-layer numbers in profiles/'}
        cpu_v, cores, spt, sample = cpu_reference_throughput(args.cpu_steps, 1) if (args.cpu_steps > 0 and world == 1) else (
            None, 0, 0, 'skipped (timed at N=1 only)' if world > 1 else 'skipped')
        line = {'metric': METRIC, 'value': round(value, 2), 'unit': 'images/sec', 'n_gpus': world, 'steps': K, 'warmup': W,
                'ms_per_step': round(ms_per_step, 4), 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
                'dtype': 'fp16x3-split (fp32-equivalent, fp32 accumulate)', 'data': 'synthetic',
                'config': {'workload': 'YOLOv5-s 640x640 forward+decode+NMS, bs64 per GPU (conf/coco_yolov5_s.yml; BASELINE.json configs[1])',
                           'global_batch': world * B, 'per_gpu_batch': B, 'parallelism': f'dp{world}', 'conf_thres': 0.001, 'iou_thres': 0.6,
                           'weights': 'synthetic, BN-calibrated (tests/golden/yolov5s_calib.npz)', 'cuda_graph': bool(args.graph),
                           'l2': 'per-step inputs 315 MB and activations > 126 MB L2 (no explicit flush needed)',
                           'collective': 'one all_gather_into_tensor of [B,300*7+1] f32 per step' if world > 1 else 'none (N=1)',
                           'kept_per_image_mean': kept_mean, 'nms_capacity_overflow': overflow},
                'gpu_launches': int(launches),
                'e2e': {'value': round(e2e_value, 2), 'unit': 'images/sec', 'h2d_bytes_per_step': h2d_f32, 'd2h_bytes_per_step': d2h_f32,
                        'ms_per_step': round(e2e_ms / K, 4), 'sm_mhz': (e2e_clk or {}).get('sm_mhz'),
                        'api': 'cvpytorch_b200.runtime.InferencePipeline.submit/result (pinned host fp32 frames in, host detections out)'},
                'e2e_uint8_frames': {'value': round(e2e8_value, 2), 'unit': 'images/sec', 'h2d_bytes_per_step': h2d8, 'd2h_bytes_per_step': d2h8,
                                     'ms_per_step': round(e2e8_ms / K, 4), 'sm_mhz': (e2e8_clk or {}).get('sm_mhz'),
                                     'api': 'InferencePipeline(uint8_frames=True): pinned host uint8 HWC frames in (ToTensor + Normalize fused into the stem '
                                            'loader, cvb_stem_s2d_u8), host detections out; an extension beyond the reference input contract'},
                'clocks': clocks, 'roofline': roof,
                'cpu_baseline': {'value': round(cpu_v, 3) if cpu_v else None, 'unit': 'images/sec', 'cores': cores, 'kind': 'port', 'sample': sample}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='images per GPU per step (BASELINE: 64)')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'], help='weak: --batch images per GPU (default); strong: --batch images in total')
    ap.add_argument('--graph', type=int, default=1, help='1 (default): the timed steps replay one captured CUDA graph (the conv-stack time for the roofline comes '
                    'from an extra eager pass); 0: eager launches with per-conv-segment events inside the timed region')
    ap.add_argument('--cpu-steps', type=int, default=3, help='bs8 CPU baseline steps timed on rank 0 (0 = skip)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)')
        run_b200_arm(args)


if __name__ == '__main__':
    main()
