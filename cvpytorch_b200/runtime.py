"""Host-side inference runtime: pinned-host -> device -> fused graph -> (all-gather) -> pinned-host, double buffered.

The reference's loop is synchronous (trainer.py:156-157 ``imgs.cuda()`` then ``model(imgs, targets, 'val')`` then
``.cpu().numpy()`` per image, yolov5.py:267-284).  Here the H2D copy of batch i+1, the fused graph of batch i and
the D2H read of batch i-1 run on three streams; nothing on the path synchronises the host except ``result()``.
"""
import torch

from . import dist as cdist


class InferencePipeline:
    def __init__(self, model, batch, height, width, device, depth=2, gather_group=None, use_cuda_graph=True, uint8_frames=False,
                 graph_collective=True):
        """uint8_frames=False: submit() takes the reference's model input, pinned fp32 [B,3,H,W] (already normalised).
        uint8_frames=True: submit() takes pinned uint8 [B,H,W,3] camera frames; ToTensor + Normalize run in the stem loader and the
        host->device copy is 4x smaller."""
        self.model = model
        self.device = device
        self.depth = depth
        self.group = gather_group
        self.G = model.build_graph(batch, height, width, device, u8_input=True) if uint8_frames else model.build_graph(batch, height, width, device)
        self.g = self.G['g']
        self.ws = self.G['ws']
        if uint8_frames:
            self.x_dev = [torch.empty((batch, height, width, 3), dtype=torch.uint8, device=device) for _ in range(depth)]
        else:
            self.x_dev = [torch.empty((batch, 3, height, width), dtype=torch.float32, device=device) for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device)
        self.out_stream = torch.cuda.Stream(device)
        self.compute_stream = torch.cuda.Stream(device)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]        # input slot filled
        self.ev_free = [torch.cuda.Event() for _ in range(depth)]      # input slot consumed by the stem
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]      # results of slot ready on device
        self.ev_host = [torch.cuda.Event() for _ in range(depth)]      # results landed in pinned host memory
        M = self.ws.max_det
        world = torch.distributed.get_world_size(gather_group) if (gather_group is not None or (
            torch.distributed.is_available() and torch.distributed.is_initialized())) else 1
        self.world = world
        self.batch = batch
        L = cdist.packed_len(batch, M)
        self.res_dev = [torch.empty((world, L), dtype=torch.float32, device=device) for _ in range(depth)]
        self.res_host = [torch.empty((world, L), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.graphs = [None] * depth
        self.n_submitted = 0
        self.h2d_bytes = self.x_dev[0].numel() * self.x_dev[0].element_size()
        self.d2h_bytes = self.res_host[0].numel() * 4
        # the collective (or, on one GPU, the copy into the result slot) is part of the captured step: the NMS kernels write the wire
        # format (ops.NmsWorkspace.packed), so a step is ONE graph launch with no eager tail
        self.graph_tail = bool(use_cuda_graph and graph_collective)
        if world > 1 and self.graph_tail:  # NCCL must have created its communicator before a capture can record the collective
            with torch.cuda.stream(self.compute_stream):
                cdist.all_gather_packed(self.ws.packed, self.res_dev[0], gather_group)
            torch.cuda.synchronize(device)
        if use_cuda_graph:
            with torch.cuda.stream(self.compute_stream):
                # one CUDA graph per input slot; activation buffers are shared (the compute stream serialises them)
                for s in range(depth):
                    self.G['holder']['x'] = self.x_dev[s]
                    self.g._graph = None
                    tail = (lambda s=s: self._tail(s)) if self.graph_tail else None
                    self.graphs[s] = self.g.capture(warmup=1 if s else 2, tail=tail)
            torch.cuda.synchronize(device)

    def _tail(self, s):
        if self.world > 1:
            cdist.all_gather_packed(self.ws.packed, self.res_dev[s], self.group)
        else:
            self.res_dev[s].view(-1).copy_(self.ws.packed)

    def submit(self, x_host_pinned):
        """Enqueue one batch (pinned host fp32 [B,3,H,W]).  Returns the slot index."""
        s = self.n_submitted % self.depth
        if self.n_submitted >= self.depth:
            self.copy_stream.wait_event(self.ev_free[s])      # stem of the previous user of this slot has read it
        with torch.cuda.stream(self.copy_stream):
            self.x_dev[s].copy_(x_host_pinned, non_blocking=True)
            self.ev_in[s].record(self.copy_stream)
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.ev_in[s])
            if self.n_submitted >= self.depth:
                self.compute_stream.wait_event(self.ev_host[s])  # result slot s has been drained to the host
            if self.graphs[s] is not None:
                self.graphs[s].replay()
            else:
                self.G['holder']['x'] = self.x_dev[s]
                self.g.run()
            self.ev_free[s].record(self.compute_stream)
            if not (self.graphs[s] is not None and self.graph_tail):
                self._tail(s)
            self.ev_done[s].record(self.compute_stream)
        with torch.cuda.stream(self.out_stream):
            self.out_stream.wait_event(self.ev_done[s])
            self.res_host[s].copy_(self.res_dev[s], non_blocking=True)
            self.ev_host[s].record(self.out_stream)
        self.n_submitted += 1
        return s

    def result(self, slot):
        """Blocks until the slot's detections are in host memory; returns (det, idx, count) CPU tensors (views)."""
        self.ev_host[slot].synchronize()
        return cdist.unpack_gathered(self.res_host[slot], self.batch, self.ws.max_det)

    def drain(self):
        for s in range(min(self.depth, self.n_submitted)):
            self.ev_host[s].synchronize()


def bind_to_gpu_numa_node(device_index):
    """Pins this process (and therefore the pinned host buffers it allocates afterwards, first touch) to the CPUs of the NUMA node
    the GPU hangs off: with 8 ranks per box every rank otherwise streams its 315 MB/step input through whichever socket the OS
    picked (SCALE_r01: e2e efficiency 0.77 at 8 GPUs).  Best effort: returns the node or None when the topology is not exposed."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f'{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        node = int(open(f'/sys/bus/pci/devices/{bus}/numa_node').read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None
