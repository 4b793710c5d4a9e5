"""Multi-GPU plumbing: one process per GPU, batch sharded by rank, ONE all-gather of the final detections.

The reference has no collective on its inference path (multi-GPU val runs independent DistributedSampler shards
and "reduces" with a no-op, trainer.py:227-231 / src/utils/distributed.py:122-125); its only gather precedent is
the pickle all_gather of the COCO evaluator (src/evaluator/eval_coco.py:464-506).  Images are independent through
conv, decode and NMS, so the shards need no data-path collective; the fixed-capacity result buffer
[B_local, max_det*6 + max_det + 1] (rows, candidate ids, count) is gathered with a single all_gather_into_tensor
(NCCL over NVLink on the box; gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world):
    """Contiguous shard [lo, hi) of rank; sizes differ by at most one."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_detections(det, det_idx, det_count):
    """det [B,M,6] f32, det_idx [B,M] i32, det_count [B] i32 -> one f32 tensor [B, M*7+1] (ints bit-cast)."""
    B, M, _ = det.shape
    return torch.cat([det.reshape(B, M * 6), det_idx.view(torch.float32).reshape(B, M),
                      det_count.view(torch.float32).reshape(B, 1)], 1).contiguous()


def unpack_detections(packed, max_det):
    B = packed.shape[0]
    M = max_det
    det = packed[:, :M * 6].reshape(B, M, 6)
    idx = packed[:, M * 6:M * 7].contiguous().view(torch.int32)
    cnt = packed[:, M * 7:].contiguous().view(torch.int32).reshape(B)
    return det, idx, cnt


def packed_len(B, max_det):
    """Length (floats) of the flat per-rank result buffer of ops.NmsWorkspace: [det B*M*6 | det_idx B*M | det_count B]."""
    return B * max_det * 7 + B


def all_gather_packed(packed, out, group=None):
    """The one collective of the data path: `packed` = ops.NmsWorkspace.packed of this rank (the NMS kernels wrote it, nothing is
    repacked), `out` = preallocated [world, len(packed)] on the same device.  Stream-ordered and CUDA-graph capturable (NCCL)."""
    dist.all_gather_into_tensor(out.view(-1), packed, group=group)
    return out


def unpack_gathered(out, B, max_det):
    """[world, B*M*7+B] (device or pinned host) -> (det [world*B,M,6] f32, det_idx [world*B,M] i32, det_count [world*B] i32), rank-major."""
    W, M = out.shape[0], max_det
    det = out[:, :B * M * 6].reshape(W * B, M, 6)
    idx = out[:, B * M * 6:B * M * 7].contiguous().view(torch.int32).reshape(W * B, M)
    cnt = out[:, B * M * 7:].contiguous().view(torch.int32).reshape(W * B)
    return det, idx, cnt


def all_gather_detections(det, det_idx, det_count, group=None, out=None):
    """Single collective: every rank ends up with the detections of the whole global batch (rank-major order).
    All ranks must hold the same local batch size (pad the last shard)."""
    packed = pack_detections(det, det_idx, det_count)
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return unpack_detections(out, det.shape[1])


def all_gather_fcos_detections(scores, classes, boxes, counts, group=None):
    """FCOS (BASELINE.json config 5, "NCCL box all-gather"): padded per-image results of FCOS.predict --
    scores [B,K] f32, classes [B,K] i32/i64, boxes [B,K,4] f32, counts [B] i32 -- packed into ONE f32 tensor [B, K*6+1]
    (ints bit-cast) and gathered with a single all_gather_into_tensor; returns the global (rank-major) tensors."""
    B, K = scores.shape
    packed = torch.cat([scores.reshape(B, K), classes.to(torch.int32).view(torch.float32).reshape(B, K), boxes.reshape(B, K * 4),
                        counts.to(torch.int32).view(torch.float32).reshape(B, 1)], 1).contiguous()
    world = dist.get_world_size(group)
    out = torch.empty((world * B, packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    G = out.shape[0]
    return (out[:, :K].contiguous(), out[:, K:2 * K].contiguous().view(torch.int32), out[:, 2 * K:6 * K].reshape(G, K, 4).contiguous(),
            out[:, 6 * K:].contiguous().view(torch.int32).reshape(G))


def all_gather_label_maps(labels, group=None, check=True):
    """Segmentation (BASELINE.json config 3): int64 label maps [B,H,W] (EncoderDecoder 'val' output) travel as uint8
    (class ids < 256; 2 MB per 1024x2048 image instead of 16 MB) in one all_gather_into_tensor and are widened to int64 again
    for API parity with encoder_decoder.py:133."""
    if check and int(labels.numel()) and (int(labels.max()) > 255 or int(labels.min()) < 0):  # check=False: no host sync (argmax of <= 256 classes)
        raise ValueError('label ids must fit uint8 for the packed gather')
    u8 = labels.to(torch.uint8).contiguous()
    world = dist.get_world_size(group)
    out = torch.empty((world * u8.shape[0],) + tuple(u8.shape[1:]), dtype=torch.uint8, device=u8.device)
    dist.all_gather_into_tensor(out, u8, group=group)
    return out.to(torch.int64)
