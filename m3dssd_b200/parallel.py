"""Multi-GPU inference: one process per GPU, images sharded over ranks as independent batches.

The forward path has no cross-image dependency (eval-mode BN, per-image align / ANAB / NMS), so the
only exchange is the one BASELINE.json's config 5 names: an all-gather of the fixed-shape per-image
detection tensors ([local_batch, topk, 14] fp32 + counts) over NCCL/NVLink, followed by the batched
NMS over the gathered set.  Every rank holds the gathered set; the NMS of gathered image g runs on
rank g // local_batch (each image is suppressed exactly once per step instead of world_size times),
and a second, tiny all-gather ([local_batch, 40, 14] kept rows + counts) leaves the full result on
every rank.  The reference's equivalent is nn.DataParallel's scatter/replicate/gather per iteration
(lib/core.py:73-74, scripts/test_rpn_3d.py:50-51); here weights are replicated once.
"""
import torch
import torch.distributed as dist

from . import ops


def shard_range(global_batch, world_size, rank):
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def nms_slice(global_batch, world_size, rank):
    """Images of the gathered set whose NMS this rank runs (same contiguous split as shard_range)."""
    return shard_range(global_batch, world_size, rank)


def gather_detections(dets, det_num, group=None):
    """all-gather [b, topk, 14] detections and [b] counts from every rank, concatenated in rank order.
    Works on any backend (NCCL on GPUs; gloo in the CPU tests).  Equal local batches required."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return dets, det_num
    out_d = torch.empty((world * dets.shape[0],) + tuple(dets.shape[1:]), dtype=dets.dtype, device=dets.device)
    out_n = torch.empty((world * det_num.shape[0],), dtype=det_num.dtype, device=det_num.device)
    dist.all_gather_into_tensor(out_d, dets.contiguous(), group=group)
    dist.all_gather_into_tensor(out_n, det_num.contiguous(), group=group)
    return out_d, out_n


class ShardedDetector:
    """Per-rank engine + gather + NMS over the gathered detections (config 5)."""

    def __init__(self, net, local_batch, height, width, precision="bf16", use_graph=True):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.engine = net.engine(local_batch, height, width, precision=precision, use_graph=use_graph)
        e = self.engine
        gb = self.world * local_batch
        dev = e.dev
        self.num_keep = torch.zeros(gb, dtype=torch.int32, device=dev)
        self.kept = torch.zeros(gb, e.max_out, 14, dtype=torch.float32, device=dev)
        lb = local_batch
        self.keep_l = torch.zeros(lb, e.topk, dtype=torch.int32, device=dev)
        self.num_keep_l = torch.zeros(lb, dtype=torch.int32, device=dev)
        self.kept_l = torch.zeros(lb, e.max_out, 14, dtype=torch.float32, device=dev)
        self.nms_ws = torch.zeros(ops.nms_workspace_bytes(lb, e.topk), dtype=torch.uint8, device=dev)
        self.launches_per_step = e.launches_per_step("decode") + 3

    def _gather_and_nms(self):
        e = self.engine
        dets, num = gather_detections(e.dets, e.det_num)
        lo, hi = nms_slice(dets.shape[0], self.world, self.rank)
        mine, nmine = dets[lo:hi], num[lo:hi]  # this rank's share of the gathered set
        ops.nms_batched(mine, nmine, float(e.conf.nms_thres), self.nms_ws, self.keep_l, self.num_keep_l)
        ops.gather_kept(mine, self.keep_l, self.num_keep_l, e.max_out, self.kept_l)
        dist.all_gather_into_tensor(self.kept, self.kept_l)
        dist.all_gather_into_tensor(self.num_keep, self.num_keep_l)

    def step_pipelined(self, images=None):
        """Asynchronous step (see Engine.detect_pipelined): batch i's tail -- decode, [all-gather,] NMS --
        runs on the engine's side stream under batch i+1's trunk.  Results are valid once
        self.engine.tail_done has fired; returns the (kept, num_keep) buffers."""
        e = self.engine
        if self.world == 1:
            return e.detect_pipelined(images)
        e.detect_pipelined(images, between=self._gather_and_nms)
        return self.kept, self.num_keep

    def step(self, images=None):
        """Returns (kept [world*local_batch, max_out, 14], num_keep [world*local_batch])."""
        e = self.engine
        if self.world == 1:
            return e.detect(images)
        e.run(images, "decode")
        self._gather_and_nms()
        return self.kept, self.num_keep
