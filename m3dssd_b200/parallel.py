"""Multi-GPU inference: one process per GPU, images sharded over ranks as independent batches.

The forward path has no cross-image dependency (eval-mode BN, per-image align / ANAB / NMS), so the
only exchange is the one BASELINE.json's config 5 names: an all-gather of the fixed-shape per-image
detection tensors ([local_batch, topk, 14] fp32 + counts) over NCCL/NVLink, and the batched NMS of the
gathered set.  Every rank ends up holding the gathered set; the NMS of gathered image g runs on rank
g // local_batch (each image is suppressed exactly once per step instead of world_size times) -- and
since that rank's share of the gathered set is exactly its own decode output, its NMS does not wait
for the collective (the gather runs asynchronously beside it).  A second, tiny all-gather
([local_batch, 40, 14] kept rows + counts) leaves the full result on every rank.  The reference's equivalent is nn.DataParallel's scatter/replicate/gather per iteration
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
    """Per-rank engine + all-gather of the detection tensors + NMS (config 5)."""

    def __init__(self, net, local_batch, height, width, precision="bf16", use_graph=True):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.engine = net.engine(local_batch, height, width, precision=precision, use_graph=use_graph)
        e = self.engine
        gb = self.world * local_batch
        dev = e.dev
        self.num_keep = torch.zeros(gb, dtype=torch.int32, device=dev)
        self.kept = torch.zeros(gb, e.max_out, 14, dtype=torch.float32, device=dev)
        # the gathered pre-NMS detection set of BASELINE config 5: [world * local_batch, topk, 14] + counts, on every rank
        self.dets_all = torch.zeros(gb, e.topk, 14, dtype=torch.float32, device=dev)
        self.det_num_all = torch.zeros(gb, dtype=torch.int32, device=dev)
        self.launches_per_step = e.launches_per_step("detect")  # our kernels only (NCCL's all-gather kernels not counted)

    def _gather_and_nms(self):
        """Runs on the engine's tail stream after decode.  The all-gather of the detection tensors is issued first,
        asynchronously (NCCL runs it on its own stream); the NMS of gathered image g belongs to rank g // local_batch,
        whose share of the gathered set IS its local decode output -- so the suppression starts at once on e.dets
        instead of waiting for the collective, and only the small all-gather of the kept rows is on the critical path.
        Both gathers are complete (stream-ordered) when this returns."""
        e = self.engine
        w1 = dist.all_gather_into_tensor(self.dets_all, e.dets, async_op=True)
        w2 = dist.all_gather_into_tensor(self.det_num_all, e.det_num, async_op=True)
        getattr(e, "_pipe", {}).get("nms", e._run_nms)()  # batched NMS + gather of this rank's images into e.kept / e.num_keep
        dist.all_gather_into_tensor(self.kept, e.kept)
        dist.all_gather_into_tensor(self.num_keep, e.num_keep)
        w1.wait()
        w2.wait()

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
