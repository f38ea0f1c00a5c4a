"""Multi-GPU plumbing: the hot path shards by image, so there is no data-path collective.

One process per GPU (torch.distributed, NCCL on GPUs / gloo in CPU tests).  Images -- and
with them their anchors, ROIs and label maps -- are split across ranks in contiguous blocks;
each rank runs the whole hot path on its own B200.  Collectives exist only around the path:
  * training: DDP bucketed gradient all-reduce (wrap_ddp),
  * evaluation: one all-gather of the fixed-shape per-image results (gather_detections).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise from the torchrun environment (RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", rank))
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def shard_range(n_items, rank, world):
    """Contiguous block of [0, n_items) owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rois(boxes, box_ind, n_images, rank, world):
    """Keep the ROIs whose image belongs to this rank; box_ind is rebased to the local batch."""
    lo, hi = shard_range(n_images, rank, world)
    sel = (box_ind >= lo) & (box_ind < hi)
    return boxes[sel], (box_ind[sel] - lo), (lo, hi)


def gather_detections(local, max_per_image=100, width=6):
    """All-gather per-image detections for evaluation.  `local`: list of [k_i,width] tensors (one
    per local image).  Every rank contributes a fixed-shape [n_local_max, max_per_image, width]
    block plus counts; returns the list for ALL images in global order on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = local[0].device if local else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    n_local = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    if world > 1:
        counts = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(counts, n_local)
        n_max = int(max(int(c.item()) for c in counts))
    else:
        counts, n_max = [n_local], len(local)
    block = torch.zeros((max(n_max, 1), max_per_image, width), dtype=torch.float32, device=dev)
    ks = torch.zeros(max(n_max, 1), dtype=torch.int64, device=dev)
    for i, d in enumerate(local):
        k = min(int(d.shape[0]), max_per_image)
        if k:
            block[i, :k] = d[:k].float()
        ks[i] = k
    if world > 1:
        blocks = [torch.zeros_like(block) for _ in range(world)]
        kss = [torch.zeros_like(ks) for _ in range(world)]
        dist.all_gather(blocks, block)
        dist.all_gather(kss, ks)
    else:
        blocks, kss = [block], [ks]
    out = []
    for r in range(world):
        for i in range(int(counts[r].item())):
            out.append(blocks[r][i, : int(kss[r][i].item())])
    return out


def wrap_ddp(module, device_ids=None, bucket_cap_mb=25):
    """DDP wrapper for the training config: gradients are all-reduced in buckets over NVLink,
    overlapped with backward.  No-op for a single process."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return module
    from torch.nn.parallel import DistributedDataParallel as DDP
    return DDP(module, device_ids=device_ids, bucket_cap_mb=bucket_cap_mb)


def max_over_ranks(value, device=None):
    """max of a python float across ranks (timing reduction)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
