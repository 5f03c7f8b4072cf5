"""One-process-per-GPU sharding of the inference path.

Images are independent, so the path shards by contiguous batch slices exactly like the reference's
``split_render_data`` (yolo_modules/yolo_gluon.py:100-124: ``[int(i*B/N), int((i+1)*B/N))`` per context) with the
weights replicated; there is NO data-path collective ("replicas only").  The only exchange is an optional
gather of the tiny ``predict`` rows (<= 320 B per image) to every rank, done with torch.distributed (NCCL on GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(batch_size: int, rank: int, world: int):
    """Slice of the global batch owned by ``rank`` (split_render_data, yolo_gluon.py:117-118)."""
    return int(rank * batch_size / world), int((rank + 1) * batch_size / world)


def shard_batch(batch, rank: int, world: int):
    lo, hi = shard_bounds(len(batch), rank, world)
    return batch[lo:hi]


def gather_rows(local_rows, batch_size: int, group=None):
    """All-gather the per-rank ``predict`` rows back into global batch order -> np.float32 (batch_size, C)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    t = torch.as_tensor(np.asarray(local_rows, np.float32))
    if world == 1:
        return t.numpy()
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    C = t.shape[1]
    sizes = [shard_bounds(batch_size, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((cap, C), dtype=torch.float32, device=dev)
    buf[: t.shape[0]] = t.to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    rows = torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)
    assert rows.shape[0] == batch_size and sizes[rank][1] - sizes[rank][0] == t.shape[0]
    return rows.cpu().numpy()
