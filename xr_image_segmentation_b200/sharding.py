"""Frame-parallel sharding across the GPUs of one box.

Frames are independent on this path (no state crosses frames), so a batch is split into contiguous chunks, one runner
(= one process, one GPU, own stream / CUDA graph / pinned buffers) per chunk.  There is NO collective on the data path:
each rank copies its own variable-length detection list to its host; only if a single consumer wants the whole batch
are the (small) per-rank lists gathered host-side, in frame order.  `torch.distributed` is plumbing for that gather and
for the benchmark's barrier / max-over-ranks timing.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous split of `total` frames over `world` ranks -> (start, count); earlier ranks take the remainder."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


@dataclass
class Detections:
    """Detections of a contiguous range of frames (rows compacted in frame order, NMS order inside a frame)."""
    first_frame: int
    counts: np.ndarray                      # [frames] detections per frame
    boxes: np.ndarray                       # [N,4] cx,cy,w,h (output_0)
    labels: np.ndarray                      # [N]   (output_1)
    scores: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    mask_bits: np.ndarray | None = None     # [N,160,5] uint32, optional

    def validate(self):
        n = int(self.counts.sum())
        assert self.boxes.shape == (n, 4) and self.labels.shape == (n,)


def merge_detections(parts: list[Detections]) -> Detections:
    """Concatenate per-rank results in frame order (host side; ranks may arrive in any order)."""
    parts = sorted(parts, key=lambda p: p.first_frame)
    exp = parts[0].first_frame
    for p in parts:
        p.validate()
        if p.first_frame != exp:
            raise ValueError(f"frame ranges are not contiguous at frame {exp}")
        exp += len(p.counts)
    bits = None
    if all(p.mask_bits is not None for p in parts):
        bits = np.concatenate([p.mask_bits for p in parts], axis=0)
    return Detections(parts[0].first_frame, np.concatenate([p.counts for p in parts]),
                      np.concatenate([p.boxes for p in parts], axis=0), np.concatenate([p.labels for p in parts]),
                      np.concatenate([p.scores for p in parts]), bits)


def gather_to_rank0(local: Detections, group=None) -> Detections | None:
    """Host-side gather of the per-rank detection lists (works on gloo and nccl groups)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if world == 1:
        return local
    out = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object(local, out, dst=0, group=group)
    return merge_detections(out) if out is not None else None
