"""world_size-2 gloo test of the multi-GPU host logic: contiguous frame sharding + host-side gather of detections
(there is no collective on the data path)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from xr_image_segmentation_b200 import sharding as S


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = S.shard_range(total, world, rank)
    rng = np.random.default_rng(100 + rank)
    counts = np.array([(f * 7) % 4 for f in range(start, start + count)])
    n = int(counts.sum())
    frames = np.repeat(np.arange(start, start + count), counts)
    local = S.Detections(start, counts, np.tile(frames[:, None], (1, 4)).astype(np.float32), frames.astype(np.int32),
                         rng.uniform(size=n).astype(np.float32))
    merged = S.gather_to_rank0(local)
    dist.barrier()
    if rank == 0:
        q.put((merged.counts.tolist(), merged.labels.tolist(), merged.boxes[:, 0].tolist()))
    else:
        assert merged is None
    dist.destroy_process_group()


def test_two_rank_gather_in_frame_order():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    total = 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, labels, bx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_counts = [(f * 7) % 4 for f in range(total)]
    assert counts == exp_counts
    assert labels == [f for f in range(total) for _ in range(exp_counts[f])]     # rows in frame order
    assert bx == [float(v) for v in labels]
