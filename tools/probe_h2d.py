#!/usr/bin/env python
"""Platform ceiling of the end-to-end leg: N ranks (one per GPU) doing NOTHING but the host->device copy of one step's
frames (64 x 640 x 640 x 3 = 78.6 MB, pinned, one cudaMemcpyAsync per copy), all at the same time.  bench.py's `e2e` at N
GPUs cannot beat N x this number; SCALE e2e efficiency is to be read against it.

  python tools/probe_h2d.py                                                       (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/probe_h2d.py

Per rank: default pinned memory and write-combined pinned memory, with and without binding the rank to the CPUs next to its
GPU (nvmlDeviceSetCpuAffinity), 30 copies each after 5 warm-up copies, CUDA events on the copy stream.  Rank 0 prints one
JSON line per variant: per-rank GB/s (min / mean), aggregate GB/s, and the frames/s ceiling that implies."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

BYTES = 64 * 640 * 640 * 3
COPIES, WARM = 30, 5


def main():
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    rt = C.CDLL("libcudart.so.12")
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    rt.cudaFreeHost.argtypes = [C.c_void_p]
    dev = torch.empty(BYTES, dtype=torch.uint8, device=f"cuda:{local}")
    stream = torch.cuda.Stream()
    results = []
    for bind in (False, True):
        if bind:
            try:
                import pynvml
                pynvml.nvmlInit()
                pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            except Exception as e:                                    # noqa: BLE001
                if rank == 0:
                    print(f"# cpu affinity not available: {e}", file=sys.stderr)
        for flags, name in ((0, "pinned"), (4, "pinned write-combined")):
            host = C.c_void_p()
            assert rt.cudaHostAlloc(C.byref(host), BYTES, flags) == 0
            C.memset(host, 7, BYTES)                                  # touch the pages from this (possibly bound) thread
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                for i in range(WARM):
                    rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), host, BYTES, 1, C.c_void_p(stream.cuda_stream))
                stream.synchronize()
                if world > 1:
                    dist.barrier()
                e0.record(stream)
                for i in range(COPIES):
                    rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), host, BYTES, 1, C.c_void_p(stream.cuda_stream))
                e1.record(stream)
                stream.synchronize()
            gbs = COPIES * BYTES / (e0.elapsed_time(e1) * 1e-3) / 1e9
            rt.cudaFreeHost(host)
            t = torch.tensor([gbs], dtype=torch.float64)
            if world > 1:
                all_ = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
                dist.all_gather(all_, t)
                per = [float(x) for x in all_]
            else:
                per = [gbs]
            results.append({"n_gpus": world, "memory": name, "cpu_affinity": bind, "per_rank_gbs_min": min(per),
                            "per_rank_gbs_mean": sum(per) / len(per), "aggregate_gbs": sum(per),
                            "frames_per_s_ceiling": sum(per) * 1e9 / (640 * 640 * 3)})
    if rank == 0:
        cpus = os.cpu_count()
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")] if os.path.isdir("/sys/devices/system/node") else []
        for r in results:
            r.update({"host_cpus": cpus, "numa_nodes": len(nodes), "bytes_per_copy": BYTES})
            print(json.dumps(r))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
