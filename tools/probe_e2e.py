"""Developer probe (GPU box): where the end-to-end step time goes.  H2D bandwidth, per-micro-batch schedule->wait time,
readback cost, and a two-runner ping-pong pipeline."""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))

import torch  # noqa: E402

from xr_image_segmentation_b200 import _lib, inference as I, weights as W  # noqa: E402

B = 64
lib = _lib.load_library()
layers, ws = W.random_weights("n", 1, None)
model = I.Model(W.write_pack("n", layers, ws), "n")
nbytes = B * 640 * 640 * 3
host = [lib.xrseg_host_alloc(nbytes) for _ in range(4)]
rng = np.random.default_rng(0)
for h in host:
    fr = rng.integers(0, 256, (B, 640, 640, 3), dtype=np.uint8)
    C.memmove(h, fr.ctypes.data, nbytes)

# H2D bandwidth
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
for chunk in (nbytes, nbytes // 4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        for o in range(0, nbytes, chunk):
            dev[o:o + chunk].copy_(src[o:o + chunk], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"H2D pinned chunk {chunk / 1e6:.1f} MB: {10 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")

for mb in (8, 16, 32, 64):
    r = I.Runner(model, device=0, max_batch=B, micro_batch=mb)
    for i in range(3):
        r.schedule_ptr(host[i % 4], B, 640, 640, 3)
        r.wait()
    t_s = t_w = t_r = 0.0
    N = 10
    t0 = time.perf_counter()
    for i in range(N):
        a = time.perf_counter()
        r.schedule_ptr(host[i % 4], B, 640, 640, 3)
        b = time.perf_counter()
        r.wait()
        c = time.perf_counter()
        r.readback(0)
        r.readback(1)
        r.masks(_lib.MASK_BITS_160)
        d = time.perf_counter()
        t_s += b - a
        t_w += c - b
        t_r += d - c
    tot = time.perf_counter() - t0
    print(f"mb={mb}: step {1e3 * tot / N:.2f} ms  (schedule call {1e3 * t_s / N:.2f}, wait {1e3 * t_w / N:.2f}, readbacks {1e3 * t_r / N:.2f})"
          f"  -> {B * N / tot:.0f} fps")
    # device-resident at this micro-batch
    for i in range(2):
        r.schedule_device(dev.data_ptr(), B, 640, 640, 3)
    r.sync()
    r.event_record(0)
    for i in range(N):
        r.schedule_device(dev.data_ptr(), B, 640, 640, 3)
    r.event_record(1)
    r.sync()
    print(f"   device-resident mb={mb}: {r.event_elapsed_ms(0, 1) / N:.2f} ms/step")
    r.close()

# two-runner ping-pong: H2D + compute of step i+1 overlap the readback of step i
for mb in (32, 64):
    rs = [I.Runner(model, device=0, max_batch=B, micro_batch=mb) for _ in range(2)]
    for i in range(4):
        rs[i % 2].schedule_ptr(host[i % 4], B, 640, 640, 3)
        rs[i % 2].wait()
    N = 20
    t0 = time.perf_counter()
    rs[0].schedule_ptr(host[0], B, 640, 640, 3)
    for i in range(1, N + 1):
        if i < N:
            rs[i % 2].schedule_ptr(host[i % 4], B, 640, 640, 3)
        p = rs[(i - 1) % 2]
        p.wait()
        p.readback(0)
        p.readback(1)
        p.masks(_lib.MASK_BITS_160)
    tot = time.perf_counter() - t0
    print(f"ping-pong mb={mb}: step {1e3 * tot / N:.2f} ms -> {B * N / tot:.0f} fps")
    for r in rs:
        r.close()
