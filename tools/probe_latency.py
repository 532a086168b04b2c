"""Where the batch-1 latency goes (BASELINE.json configs[3]): wall-clock split of one frame over the public calls.
   python tools/probe_latency.py"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xr_image_segmentation_b200 import inference as I, weights as W, _lib
lib = _lib.load_library()
layers, ws = W.random_weights("n", 1, None)
model = I.Model(W.write_pack("n", layers, ws), "n")
r1 = I.Runner(model, max_batch=1, resize_mode=_lib.RESIZE_LETTERBOX)
fb = 960 * 1280 * 3
hf = lib.xrseg_host_alloc(fb)
C.memmove(hf, np.random.default_rng(4).integers(0, 256, fb, dtype=np.uint8).ctypes.data, fb)
names = ["schedule", "wait", "readback boxes", "readback labels", "masks (bits)"]
T = []
for i in range(600):
    t = [time.perf_counter()]
    r1.schedule_ptr(hf, 1, 960, 1280, 3); t.append(time.perf_counter())
    r1.wait(); t.append(time.perf_counter())
    r1.readback(0); t.append(time.perf_counter())
    r1.readback(1); t.append(time.perf_counter())
    r1.masks(_lib.MASK_BITS_160); t.append(time.perf_counter())
    T.append(np.diff(t))
T = np.array(T[100:]) * 1e6
for n, v in zip(names, np.median(T, axis=0)):
    print(f"{n:18s} {v:7.1f} us")
print(f"{'total':18s} {np.median(T.sum(axis=1)):7.1f} us   detections {int(r1.counts()[0])}")
T2 = []
for i in range(600):
    t0 = time.perf_counter()
    r1.schedule_ptr(hf, 1, 960, 1280, 3)
    c, b, l, m = r1.collect(_lib.MASK_BITS_160)
    T2.append(time.perf_counter() - t0)
print(f"schedule + collect() {np.median(np.array(T2[100:])) * 1e6:7.1f} us (one call, one synchronisation)")
assert np.array_equal(b, r1.readback(0)) and np.array_equal(l, r1.readback(1)) and np.array_equal(m, r1.masks(_lib.MASK_BITS_160))
tm = r1.timings() if hasattr(r1, "timings") else None
print("device timings", tm)
# per-launch device time of the batch-1 pass (launches timed alone: an upper bound of what the graph + PDL overlap)
ops = r1.profile_ops(20)
tot = sum(o[1] for o in ops)
print(f"batch-1 launches {len(ops)}, sum of isolated launch times {tot * 1e3:.0f} us")
grp = {}
for name, ms, fl, by in ops:
    k = name.split(".")[0] if not name.startswith("post") else "post"
    grp[k] = grp.get(k, 0) + ms * 1e3
print({k: round(v, 1) for k, v in grp.items()})
print(sorted(((round(ms * 1e3, 1), n) for n, ms, _, _ in ops), reverse=True)[:12])
# device time of the whole batch-1 run (events around schedule .. done)
import torch
r1.event_record(0)
for _ in range(50):
    r1.schedule_ptr(hf, 1, 960, 1280, 3)
    r1.wait()
r1.event_record(1); r1.sync()
print(f"50 frames back to back: {r1.event_elapsed_ms(0, 1) / 50 * 1e3:.1f} us per frame on the stream (H2D + network + post)")
