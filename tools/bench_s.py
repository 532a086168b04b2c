"""Developer probe (GPU box): YOLO11s-seg shapes (BASELINE.json configs[2] per-GPU share: 64 frames) through the same path."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xr_image_segmentation_b200 import inference as I, weights as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layers, ws = W.random_weights("s", 3, None)
model = I.Model(W.write_pack("s", layers, ws), "s")
rs = [I.Runner(model, device=0, max_batch=B) for _ in range(2)]
frames = np.random.default_rng(2).integers(0, 256, (B, 640, 640, 3), dtype=np.uint8)
dev = torch.from_numpy(frames.reshape(-1)).cuda()
for r in rs:
    for _ in range(2):
        r.schedule_device(dev.data_ptr(), B, 640, 640, 3)
    r.wait()
print("dets/frame", float(rs[0].counts().mean()), "launches", rs[0].launch_count())
N = 20
for r in rs:
    r.sync()
rs[0].event_record(0)
for i in range(N):
    rs[i % 2].schedule_device(dev.data_ptr(), B, 640, 640, 3)
for r in rs:
    r.sync()
rs[0].event_record(1)
rs[0].sync()
ms = rs[0].event_elapsed_ms(0, 1) / N
print(f"YOLO11s-seg batch {B}: {ms:.3f} ms/step, {B / ms * 1e3:.0f} frames/s (2 streams)")
ops = rs[0].profile_ops(3)
tot = sum(o[1] for o in ops)
fl = sum(o[2] for o in ops)
print(f"sum of launches {tot:.3f} ms, {fl / 1e9 / B:.2f} GFLOP/frame, conv stack {fl / tot / 1e9:.0f} TFLOP/s")
for o in sorted(ops, key=lambda o: -o[1])[:8]:
    print(f"  {o[0]:16s} {o[1] * 1e3:7.1f} us")
