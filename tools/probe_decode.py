import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from xr_image_segmentation_b200 import inference as I, weights as W
B = 64
layers, ws = W.random_weights("n", 1, None)
model = I.Model(W.write_pack("n", layers, ws), "n")
frames = np.random.default_rng(0).integers(0, 256, (B, 640, 640, 3), dtype=np.uint8)
dev = torch.from_numpy(frames.reshape(-1)).cuda()
for score in (0.0, 0.9999):
    r = I.Runner(model, device=0, max_batch=B, score=score)
    for _ in range(2):
        r.schedule_device(dev.data_ptr(), B, 640, 640, 3)
    r.wait()
    ops = {o[0]: o[1] * 1e3 for o in r.profile_ops(10)}
    print("score", score, "dets", int(r.counts().sum()), {k: round(v, 1) for k, v in ops.items() if k.startswith("post.")})
    r.close()
