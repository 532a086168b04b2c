"""ncu target: two passes of the whole hot path at BASELINE configs[1] (YOLO11n-seg, batch 64, synthetic frames,
random-init weights).  Pass 1 captures the CUDA graph; profile pass 2 with  --launch-skip N --launch-count N  where
N = the launch count this script prints (one stem launch + the graph's kernel nodes)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xr_image_segmentation_b200 import inference as I, weights as W  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layers, ws = W.random_weights("n", 1, None)
model = I.Model(W.write_pack("n", layers, ws), "n")
runner = I.Runner(model, device=0, max_batch=B)
frames = np.random.default_rng(0).integers(0, 256, (B, 640, 640, 3), dtype=np.uint8)
dev = torch.from_numpy(frames.reshape(-1)).cuda()
torch.cuda.synchronize()
for _ in range(2):
    runner.schedule_device(dev.data_ptr(), B, 640, 640, 3)
    runner.wait()
print("launches per pass", runner.launch_count(), "dets", int(runner.counts().sum()))
runner.close()
