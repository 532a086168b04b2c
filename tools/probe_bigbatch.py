"""One pass of the hot path at a large batch against the same frames at batch 8 (developer probe):
   python tools/probe_bigbatch.py <n|s> <batch> [graph 0/1]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xr_image_segmentation_b200 import inference as I, weights as W  # noqa: E402

scale, B = sys.argv[1], int(sys.argv[2])
graph = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
layers, ws = W.random_weights(scale, 3 if scale == "s" else 1, None)
model = I.Model(W.write_pack(scale, layers, ws), scale)
fr = np.random.default_rng(0).integers(0, 256, (8, 640, 640, 3), dtype=np.uint8)


def run(b):
    r = I.Runner(model, max_batch=b, use_cuda_graph=graph)
    dev = torch.from_numpy(np.tile(fr, (b // 8, 1, 1, 1)).reshape(-1)).cuda()
    r.schedule_device(dev.data_ptr(), b, 640, 640, 3)
    r.wait()
    c = r.counts().copy()
    cl = np.concatenate([r.fetch(f"cls_logits.{i}").reshape(b, 80, -1) for i in range(3)], axis=2)
    bl = np.concatenate([r.fetch(f"box_logits.{i}").reshape(b, 64, -1) for i in range(3)], axis=2)
    pr = r.fetch("protos").reshape(b, 32, -1)
    r.close()
    return c, cl, bl, pr


c8, cl8, bl8, pr8 = run(8)
c, cl, bl, pr = run(B)
print(scale, B, "dets/frame", float(c.mean()), "batch-8 counts", c8.tolist(), "last 8", c[-8:].tolist())
for name, a, ref in (("cls", cl, cl8), ("box", bl, bl8), ("proto", pr, pr8)):
    d = np.abs(a[-8:] - ref)
    print(f"  {name}: max |diff| vs batch 8 = {d.max():.4g}, mean {d.mean():.3g}, ref abs mean {np.abs(ref).mean():.3g}", flush=True)
