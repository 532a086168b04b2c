"""In-graph cost of groups of launches: serial batch-64 steps (one runner, CUDA graph + PDL + branch streams as in the product)
with the launches of a group left out (XRSEG_SKIP, libxrseg_debug.so only).  step(all) - step(without group) = what the
group costs inside the pipeline, as opposed to the isolated per-launch times of xrseg_profile_ops.
   python tools/probe_skip.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# every run also skips the post-processing ("post."): stale head tensors would otherwise flood the NMS with candidates
GROUPS = {
    "none": "post.",
    "b0 stem": "post.,b0",
    "b1..b3 (320/160 backbone)": "post.,b1,b2.,b3",
    "b4, n16 (80x80 backbone/neck)": "post.,b4.,n16.",
    "stage40 (b5, b6, n13, n17, n19, h4)": "post.,b5,b6.,n13.,n17,n19.,h4.",
    "stage20 (b7..b10, n20, n22, h5)": "post.,b7,b8.,b9.,b10.,c2psa,sppf,n20,n22.,h5.",
    "upsamples": "post.,upsample",
    "h3 box+coef": "post.,h3.box,h3.coef",
    "h3 cls": "post.,h3.cls",
    "proto": "post.,proto.",
    "proto.cv3": "post.,proto.cv3",
    "proto.up": "post.,proto.up",
    "all dw (cls.Xdw, pe)": "post.,h3.cls.0dw,h3.cls.1dw,h4.cls.0dw,h4.cls.1dw,h5.cls.0dw,h5.cls.1dw,b10.attn.pe",
    "everything but post": "post.,b,n,h,c2psa,sppf,upsample,proto",
}
CHILD = r'''
import os, sys, json
import numpy as np, torch
sys.path.insert(0, %r)
from xr_image_segmentation_b200 import inference as I, weights as W
B = 64
layers, ws = W.random_weights("n", 1, None)
model = I.Model(W.write_pack("n", layers, ws), "n")
n_r = int(os.environ.get("NRUN", "1"))
rs = [I.Runner(model, max_batch=B, debug=True) for _ in range(n_r)]
frames = np.random.default_rng(0).integers(0, 256, (B, 640, 640, 3), dtype=np.uint8)
dev = torch.from_numpy(frames.reshape(-1)).cuda()
for r in rs:
    for _ in range(3):
        r.schedule_device(dev.data_ptr(), B, 640, 640, 3)
    r.sync()
N = 40
rs[0].event_record(0)
for i in range(N):
    rs[i %% n_r].schedule_device(dev.data_ptr(), B, 640, 640, 3)
for r in rs:
    r.sync()
rs[0].event_record(1)
rs[0].sync()
print(json.dumps({"ms": rs[0].event_elapsed_ms(0, 1) / N, "launches": rs[0].launch_count()}))
''' % ROOT
base = {}
for nrun in (1, 4):
    for name, skip in GROUPS.items():
        env = dict(os.environ, XRSEG_SKIP=skip, NRUN=str(nrun))
        out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception:
            print(name, "FAILED", out.stderr[-400:])
            continue
        if name == "none":
            base[nrun] = d["ms"]
        print(f"runners {nrun}  without {name:40s} {d['ms']:.3f} ms/step  launches {d['launches']:3d}  -> group costs {base[nrun] - d['ms']:.3f} ms")
