"""BASELINE.json configs[4] (post-process stress) on the GPU box: 8400 anchors x 80 classes, 300 planted objects x 3
overlapping anchors per frame, 32 x 160 x 160 prototypes, fp32 head tensors fed through xrseg_debug_post (the oracle-tensor
path: scalar fp32 mask kernel, IEEE sigmoid).  XRSEG_DBG_TIME makes the library time every launch with CUDA events.
   python tools/bench_post.py [batch]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["XRSEG_DBG_TIME"] = "1"
from xr_image_segmentation_b200 import _lib, inference as I, weights as W  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layers, ws = W.random_weights("n", 1, None)
r = I.Runner(I.Model(W.write_pack("n", layers, ws), "n"), max_batch=B, max_det=300)
rng = np.random.default_rng(5)
A = 8400
box = rng.standard_normal((B, A, 64)).astype(np.float32)
box.reshape(B, A, 4, 16)[..., 1] += 6.0
cls = (rng.standard_normal((B, A, 80)) - 6).astype(np.float32)
for f in range(B):
    for a in rng.choice(6400, 300, replace=False):
        for d in (0, 1, 80):
            if a + d < 6400:
                cls[f, a + d, rng.integers(0, 80)] = 2.0 + rng.standard_normal()
coef = rng.standard_normal((B, A, 32)).astype(np.float32)
proto = rng.standard_normal((B, 32, 25600)).astype(np.float32)
print("--- fp32 tensors, bit-exact kernels (xrseg_debug_post) ---", file=sys.stderr, flush=True)
r.debug_post(box, cls, coef, proto)
r.wait()
print("kept per frame:", r.counts()[:8].tolist(), flush=True)
print("--- same tensors rounded to fp16, product kernels (xrseg_debug_post_f16) ---", file=sys.stderr, flush=True)
r.debug_post(box, cls, coef, proto, f16=True)
r.wait()
print("kept per frame (fp16):", r.counts()[:8].tolist(), flush=True)
n = int(r.counts()[0])
for mode, name, per in ((_lib.MASK_BITS_160, "bit-packed 160x160", 3200), (_lib.MASK_UPSAMPLE_640, "640x640 u8", 409600)):
    r.masks(mode, first=0, count=n)              # first call grows the scratch buffer and faults the host pages in
    t0 = time.perf_counter()
    m = r.masks(mode, first=0, count=n)
    dt = time.perf_counter() - t0
    print(f"xrseg_masks {name}: frame 0, {n} masks, {n * per / 1e6:.1f} MB to the host in {dt * 1e3:.2f} ms (kernel + D2H)", flush=True)
r.close()
