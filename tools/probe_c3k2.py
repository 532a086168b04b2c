"""Developer probe: the opt-in whole-block C3k2 kernel (XRSEG_FUSE_C3K2=1) inside the network against the default path,
and the debug hook on the network's own tensors."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xr_image_segmentation_b200 import inference as I, weights as W
layers, ws = W.random_weights("n", 1, None)
model = I.Model(W.write_pack("n", layers, ws), "n")
fr = np.random.default_rng(0).integers(0, 256, (2, 640, 640, 3), dtype=np.uint8)
out = {}
for flag in ("0", "1"):
    os.environ["XRSEG_FUSE_C3K2"] = flag
    r = I.Runner(model, max_batch=2)
    r.schedule(fr); r.wait()
    out[flag] = {n: r.fetch(n) for n in ("b1", "b2.cv2", "b3")}
    print(flag, "counts", r.counts().tolist(), "launches", r.launch_count())
    r.close()
for n in ("b1", "b2.cv2", "b3"):
    a, b = out["0"][n], out["1"][n]
    print(n, a.shape, "max diff", float(np.abs(a - b).max()), "nan", int(np.isnan(b).sum()), "ref absmean", float(np.abs(a).mean()), "fused absmean", float(np.abs(b).mean()))
d = np.abs(out["0"]["b2.cv2"] - out["1"]["b2.cv2"])
print("per-channel max diff", d.max(axis=(0, 2, 3)).round(3).tolist())
print("rows with diff", np.nonzero(d.max(axis=(0, 1, 3)) > 0.05)[0][:20].tolist(), "cols", np.nonzero(d.max(axis=(0, 1, 2)) > 0.05)[0][:20].tolist())
# the debug hook on the network's own tensors and weights
idx = {l.name: i for i, l in enumerate(layers)}
g = lambda n: ws[idx[n]]
(w1, b1), (wm1, bm1), (wm2, bm2), (w2, b2) = g("b2.cv1"), g("b2.m0.cv1"), g("b2.m0.cv2"), g("b2.cv2")
print("shapes", w1.shape, wm1.shape, wm2.shape, w2.shape)
y = I.debug_c3k2(out["0"]["b1"], w1.reshape(w1.shape[0], -1), b1, wm1, bm1, wm2, bm2, w2.reshape(w2.shape[0], -1), b2)
print("hook vs unfused network: max diff", float(np.abs(y - out["0"]["b2.cv2"]).max()), " hook vs fused network:", float(np.abs(y - out["1"]["b2.cv2"]).max()))
