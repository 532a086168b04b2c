"""Developer probe (GPU box): per-role cycle counters of the TMA-fed tcgen05 conv kernel (halo 3x3, flat 1x1, s2)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xr_image_segmentation_b200 import inference as I
rng = np.random.default_rng(0)
#        B  cin cout  h    w   k  s
cases = [(32, 16, 8, 160, 160, 3, 1), (32, 32, 32, 160, 160, 1, 1), (32, 48, 64, 160, 160, 1, 1)]
os.environ["XRSEG_DBG_TIME"] = "1"
for B, cin, cout, h, w, k, s in cases:
    x = rng.standard_normal((B, cin, h, w), dtype=np.float32)
    wt = rng.standard_normal((cout, cin, k, k), dtype=np.float32) * np.float32(1 / np.sqrt(cin * k * k))
    b = rng.standard_normal(cout, dtype=np.float32)
    for skip in (0, 2, 1):
        os.environ["XRSEG_DBG_SKIP"] = str(skip)
        print("case", (B, cin, cout, h, w, k, s), "skip", skip, flush=True)
        I.debug_conv(x, wt, b, k, s, 1, variant=0)
