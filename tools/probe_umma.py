"""Developer probe (GPU box): which role bounds the thread-gather tcgen05 conv kernel.  Times one launch per case with
parts switched off (XRSEG_DBG_SKIP bits: 1 = no MMA issue, 2 = no epilogue math/stores, 4 = no A loads)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xr_image_segmentation_b200 import inference as I
rng = np.random.default_rng(0)
B = 16
#        cin cout  h    w   k  s
cases = [(32, 32, 160, 160, 1, 1), (16, 32, 320, 320, 3, 2), (64, 64, 160, 160, 3, 2), (48, 64, 160, 160, 1, 1),
         (128, 128, 80, 80, 1, 1)]
os.environ["XRSEG_DBG_TIME"] = "1"
for cin, cout, h, w, k, s in cases:
    x = rng.standard_normal((B, cin, h, w), dtype=np.float32)
    wt = rng.standard_normal((cout, cin, k, k), dtype=np.float32) * np.float32(1 / np.sqrt(cin * k * k))
    b = rng.standard_normal(cout, dtype=np.float32)
    for skip in (0, 1, 2, 4, 6, 7):
        os.environ["XRSEG_DBG_SKIP"] = str(skip)
        print("case", (cin, cout, h, w, k, s), "skip", skip, flush=True)
        I.debug_conv(x, wt, b, k, s, 1, variant=1)
