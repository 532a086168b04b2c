import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xr_image_segmentation_b200 import inference as I
rng = np.random.default_rng(0)
B = 16
cases = [(64, 64, 160, 160), (64, 64, 80, 80), (16, 16, 160, 160)]
os.environ["XRSEG_DBG_TIME"] = "1"
for cin, cout, h, w in cases:
    x = rng.standard_normal((B, cin, h, w), dtype=np.float32)
    wt = rng.standard_normal((cout, cin, 3, 3), dtype=np.float32) * np.float32(1 / np.sqrt(cin * 9))
    b = rng.standard_normal(cout, dtype=np.float32)
    for variant in (0,):
        for skip in (0, 1, 2, 7):
            os.environ["XRSEG_DBG_SKIP"] = str(skip)
            print("case", (cin, cout, h, w), "variant", variant, "skip", skip, flush=True)
            I.debug_conv(x, wt, b, 3, 1, 1, variant=variant)
