import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xr_image_segmentation_b200 import inference as I
rng = np.random.default_rng(0)
B = 16
cin, cout, h, w = 64, 64, 160, 160
x = rng.standard_normal((B, cin, h, w), dtype=np.float32)
wt = rng.standard_normal((cout, cin, 3, 3), dtype=np.float32) * np.float32(1 / np.sqrt(cin * 9))
b = rng.standard_normal(cout, dtype=np.float32)
for skip in (7, 0):
    os.environ["XRSEG_DBG_SKIP"] = str(skip)
    I.debug_conv(x, wt, b, 3, 1, 1, variant=0)
