"""ncu target: five representative convolutions of the network at batch 16 through the tcgen05 path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xr_image_segmentation_b200 import inference as I  # noqa: E402

rng = np.random.default_rng(0)
B = 16
CASES = [  # cin, cout, h, w, k, s   (layer)
    (64, 64, 160, 160, 3, 1),    # proto.cv2: halo mode, the largest layer
    (48, 64, 160, 160, 1, 1),    # b2.cv2: 1x1
    (16, 32, 320, 320, 3, 2),    # b1: stride-2 gather
    (16, 16, 160, 160, 3, 1),    # b2.m0 (padded 8): halo, tiny channels
    (256, 64, 80, 80, 1, 1),     # n16.cv1: 1x1, K = 256
]
for cin, cout, h, w, k, s in CASES:
    x = rng.standard_normal((B, cin, h, w), dtype=np.float32)
    wt = rng.standard_normal((cout, cin, k, k), dtype=np.float32) * np.float32(1 / np.sqrt(cin * k * k))
    b = rng.standard_normal(cout, dtype=np.float32)
    y = I.debug_conv(x, wt, b, k, s, 1)
    print(cin, cout, h, w, k, s, float(np.abs(y).mean()))
