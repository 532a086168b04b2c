"""Developer probe (GPU box): per-role cycle counters of the TMA-fed tcgen05 conv kernel (halo 3x3, flat 1x1, s2) at the
network's own shapes (libxrseg_debug.so, PROBE instantiation: XRSEG_DBG_TIME=1; XRSEG_DBG_SKIP bits 1 no MMAs, 2 no stores,
4 no loads).   python tools/probe_tma.py [batch]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xr_image_segmentation_b200 import inference as I
rng = np.random.default_rng(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
#        name        cin cout  h    w   k  s
cases = [("b2.cv2", 48, 64, 160, 160, 1, 1), ("proto.cv2", 64, 64, 160, 160, 3, 1), ("b3", 64, 64, 160, 160, 3, 2)]
if os.environ.get("CASES") == "generic":  # layers on the generic epilogue path (chunk count does not divide the group count)
    cases = [("h3.cls.1pw", 80, 80, 80, 80, 1, 1), ("h3.box.0+coef.0", 64, 96, 80, 80, 3, 1), ("b9.cv2", 512, 256, 20, 20, 1, 1),
             ("n13.cv1", 384, 128, 40, 40, 1, 1), ("h4.box.0+coef.0", 128, 96, 40, 40, 3, 1)]
if os.environ.get("CASES") == "small":   # the latency-bound launches of the 20x20 / 40x40 stages
    cases = [("b8.m0.m0.cv2", 64, 64, 20, 20, 3, 1), ("b8.cv2", 384, 256, 20, 20, 1, 1), ("n13.m0.cv2", 32, 64, 40, 40, 3, 1), ("b10.ffn.1", 256, 128, 20, 20, 1, 1)]
os.environ["XRSEG_DBG_TIME"] = "1"
for name, cin, cout, h, w, k, s in cases:
    x = rng.standard_normal((B, cin, h, w), dtype=np.float32)
    wt = rng.standard_normal((cout, cin, k, k), dtype=np.float32) * np.float32(1 / np.sqrt(cin * k * k))
    b = rng.standard_normal(cout, dtype=np.float32)
    for skip in [int(x) for x in os.environ.get("SKIPS", "0,2,16,32,1").split(",")]:
        os.environ["XRSEG_DBG_SKIP"] = str(skip)
        print("case", name, (B, cin, cout, h, w, k, s), "skip", skip, file=sys.stderr, flush=True)
        I.debug_conv(x, wt, b, k, s, 1, variant=0)
