import os, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xr_image_segmentation_b200 import inference as I
rng = np.random.default_rng(0)
def run(cin, cout, h, w, taps, variant=0):
    x = rng.standard_normal((1, cin, h, w), dtype=np.float32)
    wt = np.zeros((cout, cin, 3, 3), np.float32)
    for (kh, kw) in taps:
        wt[:, :, kh, kw] = rng.standard_normal((cout, cin), dtype=np.float32) / np.sqrt(cin)
    b = np.zeros(cout, np.float32)
    ref = F.conv2d(torch.from_numpy(x).half().float(), torch.from_numpy(wt).half().float(), torch.from_numpy(b), padding=1).numpy()
    y = I.debug_conv(x, wt, b, 3, 1, 0, variant=variant)
    err = np.abs(y - ref)
    print(f"cin={cin} cout={cout} {h}x{w} taps={taps} var={variant}: max err {err.max():.4f} (ref absmax {np.abs(ref).max():.2f}) bad frac {(err > 0.02).mean():.3f}", flush=True)
ALL = [(a, b) for a in range(3) for b in range(3)]
for mode in (0, 16, 32):
    for cin in (32, 64):
        for (h, w) in ((8, 8), (20, 20)):
            for taps in ([(1, 1)], [(1, 0)], [(2, 1)], ALL):
                run(cin, 16, h, w, taps, variant=mode)
