"""Per-layer phase clocks of the chain kernel (CTA 0): XRSEG_CHAIN_PROBE=1, libxrseg_debug.so, graph off.
   python tools/probe_chain.py [batch]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["XRSEG_CHAIN_PROBE"] = "1"
from xr_image_segmentation_b200 import inference as I, weights as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layers, ws = W.random_weights("n", 1, None)
r = I.Runner(I.Model(W.write_pack("n", layers, ws), "n"), max_batch=B, debug=True, use_cuda_graph=False)
fr = np.random.default_rng(0).integers(0, 256, (B, 640, 640, 3), dtype=np.uint8)
for i in range(2):
    if i == 1:
        print("---- second pass ----", file=sys.stderr, flush=True)
    r.schedule(fr)
    r.wait()
r.close()
