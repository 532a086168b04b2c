"""Generates the committed golden fixtures from the reference's own assets (run in the build container, where
/root/reference exists; the GPU box only sees the committed outputs).

  python tests/golden/make_golden.py

Writes, next to this file:
  yolo11n_seg.xrsw   the reference asset's 100 convolutions (uint8 tensors + per-tensor scale / zero point, exactly
                     the constants of yolo11n-seg-sentis.sentis) re-containered as an XRSW pack
  inputs.npz         sample frames of the reference (Assets/Resources/Images/*.jpg, bus.png) as raw uint8 RGB
  expected.npz       outputs of the ORACLE (oracle/graph.py = interpreter over the asset's 499 chains) on those frames:
                     the four graph outputs, NMS keep indices, strided samples of the head tensors, and the C#
                     post-processing results (ParseBoxes / DrawBoxes / DrawMask)
  labels.txt         the 80 class names (Assets/Resources/Model/yolo11n-labels.txt)
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import postprocess as pp  # noqa: E402
from oracle import preprocess as pre  # noqa: E402
from oracle.graph import (GraphInterpreter, V_COEF, V_CORNERS, V_HEAD_RAW, V_KEEP, V_PROTO, V_SCORES)  # noqa: E402
from oracle.sentis import load_sentis  # noqa: E402
from xr_image_segmentation_b200 import weights as W  # noqa: E402

REF = "/root/reference"
SENTIS = f"{REF}/Assets/Resources/Model/yolo11n-seg-sentis.sentis"
IMAGES = {
    "coco139": f"{REF}/Assets/Resources/Images/000000000139.jpg",
    "coco632": f"{REF}/Assets/Resources/Images/000000000632.jpg",
    "coco2006": f"{REF}/Assets/Resources/Images/000000002006.jpg",
    "coco4495": f"{REF}/Assets/Resources/Images/000000004495.jpg",
    "coco7108": f"{REF}/Assets/Resources/Images/000000007108.jpg",
    "bus": f"{REF}/bus.png",
}
SCREEN = (1920.0, 1080.0)   # a Screen.width/height for the C# box conventions


def quantized_tensors(model):
    """(Tensor8 w, Tensor8 b) per Conv/ConvTranspose chain in file order (DFL conv 405 has no bias: skipped)."""
    deq = {}
    out = []
    for c in model.chains:
        if c.op == "DequantizeUint8":
            deq[c.outputs[0]] = W.Tensor8(model.values[c.inputs[0]].data, model.values[c.args[0]], model.values[c.args[1]])
        elif c.op in ("Conv", "ConvTranspose"):
            if len(c.inputs) < 3 or c.inputs[2] < 0:
                continue
            out.append((deq[c.inputs[1]], deq[c.inputs[2]]))
    return out


def main():
    model = load_sentis(SENTIS)
    layers = W.layer_table("n")
    pack = W.write_pack("n", layers, quantized_tensors(model))
    open(os.path.join(HERE, "yolo11n_seg.xrsw"), "wb").write(pack)
    open(os.path.join(HERE, "labels.txt"), "w").write(open(f"{REF}/Assets/Resources/Model/yolo11n-labels.txt").read())

    gi = GraphInterpreter(model)
    inputs, exp = {}, {}
    for name, path in IMAGES.items():
        img = np.asarray(Image.open(path).convert("RGB"))
        inputs[name] = img
        x = torch.from_numpy(pre.to_tensor(img))
        out = gi.run(x, keep={V_HEAD_RAW, V_COEF, V_PROTO, V_KEEP, V_CORNERS, V_SCORES})
        boxes, labels = out[model.outputs[0]].numpy(), out[model.outputs[1]].numpy()
        coefs, masks = out[model.outputs[2]].numpy(), out[model.outputs[3]].numpy()
        exp[f"{name}.keep"] = out[V_KEEP].numpy().astype(np.int32)
        exp[f"{name}.boxes"] = boxes
        exp[f"{name}.labels"] = labels
        exp[f"{name}.coefs"] = coefs
        exp[f"{name}.mask_bits"] = np.packbits(masks > np.float32(0.5), axis=-1)
        exp[f"{name}.mask_prob_sample"] = masks[:, ::8, ::8].copy()
        head = out[V_HEAD_RAW][0].numpy()            # [144,8400]
        exp[f"{name}.head_sample"] = head[:, ::25].copy()
        exp[f"{name}.coef_sample"] = out[V_COEF][0].numpy()[:, ::25].copy()
        exp[f"{name}.proto_sample"] = out[V_PROTO].numpy()[:, ::64].copy()
        exp[f"{name}.scores_sample"] = out[V_SCORES][0, 0].numpy()[::5].copy()
        exp[f"{name}.input_sample"] = x[0, :, ::16, ::16].numpy().copy()
        # C# post-processing (IEE:529-559, IEB:37-81, IEM:82-119,232-247)
        pb, _ = pp.parse_boxes(boxes, labels, *SCREEN)
        db, _ = pp.draw_boxes(boxes, labels, *SCREEN)
        exp[f"{name}.parse_boxes"] = pb
        exp[f"{name}.draw_boxes"] = db
        n = min(len(db), 200)
        dm = np.stack([pp.draw_mask_bits(masks[i], db[i], int(SCREEN[0]), int(SCREEN[1])) for i in range(n)]) \
            if n else np.zeros((0, 160, 160), np.uint8)
        exp[f"{name}.draw_mask_bits"] = np.packbits(dm.astype(bool), axis=-1)
        print(name, img.shape, "keep", exp[f"{name}.keep"], "labels", labels)
    np.savez_compressed(os.path.join(HERE, "inputs.npz"), **inputs)
    np.savez_compressed(os.path.join(HERE, "expected.npz"), **exp)
    for f in ("yolo11n_seg.xrsw", "inputs.npz", "expected.npz", "labels.txt"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
