"""ORACLE (test infrastructure, not product): shape-generic fp32 YOLO11-seg forward pass.

The shipped asset hard-codes batch 1 and the `n` widths in its reshape constants
(SURVEY.md "Hard parts").  This module restates the SAME topology -- SURVEY.md Appendix A,
i.e. the 499 chains of /root/reference/Assets/Resources/Model/yolo11n-seg-sentis.sentis --
for any batch and for the `n` / `s` width rules, so that random-init
networks of BASELINE.json configs 2/3 have a CPU reference.  `tests/test_oracle.py`
checks that, loaded with the asset's own weights, it reproduces the graph interpreter
(oracle/graph.py) on the reference's sample images.

Convolutions are consumed strictly in the asset's chain order (Appendix B order, the DFL
1x1 conv 405 excluded because its weights 0..15 are structural): that order is the
canonical weight order of the `XRSW` weight pack handed to the CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from . import postprocess as pp


@dataclass(frozen=True)
class Spec:
    name: str
    ch: tuple            # backbone widths at strides 2,4,8,16,32 after width scaling
    c3k2_mid: tuple      # output widths of the two early C3k2 (layers 2 and 4)
    cls_mid: int         # class-branch width  = max(ch[2], min(nc, 100))
    box_mid: int         # box-branch width    = max(16, ch[2] // 4, 64)
    coef_mid: int        # coef-branch width   = max(ch[2] // 4, 32)
    proto_mid: int       # Proto width         = 256 * width
    heads: int           # C2PSA heads         = (ch[4] // 2) // 64
    nc: int = 80
    nm: int = 32
    reg_max: int = 16


def make_spec(scale: str) -> Spec:
    width = {"n": 0.25, "s": 0.5}[scale]
    base = [64, 128, 256, 512, 1024]
    ch = tuple(int(min(c, 1024) * width) for c in base)           # n: 16,32,64,128,256
    c3k2_mid = (int(256 * width), int(512 * width))                # n: 64,128
    return Spec(scale, ch, c3k2_mid, cls_mid=max(ch[2], 80), box_mid=max(16, ch[2] // 4, 64),
                coef_mid=max(ch[2] // 4, 32), proto_mid=int(256 * width), heads=(ch[4] // 2) // 64)


def silu(x):
    return x * torch.sigmoid(x)


class _ConvFeed:
    """Hands out (w, b) pairs in canonical order and applies the convolution."""

    def __init__(self, weights=None, record=None, trace=None):
        self.weights = weights
        self.i = 0
        self.record = record  # list to append layer specs to (trace mode)
        self.trace = trace    # dict: layer name -> output tensor of that convolution (+ activation), for per-layer parity sweeps

    def __call__(self, x, cout, k=1, s=1, act=True, groups=1, transposed=False, name=""):
        cin = x.shape[1]
        if self.record is not None:
            self.record.append(dict(name=name, cin=cin, cout=cout, k=k, s=s, groups=groups, act=act,
                                    transposed=transposed, h_in=x.shape[2], w_in=x.shape[3]))
            if transposed:
                return x.new_empty((x.shape[0], cout, x.shape[2] * s, x.shape[3] * s))
            ho = (x.shape[2] + 2 * (k // 2) - k) // s + 1
            wo = (x.shape[3] + 2 * (k // 2) - k) // s + 1
            return x.new_empty((x.shape[0], cout, ho, wo))
        w, b = self.weights[self.i]
        self.i += 1
        w = torch.as_tensor(w)
        b = torch.as_tensor(b)
        if transposed:
            assert tuple(w.shape) == (cin, cout, k, k), (name, w.shape)
            y = F.conv_transpose2d(x, w, b, stride=s)
        else:
            assert tuple(w.shape) == (cout, cin // groups, k, k), (name, tuple(w.shape), (cout, cin // groups, k, k))
            y = F.conv2d(x, w, b, stride=s, padding=k // 2, groups=groups)
        y = silu(y) if act else y
        if self.trace is not None:
            self.trace[name] = y
            self.trace["in:" + name] = x
        return y


def _bottleneck(x, conv, c_mid, c_out, name):
    y = conv(x, c_mid, 3, name=name + ".cv1")
    y = conv(y, c_out, 3, name=name + ".cv2")
    return x + y


def _c3k2(x, conv, c_out, c_hidden, c3k, name):
    """C3k2 (chains 10-25 plain, 56-93 with the C3k inner block)."""
    y = conv(x, 2 * c_hidden, 1, name=name + ".cv1")
    a, b = y[:, :c_hidden], y[:, c_hidden:]
    if not c3k:
        m = _bottleneck(b, conv, c_hidden // 2, c_hidden, name + ".m0")
    else:
        c_ = c_hidden // 2
        t = conv(b, c_, 1, name=name + ".m0.cv1")
        t = _bottleneck(t, conv, c_, c_, name + ".m0.m0")
        t = _bottleneck(t, conv, c_, c_, name + ".m0.m1")
        u = conv(b, c_, 1, name=name + ".m0.cv2")
        m = conv(torch.cat([t, u], 1), c_hidden, 1, name=name + ".m0.cv3")
    return conv(torch.cat([a, b, m], 1), c_out, 1, name=name + ".cv2")


def _sppf(x, conv, c_out, name):
    c_ = x.shape[1] // 2
    y0 = conv(x, c_, 1, name=name + ".cv1")
    y1 = F.max_pool2d(y0, 5, 1, 2)
    y2 = F.max_pool2d(y1, 5, 1, 2)
    y3 = F.max_pool2d(y2, 5, 1, 2)
    return conv(torch.cat([y0, y1, y2, y3], 1), c_out, 1, name=name + ".cv2")


def _c2psa(x, conv, heads, name):
    """C2PSA (chains 152-190): 2-head (n) / 4-head (s) attention on the 20x20 map."""
    c_total = x.shape[1]
    c = c_total // 2
    y = conv(x, c_total, 1, name=name + ".cv1")
    a, b = y[:, :c], y[:, c:]
    B, _, H, W = b.shape
    N = H * W
    hd = c // heads
    kd = hd // 2
    qkv = conv(b, c + 2 * heads * kd, 1, act=False, name=name + ".attn.qkv")
    qkv = qkv.reshape(B, heads, 2 * kd + hd, N)
    q, k, v = qkv[:, :, :kd], qkv[:, :, kd:2 * kd], qkv[:, :, 2 * kd:]
    attn = torch.matmul(q.transpose(-2, -1), k) * float(np.float32(kd ** -0.5))
    attn = torch.softmax(attn, dim=-1)
    o = torch.matmul(v, attn.transpose(-2, -1)).reshape(B, c, H, W)
    pe = conv(v.reshape(B, c, H, W), c, 3, act=False, groups=c, name=name + ".attn.pe")
    o = conv(o + pe, c, 1, act=False, name=name + ".attn.proj")
    b = b + o
    f = conv(b, 2 * c, 1, name=name + ".ffn.0")
    f = conv(f, c, 1, act=False, name=name + ".ffn.1")
    b = b + f
    return conv(torch.cat([a, b], 1), c_total, 1, name=name + ".cv2")


def _up2(x):
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def _head_box_cls(x, conv, sp: Spec, name):
    b = conv(x, sp.box_mid, 3, name=name + ".box.0")
    b = conv(b, sp.box_mid, 3, name=name + ".box.1")
    b = conv(b, 4 * sp.reg_max, 1, act=False, name=name + ".box.2")
    c = conv(x, x.shape[1], 3, groups=x.shape[1], name=name + ".cls.0dw")
    c = conv(c, sp.cls_mid, 1, name=name + ".cls.0pw")
    c = conv(c, sp.cls_mid, 3, groups=sp.cls_mid, name=name + ".cls.1dw")
    c = conv(c, sp.cls_mid, 1, name=name + ".cls.1pw")
    c = conv(c, sp.nc, 1, act=False, name=name + ".cls.2")
    return b, c


def _head_coef(x, conv, sp: Spec, name):
    m = conv(x, sp.coef_mid, 3, name=name + ".coef.0")
    m = conv(m, sp.coef_mid, 3, name=name + ".coef.1")
    return conv(m, sp.nm, 1, act=False, name=name + ".coef.2")


def forward_raw(x: torch.Tensor, conv: _ConvFeed, sp: Spec, taps: dict | None = None):
    """x f32 [B,3,H,W] (0..1).  Returns raw head tensors in anchor-major layout:
    box_logits [B,A,64], cls_logits [B,A,80], coefs [B,A,32], protos [B,32,H/4,W/4]."""
    c1, c2, c3, c4, c5 = sp.ch
    t = conv(x, c1, 3, 2, name="b0")
    t = conv(t, c2, 3, 2, name="b1")
    t = _c3k2(t, conv, sp.c3k2_mid[0], sp.c3k2_mid[0] // 4, False, "b2")
    t = conv(t, sp.c3k2_mid[0], 3, 2, name="b3")
    f4 = _c3k2(t, conv, sp.c3k2_mid[1], sp.c3k2_mid[1] // 4, False, "b4")      # stride 8
    t = conv(f4, sp.c3k2_mid[1], 3, 2, name="b5")
    f6 = _c3k2(t, conv, sp.c3k2_mid[1], sp.c3k2_mid[1] // 2, True, "b6")       # stride 16
    t = conv(f6, c5, 3, 2, name="b7")
    t = _c3k2(t, conv, c5, c5 // 2, True, "b8")
    t = _sppf(t, conv, c5, "b9")
    f10 = _c2psa(t, conv, sp.heads, "b10")                                     # stride 32
    t = torch.cat([_up2(f10), f6], 1)
    f13 = _c3k2(t, conv, sp.c3k2_mid[1], sp.c3k2_mid[1] // 2, False, "n13")
    t = torch.cat([_up2(f13), f4], 1)
    p3 = _c3k2(t, conv, sp.c3k2_mid[0], sp.c3k2_mid[0] // 2, False, "n16")
    b3, k3 = _head_box_cls(p3, conv, sp, "h3")
    t = conv(p3, sp.c3k2_mid[0], 3, 2, name="n17")
    p4 = _c3k2(torch.cat([t, f13], 1), conv, sp.c3k2_mid[1], sp.c3k2_mid[1] // 2, False, "n19")
    b4, k4 = _head_box_cls(p4, conv, sp, "h4")
    t = conv(p4, sp.c3k2_mid[1], 3, 2, name="n20")
    p5 = _c3k2(torch.cat([t, f10], 1), conv, c5, c5 // 2, True, "n22")
    b5, k5 = _head_box_cls(p5, conv, sp, "h5")
    m3 = _head_coef(p3, conv, sp, "h3")
    m4 = _head_coef(p4, conv, sp, "h4")
    m5 = _head_coef(p5, conv, sp, "h5")
    pr = conv(p3, sp.proto_mid, 3, name="proto.cv1")
    pr = conv(pr, sp.proto_mid, 2, 2, act=False, transposed=True, name="proto.up")
    pr = conv(pr, sp.proto_mid, 3, name="proto.cv2")
    pr = conv(pr, sp.nm, 1, name="proto.cv3")

    def flat(ts):
        return torch.cat([z.flatten(2) for z in ts], 2).transpose(1, 2).contiguous()

    if taps is not None:
        taps.update(p3=p3, p4=p4, p5=p5, f10=f10)
    return dict(box_logits=flat([b3, b4, b5]), cls_logits=flat([k3, k4, k5]), coefs=flat([m3, m4, m5]),
                protos=pr, sizes=[tuple(z.shape[2:]) for z in (p3, p4, p5)])


def layer_table(scale: str, hw=(640, 640)) -> list[dict]:
    """Ordered list of every convolution (canonical weight order) with its shape."""
    rec: list[dict] = []
    sp = make_spec(scale)
    with torch.no_grad():
        forward_raw(torch.empty(1, 3, *hw, device="meta"), _ConvFeed(record=rec), sp)
    return rec


def weight_shape(l: dict) -> tuple:
    if l["transposed"]:
        return (l["cin"], l["cout"], l["k"], l["k"])
    return (l["cout"], l["cin"] // l["groups"], l["k"], l["k"])


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
def weights_from_sentis(model) -> list[tuple[np.ndarray, np.ndarray]]:
    """Dequantized (w, b) of every Conv / ConvTranspose chain in file order, DFL conv excluded.

    ↔ the 202 `DequantizeUint8` layers, w = (q - zp) * scale per tensor."""
    from .sentis import TensorValue

    deq = {}
    out = []
    for c in model.chains:
        if c.op == "DequantizeUint8":
            q = model.values[c.inputs[0]].data
            scale, zp = model.values[c.args[0]], model.values[c.args[1]]
            deq[c.outputs[0]] = ((q.astype(np.float32) - np.float32(zp)) * np.float32(scale)).astype(np.float32)
        elif c.op in ("Conv", "ConvTranspose"):
            w = deq[c.inputs[1]]
            if len(c.inputs) < 3 or c.inputs[2] < 0:
                continue  # DFL conv (chain 405) has no bias and structural weights
            out.append((w, deq[c.inputs[2]]))
    return out


CLS_FINAL = ("h3.cls.2", "h4.cls.2", "h5.cls.2")


def random_weights(scale: str, seed: int, cls_bias: float | None = None, gain: float = 1.5, bias_std: float = 0.2):
    """Random-init weights for BASELINE.json configs 2/3 -- the same frozen recipe as the product-side generator
    (xr_image_segmentation_b200/weights.py, checked equal by tests/test_abi.py): w ~ N(0, (1.5/sqrt(fan_in))^2),
    b ~ N(0, 0.2^2), final class-conv biases N(cls_bias, 0.05^2) with cls_bias -5.15 (n) / -12.4 (s; see weights.INIT_CLS_BIAS)."""
    cls_bias = {"n": -5.15, "s": -12.4}[scale] if cls_bias is None else cls_bias
    rng = np.random.default_rng(seed)
    out = []
    for l in layer_table(scale):
        shp = weight_shape(l)
        fan_in = (l["cin"] // l["groups"]) * l["k"] * l["k"] if not l["transposed"] else l["cin"]
        w = (rng.standard_normal(shp, dtype=np.float32) * np.float32(gain / np.sqrt(fan_in))).astype(np.float32)
        b = (rng.standard_normal(l["cout"], dtype=np.float32) * np.float32(bias_std)).astype(np.float32)
        if l["name"] in CLS_FINAL:
            b = (b * np.float32(0.25) + np.float32(cls_bias)).astype(np.float32)
        out.append((w, b))
    return out


# --------------------------------------------------------------------------------------
# full pipeline
# --------------------------------------------------------------------------------------
@torch.no_grad()
def run_raw(weights, images: torch.Tensor, scale: str = "n", taps=None, trace=None):
    """trace: optional dict filled with every convolution's output by layer name."""
    sp = make_spec(scale)
    feed = _ConvFeed(weights=weights, trace=trace)
    out = forward_raw(images.float(), feed, sp, taps)
    assert feed.i == len(weights), "unused weights"
    return out


def postprocess_frame(box_logits, cls_logits, coefs, protos, sizes, iou_thr=0.43, score_thr=0.301,
                      max_det=-1, strides=(8, 16, 32)):
    """One frame of the baked tail (chains 400-416, 455-498): numpy fp32 in, the four graph outputs out.

    box_logits [A,64], cls_logits [A,80], coefs [A,32], protos [32,P]."""
    ax, ay, st = pp.make_anchors(sizes, strides)
    boxes = pp.dfl_decode(box_logits, ax, ay, st)
    score, label = pp.class_scores(cls_logits)
    corners = pp.cxcywh_to_corners(boxes)
    keep = pp.nms_onnx(corners, score, iou_thr, score_thr, max_det)
    out_coefs = coefs[keep].astype(np.float32)
    return dict(keep=keep, boxes=boxes[keep], labels=label[keep], coefs=out_coefs,
                masks=pp.mask_probs(out_coefs, protos, _proto_hw(protos)),
                all_boxes=boxes, all_scores=score, all_labels=label, corners=corners)


def _proto_hw(protos):
    p = protos.shape[1]
    s = int(round(np.sqrt(p)))
    assert s * s == p
    return (s, s)


@torch.no_grad()
def run_model(weights, images: torch.Tensor, scale: str = "n", **kw):
    """images f32 [B,3,640,640] -> list of per-frame dicts (output_0..3 + intermediates)."""
    raw = run_raw(weights, images, scale)
    res = []
    B = images.shape[0]
    for i in range(B):
        protos = raw["protos"][i].reshape(raw["protos"].shape[1], -1).numpy()
        r = postprocess_frame(raw["box_logits"][i].numpy(), raw["cls_logits"][i].numpy(),
                              raw["coefs"][i].numpy(), protos, raw["sizes"], **kw)
        r["protos"] = protos
        res.append(r)
    return res, raw
