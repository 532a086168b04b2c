"""ORACLE (test infrastructure, not product): CPU restatement of the post-processing on the
reference's hot path -- the converter tail baked into the graph and the C# box/mask code.

Every function cites the reference lines it follows.  Arithmetic is numpy float32 with
one IEEE operation per numpy call (no FMA contraction) in the order written -- except the
mask dot product, which is an explicit FMA chain -- so that the CUDA kernels, written with
the matching __fadd_rn / __fmul_rn / __fdiv_rn / __fmaf_rn intrinsics, agree bit for bit.

PARITY UNPINNED by the reference (no tests / golden vectors exist); the
choices marked "oracle choice" are documented in DESIGN.md.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------------------
# Decode: graph chains 400-416 (DFL) and 455-465 (converter tail, CONV:63-73)
# --------------------------------------------------------------------------------------
def make_anchors(sizes=((80, 80), (40, 40), (20, 20)), strides=(8, 16, 32)):
    """Anchor centres (x+0.5, y+0.5) in grid units, row-major per scale, P3|P4|P5, and strides.

    ↔ the constants `anchors [1,2,8400]` / `strides [1,8400]` stored in the asset
    (SURVEY.md §8 a8, Appendix C "Notable constants")."""
    ax, ay, st = [], [], []
    for (h, w), s in zip(sizes, strides):
        ys, xs = np.meshgrid(np.arange(h, dtype=f32) + f32(0.5), np.arange(w, dtype=f32) + f32(0.5), indexing="ij")
        ax.append(xs.reshape(-1))
        ay.append(ys.reshape(-1))
        st.append(np.full(h * w, s, f32))
    return np.concatenate(ax), np.concatenate(ay), np.concatenate(st)


def dfl_decode(box_logits: np.ndarray, ax, ay, stride) -> np.ndarray:
    """box_logits f32 [A,64] (4 sides x 16 bins, side-major) -> [A,4] cx,cy,w,h in input pixels.

    Follows graph chains 401-415: softmax over the 16 bins, expectation with weights 0..15
    (1x1 conv 405), lt = d[0:2], rb = d[2:4], x1y1 = anchor - lt, x2y2 = anchor + rb,
    c = (x1y1 + x2y2) * 0.5, wh = x2y2 - x1y1, all * stride."""
    a = box_logits.astype(f32).reshape(-1, 4, 16)
    m = a.max(axis=2, keepdims=True)
    e = np.exp((a - m).astype(f32)).astype(f32)
    p = (e / e.sum(axis=2, keepdims=True, dtype=f32)).astype(f32)
    d = np.zeros(p.shape[:2], f32)
    for k in range(16):                      # sequential expectation, fp32
        d = (d + p[:, :, k] * f32(k)).astype(f32)
    x1 = ax - d[:, 0]
    y1 = ay - d[:, 1]
    x2 = ax + d[:, 2]
    y2 = ay + d[:, 3]
    cx = (x1 + x2) * f32(0.5)
    cy = (y1 + y2) * f32(0.5)
    w = x2 - x1
    h = y2 - y1
    return (np.stack([cx, cy, w, h], axis=1) * stride[:, None]).astype(f32)


def class_scores(cls_logits: np.ndarray):
    """cls_logits f32 [A,80] -> (score [A] f32, label [A] i32).

    Sigmoid (chain 416), ReduceMax (464, CONV:69), ArgMax first-max-wins (471, CONV:70)."""
    prob = (f32(1) / (f32(1) + np.exp(-cls_logits.astype(f32)).astype(f32))).astype(f32)
    score = prob.max(axis=1)
    label = np.argmax(prob == score[:, None], axis=1).astype(np.int32)
    return score, label


def cxcywh_to_corners(b: np.ndarray) -> np.ndarray:
    """[A,4] cx,cy,w,h -> [A,4] x1,y1,x2,y2; ↔ MatMul with the 4x4 map, CONV:40-46,73 (chain 459)."""
    cx, cy, w, h = (b[:, i].astype(f32) for i in range(4))
    hw = w * f32(0.5)
    hh = h * f32(0.5)
    return np.stack([cx - hw, cy - hh, cx + hw, cy + hh], axis=1).astype(f32)


# --------------------------------------------------------------------------------------
# NMS: graph chain 466 (`Functional.NMS(corners, scores, iou, score)`, CONV:76)
# --------------------------------------------------------------------------------------
def iou_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """IoU of one box `a` [4] against boxes `b` [K,4]; corners x1,y1,x2,y2; fp32, no FMA.

    oracle choice (the ONNX NonMaxSuppression definition the Inference Engine op mirrors):
    inter = max(0, min(x2) - max(x1)) * max(0, min(y2) - max(y1));
    iou = inter / (areaA + areaB - inter)."""
    a = a.astype(f32)
    b = b.astype(f32)
    iw = np.maximum(f32(0), np.minimum(a[2], b[:, 2]) - np.maximum(a[0], b[:, 0]))
    ih = np.maximum(f32(0), np.minimum(a[3], b[:, 3]) - np.maximum(a[1], b[:, 1]))
    inter = (iw * ih).astype(f32)
    area_a = ((a[2] - a[0]) * (a[3] - a[1])).astype(f32)
    area_b = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(f32)
    union = ((area_a + area_b).astype(f32) - inter).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / union).astype(f32)


def nms_onnx(corners: np.ndarray, scores: np.ndarray, iou_thr: float, score_thr: float,
             max_out: int = -1) -> np.ndarray:
    """Class-agnostic greedy NMS, returns kept anchor indices in selection (score-descending) order.

    candidates: score > score_thr; order: score descending, index ascending on ties (stable);
    a candidate is dropped if IoU with any already-kept box is > iou_thr (strict).
    `max_out` < 0 = unlimited (asset constant maxOutputBoxesPerClass = -1)."""
    corners = np.asarray(corners, f32)
    scores = np.asarray(scores, f32)
    cand = np.nonzero(scores > f32(score_thr))[0]
    order = cand[np.argsort(-scores[cand], kind="stable")]
    keep: list[int] = []
    thr = f32(iou_thr)
    kept_boxes = np.zeros((0, 4), f32)
    for i in order:
        if len(keep):
            iou = iou_f32(corners[i], kept_boxes)
            if np.any(iou > thr):
                continue
        keep.append(int(i))
        kept_boxes = np.concatenate([kept_boxes, corners[i:i + 1]], axis=0)
        if 0 <= max_out == len(keep):
            break
    return np.asarray(keep, np.int32)


# --------------------------------------------------------------------------------------
# Mask assembly: graph chains 496-498 (CONV:87-97)
# --------------------------------------------------------------------------------------
def mask_logits(coefs: np.ndarray, protos: np.ndarray) -> np.ndarray:
    """coefs f32 [N,32] x protos f32 [32,P] -> [N,P]: acc = fma(coef[:,k], proto[k,:], acc) for k = 0..31 in fp32.

    oracle choice: the summation order of the reference's MatMul (chain 496) is not observable; a sequential fused
    multiply-add chain is what a GPU (and Burst on ARM) emits for a 32-term dot product and what the CUDA kernel
    reproduces exactly.  The fp32 FMA is emulated in float64: the product of two fp32 values is exact in float64 and
    the single float64 add is rounded to fp32 (double rounding can differ in the last bit with probability ~2^-29
    per operation, far below anything the threshold tests can see)."""
    c64 = coefs.astype(np.float32).astype(np.float64)
    p64 = protos.astype(np.float32).astype(np.float64)
    acc = np.zeros((coefs.shape[0], protos.shape[1]), f32)
    for k in range(coefs.shape[1]):
        acc = (acc.astype(np.float64) + c64[:, k:k + 1] * p64[k:k + 1, :]).astype(f32)
    return acc


def sigmoid_f32(x: np.ndarray) -> np.ndarray:
    x = x.astype(f32)
    return (f32(1) / (f32(1) + np.exp(-x).astype(f32))).astype(f32)


def mask_probs(coefs, protos, hw=(160, 160)) -> np.ndarray:
    """-> output_3 f32 [N,160,160] (Sigmoid 497 + Reshape 498)."""
    return sigmoid_f32(mask_logits(coefs, protos)).reshape(-1, *hw)


# --------------------------------------------------------------------------------------
# C# box post-processing: IEExecutor.ParseBoxes (IEE:529-559), IEBoxer.DrawBoxes (IEB:37-81)
# --------------------------------------------------------------------------------------
def load_labels(text: str) -> list[str]:
    """↔ IEBoxer.Start: Split on '\\n','\\r', RemoveEmptyEntries (IEB:33)."""
    return [s for s in text.replace("\r", "\n").split("\n") if s != ""]


def get_class_name(labels: list[str], label_id: int) -> str:
    """↔ IEBoxer.GetClassName (IEB:183-188)."""
    if labels is None or label_id < 0 or label_id >= len(labels):
        return "unknown"
    return labels[label_id].replace(" ", "_")


def parse_boxes(boxes: np.ndarray, label_ids: np.ndarray, screen_w: float, screen_h: float, cap: int = 50):
    """↔ IEExecutor.ParseBoxes (IEE:529-559): centred, Y-up screen coordinates, first `cap` rows.

    Returns f32 [n,4] (CenterX, CenterY, Width, Height) and the label ids of those rows."""
    sx = f32(screen_w) / f32(640)
    sy = f32(screen_h) / f32(640)
    n = min(boxes.shape[0], cap)
    b = boxes[:n].astype(f32)
    out = np.empty((n, 4), f32)
    out[:, 0] = (b[:, 0] - f32(320)) * sx
    out[:, 1] = (f32(320) - b[:, 1]) * sy
    out[:, 2] = b[:, 2] * sx
    out[:, 3] = b[:, 3] * sy
    return out, np.asarray(label_ids[:n], np.int32)


def draw_boxes(boxes: np.ndarray, label_ids: np.ndarray, image_w: float, image_h: float, cap: int = 200):
    """↔ IEBoxer.DrawBoxes (IEB:37-81): Y-down coordinates relative to the image centre."""
    sx = f32(image_w) / f32(640)
    sy = f32(image_h) / f32(640)
    hw = f32(image_w) / f32(2)
    hh = f32(image_h) / f32(2)
    n = min(boxes.shape[0], cap)
    b = boxes[:n].astype(f32)
    out = np.empty((n, 4), f32)
    out[:, 0] = b[:, 0] * sx - hw
    out[:, 1] = b[:, 1] * sy - hh
    out[:, 2] = b[:, 2] * sx
    out[:, 3] = b[:, 3] * sy
    return out, np.asarray(label_ids[:n], np.int32)


# --------------------------------------------------------------------------------------
# C# mask post-processing: IEMasker.DrawMask / DrawSingleMask / PixelInBoundingBox
# --------------------------------------------------------------------------------------
MASK_W = 160
MASK_H = 160


def pixel_in_bounding_box(box, x: np.ndarray, y: np.ndarray, image_w: int, image_h: int) -> np.ndarray:
    """↔ IEMasker.PixelInBoundingBox (IEM:232-247), vectorised over integer pixel coords x,y.

    box = (CenterX, CenterY, Width, Height) as produced by parse_boxes / draw_boxes."""
    cxb, cyb, wb, hb = (f32(v) for v in box)
    xs = f32(MASK_W) / f32(image_w)
    ys = f32(MASK_H) / f32(image_h)
    cx = (cxb * xs) + f32(MASK_W // 2)
    cy = f32(MASK_H // 2) - (cyb * ys)
    hw = wb * xs / f32(2)
    hh = hb * ys / f32(2)
    xf = x.astype(f32)
    yf = y.astype(f32)
    return (xf >= (cx - hw)) & (xf <= (cx + hw)) & (yf >= (cy - hh)) & (yf <= (cy + hh))


def draw_mask_bits(mask: np.ndarray, box, image_w: int, image_h: int, thr: float = 0.5) -> np.ndarray:
    """One detection of IEMasker.DrawMask (IEM:98-113) / DrawSingleMask (IEM:167-185).

    mask f32 [160,160] (row y of output_3) -> uint8 [160,160] in TEXTURE order: element
    [posY, posX] with posY = 159 - y (IEM:103,174), 1 where the C# writes the colour, 0 where
    it writes Color.clear."""
    ys, xs = np.meshgrid(np.arange(MASK_H), np.arange(MASK_W), indexing="ij")
    pos_y = MASK_H - ys - 1
    on = (mask.astype(f32) > f32(thr)) & pixel_in_bounding_box(box, xs, pos_y, image_w, image_h)
    out = np.zeros((MASK_H, MASK_W), np.uint8)
    out[pos_y, xs] = on.astype(np.uint8)
    return out


# --------------------------------------------------------------------------------------
# Extension (not in the reference): Ultralytics-style crop + 640 upsample
# --------------------------------------------------------------------------------------
def crop_mask_native(prob: np.ndarray, box_cxcywh: np.ndarray, thr: float = 0.5, in_size: int = 640) -> np.ndarray:
    """Geometrically consistent crop at proto resolution: pixel (y,x) kept when its index lies in
    [x1, x2) x [y1, y2) of the box scaled by 160/640; uint8 [160,160], image row order."""
    cx, cy, w, h = (f32(v) for v in box_cxcywh)
    s = f32(MASK_W) / f32(in_size)
    x1 = (cx - w * f32(0.5)) * s
    x2 = (cx + w * f32(0.5)) * s
    y1 = (cy - h * f32(0.5)) * s
    y2 = (cy + h * f32(0.5)) * s
    ys, xs = np.meshgrid(np.arange(MASK_H, dtype=f32), np.arange(MASK_W, dtype=f32), indexing="ij")
    inside = (xs >= x1) & (xs < x2) & (ys >= y1) & (ys < y2)
    return ((prob.astype(f32) > f32(thr)) & inside).astype(np.uint8)


def upsample_mask_640(logit: np.ndarray, box_cxcywh: np.ndarray, out_size: int = 640) -> np.ndarray:
    """Bilinear (align_corners=False, edge clamp) upsample of mask LOGITS 160 -> 640, crop to the
    box in input pixels, threshold at logit > 0 (== sigmoid > 0.5); uint8 [640,640]."""
    h, w = logit.shape
    scale = f32(h) / f32(out_size)
    o = (np.arange(out_size, dtype=f32) + f32(0.5)) * scale - f32(0.5)
    o = np.maximum(o, f32(0))
    i0 = np.floor(o).astype(np.int64)
    i0 = np.minimum(i0, h - 1)
    i1 = np.minimum(i0 + 1, h - 1)
    t = (o - i0.astype(f32)).astype(f32)
    L = logit.astype(f32)
    top = (L[i0][:, i0] * (f32(1) - t)[None, :]).astype(f32) + (L[i0][:, i1] * t[None, :]).astype(f32)
    bot = (L[i1][:, i0] * (f32(1) - t)[None, :]).astype(f32) + (L[i1][:, i1] * t[None, :]).astype(f32)
    val = (top.astype(f32) * (f32(1) - t)[:, None]).astype(f32) + (bot.astype(f32) * t[:, None]).astype(f32)
    cx, cy, bw, bh = (f32(v) for v in box_cxcywh)
    x1, x2 = cx - bw * f32(0.5), cx + bw * f32(0.5)
    y1, y2 = cy - bh * f32(0.5), cy + bh * f32(0.5)
    ys, xs = np.meshgrid(np.arange(out_size, dtype=f32), np.arange(out_size, dtype=f32), indexing="ij")
    inside = (xs >= x1) & (xs < x2) & (ys >= y1) & (ys < y2)
    return ((val > f32(0)) & inside).astype(np.uint8)


# --------------------------------------------------------------------------------------
# RGB-D point extraction and target association
# --------------------------------------------------------------------------------------
def extract_points(mask_prob: np.ndarray, raw_box, depth_half: np.ndarray, screen_w: float, screen_h: float, cam_pos, cam_rot,
                   focal, principal, sensor_res, step: int = 5, thr: float = 0.5, max_points: int = 8000) -> np.ndarray:
    """↔ IEExecutor.ExtractDepthData (IEE:561-651) + DepthExtractionJob.Execute (IEE:86-156) + the in-order compaction of
    CollectJobResults (IEE:653-667).  mask_prob f32 [160,160] (output_3 row), raw_box = output_0 row (cx,cy,w,h @640),
    depth_half uint16 [H,W] half floats.  Returns f32 [n,4]: world x, y, z, depth (m).

    fp32, one IEEE operation per step, in the C# expression order; math.normalize = rsqrt(dot(v,v)) * v and
    math.mul(quaternion, float3) = v + q.w * t + cross(q.xyz, t), t = 2 * cross(q.xyz, v) (Unity.Mathematics)."""
    sx, sy = f32(screen_w) / f32(640), f32(screen_h) / f32(640)
    cx, cy, w, h = (f32(v) for v in raw_box)
    bx, by, bw, bh = (cx - f32(320)) * sx, (f32(320) - cy) * sy, w * sx, h * sy            # ParseBoxes, IEE:548-551
    rcx, rcy, rw, rh = bx / sx + f32(320), f32(320) - by / sy, bw / sx, bh / sy           # IEE:586-589
    dh, dw = depth_half.shape
    depth = depth_half.reshape(-1).view(np.float16).astype(f32)
    qx, qy, qz, qw = (f32(v) for v in cam_rot)
    pos = [f32(v) for v in cam_pos]
    out = []
    total_x = 160 // step
    for index in range(total_x * total_x):
        ly, lx = divmod(index, total_x)
        y, x = ly * step, lx * step
        if y >= 160 or x >= 160 or not (mask_prob[y, x] > f32(thr)):
            continue
        nx, ny = f32(x) / f32(160), f32(y) / f32(160)
        ipx = f32(f32(rcx - f32(rw * f32(0.5))) + f32(nx * rw))
        ipy = f32(f32(rcy - f32(rh * f32(0.5))) + f32(ny * rh))
        u = min(max(f32(ipx / f32(640)), f32(0)), f32(1))
        v = min(max(f32(ipy / f32(640)), f32(0)), f32(1))
        omv = f32(f32(1) - v)
        dx = int(f32(u * f32(dw - 1)))
        dy = int(f32(omv * f32(dh - 1)))
        di = dy * dw + dx
        if di < 0 or di >= dw * dh:
            continue
        d = depth[di]
        if not (d > f32(0.1) and d < f32(3.0)):
            continue
        cpx, cpy = f32(u * f32(sensor_res[0])), f32(omv * f32(sensor_res[1]))
        vx = f32(f32(cpx - f32(principal[0])) / f32(focal[0]))
        vy = f32(f32(cpy - f32(principal[1])) / f32(focal[1]))
        vz = f32(1)
        dot = f32(f32(f32(vx * vx) + f32(vy * vy)) + f32(vz * vz))
        inv = f32(f32(1) / np.sqrt(dot, dtype=f32))
        vx, vy, vz = f32(inv * vx), f32(inv * vy), f32(inv * vz)
        tx = f32(f32(2) * f32(f32(qy * vz) - f32(qz * vy)))
        ty = f32(f32(2) * f32(f32(qz * vx) - f32(qx * vz)))
        tz = f32(f32(2) * f32(f32(qx * vy) - f32(qy * vx)))
        wx = f32(f32(vx + f32(qw * tx)) + f32(f32(qy * tz) - f32(qz * ty)))
        wy = f32(f32(vy + f32(qw * ty)) + f32(f32(qz * tx) - f32(qx * tz)))
        wz = f32(f32(vz + f32(qw * tz)) + f32(f32(qx * ty) - f32(qy * tx)))
        out.append((f32(pos[0] + f32(wx * d)), f32(pos[1] + f32(wy * d)), f32(pos[2] + f32(wz * d)), d))
        if len(out) >= max_points:
            break
    return np.asarray(out, f32).reshape(-1, 4)


def associate(boxes: np.ndarray, label_ids: np.ndarray, locked_cx: float, locked_cy: float, locked_label: int,
              screen_w: float, screen_h: float, max_dist: float = 300.0):
    """↔ the locked-target search of IEExecutor.ProcessInferenceResult (IEE:488-507) over ParseBoxes' list (cap 50):
    nearest box of the same class, strict `<` (first of equal distances); (-1, dist) when not closer than max_dist."""
    pb, lab = parse_boxes(boxes, label_ids, screen_w, screen_h)
    best, mind = -1, f32(3.402823466e+38)
    for i in range(len(pb)):
        if int(lab[i]) != int(locked_label):
            continue
        dx, dy = f32(pb[i, 0] - f32(locked_cx)), f32(pb[i, 1] - f32(locked_cy))
        dist = np.sqrt(f32(f32(dx * dx) + f32(dy * dy)), dtype=f32)
        if dist < mind:
            mind, best = dist, i
    return (best if best != -1 and mind < f32(max_dist) else -1), float(mind)
