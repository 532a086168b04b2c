"""ORACLE (test infrastructure, not product): CPU restatement of the frame -> tensor step.

↔ `TextureConverter.ToTensor(inputTexture, 640, 640, 3)`
(/root/reference/Assets/Scripts/InferenceEngine/IEExecutor.cs:370): the camera texture is
STRETCHED (no letterbox) to 640x640, RGB, floats in 0..1, NCHW.

The resampling filter / colour-space handling of that call live in the un-vendored
`com.unity.ai.inference` 2.2.1 package and are UNPINNED.  oracle choice: a GPU
blit is a bilinear texture fetch, so: sample position (x + 0.5) * src_w / 640 - 0.5 with edge
clamp, 2x2 bilinear weights in fp32, uint8 / 255 applied AFTER interpolation, top row first,
no sRGB conversion.  `letterbox()` is the extension for BASELINE.json config 4.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _axis(n_out: int, n_in: int, scale=None, offset=0.0):
    s = f32(n_in) / f32(n_out) if scale is None else f32(scale)
    pos = (np.arange(n_out, dtype=f32) + f32(0.5) - f32(offset)) * s - f32(0.5)
    i0f = np.floor(pos)
    t = (pos - i0f).astype(f32)
    i0 = i0f.astype(np.int64)
    i1 = i0 + 1
    return np.clip(i0, 0, n_in - 1), np.clip(i1, 0, n_in - 1), t


def to_tensor(img_u8: np.ndarray, out_hw=(640, 640)) -> np.ndarray:
    """img_u8 [H,W,3|4] uint8 -> f32 [1,3,640,640] in 0..1 (alpha dropped)."""
    img = img_u8[..., :3].astype(f32)
    H, W = img.shape[:2]
    oh, ow = out_hw
    y0, y1, ty = _axis(oh, H)
    x0, x1, tx = _axis(ow, W)
    tx_ = tx[None, :, None]
    ty_ = ty[:, None, None]
    top = (img[y0][:, x0] * (f32(1) - tx_)).astype(f32) + (img[y0][:, x1] * tx_).astype(f32)
    bot = (img[y1][:, x0] * (f32(1) - tx_)).astype(f32) + (img[y1][:, x1] * tx_).astype(f32)
    val = (top.astype(f32) * (f32(1) - ty_)).astype(f32) + (bot.astype(f32) * ty_).astype(f32)
    val = (val.astype(f32) / f32(255)).astype(f32)
    return np.ascontiguousarray(val.transpose(2, 0, 1)[None])


def letterbox(img_u8: np.ndarray, out_hw=(640, 640), pad_value: int = 114) -> np.ndarray:
    """Extension (not in the reference): aspect-preserving resize + centred padding with 114."""
    img = img_u8[..., :3].astype(f32)
    H, W = img.shape[:2]
    oh, ow = out_hw
    r = min(oh / H, ow / W)
    nh, nw = int(round(H * r)), int(round(W * r))
    top, left = (oh - nh) // 2, (ow - nw) // 2
    y0, y1, ty = _axis(nh, H)
    x0, x1, tx = _axis(nw, W)
    tx_ = tx[None, :, None]
    ty_ = ty[:, None, None]
    a = (img[y0][:, x0] * (f32(1) - tx_)).astype(f32) + (img[y0][:, x1] * tx_).astype(f32)
    b = (img[y1][:, x0] * (f32(1) - tx_)).astype(f32) + (img[y1][:, x1] * tx_).astype(f32)
    val = (a.astype(f32) * (f32(1) - ty_)).astype(f32) + (b.astype(f32) * ty_).astype(f32)
    out = np.full((oh, ow, 3), f32(pad_value), f32)
    out[top:top + nh, left:left + nw] = val
    out = (out / f32(255)).astype(f32)
    return np.ascontiguousarray(out.transpose(2, 0, 1)[None])
