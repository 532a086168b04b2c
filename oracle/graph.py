"""ORACLE (test infrastructure, not product): fp32 CPU interpreter for the reference's
serialized Inference Engine graph.

Restates what `Worker.ScheduleIterable(input)` + the `MoveNext()` loop compute
(/root/reference/Assets/Scripts/InferenceEngine/IEExecutor.cs:371,397): the 499 layers
of `yolo11n-seg-sentis.sentis`, whose tail (decode → ReduceMax/ArgMax → corner MatMul →
NMS → gathers → mask MatMul → Sigmoid) was appended by
Assets/Scripts/InferenceEngine/Editor/IEModelEditorConverter.cs:31-106.

The arithmetic lives in the un-vendored package `com.unity.ai.inference` 2.2.1
(Packages/manifest.json:4), so op semantics are restated from their published
ONNX-equivalent definitions; per-op attribute order is the one observed in the asset
(SURVEY.md Appendix C).  PARITY UNPINNED by the reference (it has no tests or golden
tensors) -- this oracle is the pin; see DESIGN.md "Oracle choices".
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .postprocess import nms_onnx
from .sentis import SentisModel, TensorValue


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x)


class GraphInterpreter:
    """Executes every chain in file order; keeps all intermediate values (`self.env`)."""

    def __init__(self, model: SentisModel):
        self.m = model
        self.const = {}
        for i, v in enumerate(model.values):
            if isinstance(v, TensorValue) and v.is_const:
                self.const[i] = torch.from_numpy(v.data)

    # -- attribute helpers --------------------------------------------------------
    def _a(self, c, k):
        return self.m.values[c.args[k]]

    @torch.no_grad()
    def run(self, images: torch.Tensor, keep: set[int] | None = None, stop_after: int | None = None):
        """images: f32 [1,3,640,640] in 0..1 (↔ TextureConverter.ToTensor output, IEExecutor.cs:370).

        Returns dict value-id -> tensor for the graph outputs and any ids in `keep`."""
        env: dict[int, torch.Tensor] = dict(self.const)
        env[self.m.inputs[0]] = images.float()
        keep = set(keep or ()) | set(self.m.outputs)
        for c in self.m.chains:
            ins = [env[i] if i >= 0 else None for i in c.inputs]
            outs = getattr(self, "op_" + c.op)(c, *ins)
            if not isinstance(outs, (list, tuple)):
                outs = [outs]
            for o, v in zip(c.outputs, outs):
                env[o] = v
            if stop_after is not None and c.index >= stop_after:
                break
        self.env = env
        return {k: env[k] for k in keep if k in env}

    # -- ops ----------------------------------------------------------------------
    def op_DequantizeUint8(self, c, q):
        scale, zp = self._a(c, 0), self._a(c, 1)
        return (q.to(torch.float32) - float(zp)) * float(np.float32(scale))

    def op_Conv(self, c, x, w, b):
        _auto, dil, group, pads, strides, _kernel, fused = (self._a(c, k) for k in range(7))
        assert fused == 0 and pads[0] == pads[2] and pads[1] == pads[3]
        return F.conv2d(x, w, b, stride=strides, padding=(pads[0], pads[1]), dilation=dil, groups=group)

    def op_ConvTranspose(self, c, x, w, b):
        _auto, _outpad, pads, strides, _kernel, fused = (self._a(c, k) for k in range(6))
        assert fused == 0 and all(p == 0 for p in pads)
        return F.conv_transpose2d(x, w, b, stride=strides)

    def op_Swish(self, c, x):
        return x * torch.sigmoid(x)

    def op_Sigmoid(self, c, x):
        return torch.sigmoid(x)

    def op_Split(self, c, x, sizes):
        axis = self._a(c, 0)
        return list(torch.split(x, [int(s) for s in sizes], dim=axis))

    def op_Add(self, c, a, b):
        return a + b

    def op_Sub(self, c, a, b):
        return a - b

    def op_Mul(self, c, a, b):
        return a * b

    def op_Concat(self, c, *xs):
        return torch.cat(xs, dim=self._a(c, 0))

    def op_MaxPool(self, c, x):
        kernel, strides, pads, _auto = (self._a(c, k) for k in range(4))
        return F.max_pool2d(x, kernel, strides, (pads[0], pads[1]))

    def op_Reshape(self, c, x, shape):
        return x.reshape([int(s) for s in shape])

    def op_Transpose(self, c, x):
        return x.permute(self._a(c, 0))

    def op_MoveDim(self, c, x):
        return torch.movedim(x, self._a(c, 0), self._a(c, 1))

    def op_MatMul(self, c, a, b):
        return torch.matmul(a, b)

    def op_ScalarMad(self, c, x):
        return x * float(np.float32(self._a(c, 1))) + float(np.float32(self._a(c, 2)))

    def op_Softmax(self, c, x):
        return torch.softmax(x, dim=self._a(c, 0))

    def op_Resize(self, c, x, scales):
        s = [float(v) for v in scales]
        assert s[0] == 1 and s[1] == 1 and s[2] == 2 and s[3] == 2
        return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)  # nearest: out[i] = in[i // 2]

    def op_Slice(self, c, x, starts, ends, axes, steps=None):
        idx = [slice(None)] * x.dim()
        for k, ax in enumerate(axes.tolist()):
            st = int(steps[k]) if steps is not None else 1
            idx[ax] = slice(int(starts[k]), min(int(ends[k]), x.shape[ax]), st)
        return x[tuple(idx)]

    def op_Squeeze(self, c, x, axes):
        for ax in sorted((int(a) for a in axes), reverse=True):
            x = x.squeeze(ax)
        return x

    def op_Unsqueeze(self, c, x, axes):
        for ax in sorted(int(a) for a in axes):
            x = x.unsqueeze(ax)
        return x

    def op_ReduceMax(self, c, x, axes):
        keepdims = self._a(c, 0)
        return torch.amax(x, dim=[int(a) for a in axes], keepdim=bool(keepdims))

    def op_ArgMax(self, c, x):
        axis, keepdims, select_last = (self._a(c, k) for k in range(3))
        assert not select_last
        # first maximum wins (torch.argmax on CPU returns the first occurrence; made explicit here)
        mx = torch.amax(x, dim=axis, keepdim=True)
        first = torch.argmax((x == mx).to(torch.int8), dim=axis, keepdim=bool(keepdims))
        return first.to(torch.int32)

    def op_NonMaxSuppression(self, c, boxes, scores, max_out, iou_thr, score_thr):
        assert self._a(c, 0) == 0  # centerPointBox = 0 -> corner boxes
        assert int(max_out) == -1  # unlimited
        b = boxes[0].numpy()
        s = scores[0, 0].numpy()
        keep = nms_onnx(b, s, float(np.float32(iou_thr)), float(np.float32(score_thr)))
        out = np.zeros((len(keep), 3), np.int32)
        out[:, 2] = keep
        return torch.from_numpy(out)

    def op_Select(self, c, x, dim, index):
        return x.select(int(dim), int(index))

    def op_Expand(self, c, x, shape):
        shp = [int(s) for s in shape]
        tgt = list(torch.broadcast_shapes(tuple(x.shape), tuple(shp)))
        return x.expand(tgt)

    def op_GatherElements(self, c, x, idx):
        return torch.gather(x, self._a(c, 0), idx.to(torch.int64))


# value ids of interest in the shipped asset
V_HEAD = 1844       # [1,116,8400]: cx,cy,w,h | 80 class probs | 32 coefs
V_HEAD_RAW = 1631   # [1,144,8400]: 64 DFL logits | 80 class logits
V_COEF = 1842       # [1,32,8400]
V_PROTO = 1984      # [32,25600]
V_CORNERS = 1860    # [1,8400,4]
V_SCORES = 1876     # [1,1,8400]
V_KEEP = 1884       # [N]
V_P3, V_P4, V_P5 = 936, 1169, 1490
