"""Host-side mirror of the reference's runner scripts on top of the CUDA library.

  IEExecutor  ↔ Assets/Scripts/InferenceEngine/IEExecutor.cs  (state machine 389-456, ParseBoxes 529-559)
  IEBoxer     ↔ Assets/Scripts/InferenceEngine/IEBoxer.cs     (DrawBoxes 37-81, GetClassName 183-188)
  IEMasker    ↔ Assets/Scripts/InferenceEngine/IEMasker.cs    (DrawMask 82-119, DrawSingleMask 124-196,
                                                               PixelInBoundingBox 232-247)

Same names, argument meaning and error behaviour; the uGUI drawing / smoothing / tracking / point-cloud parts of
those scripts are out of scope -- methods return the data the C# would draw.  Every number is computed
by a CUDA kernel of libxrseg.so (xrseg_decode / xrseg_masks); nothing is recomputed in Python.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass

import numpy as np

from . import _lib
from .inference import BackendType, InputTensor, ModelLoader, Tensor, TextureConverter, Worker

YOLO11_MASK_HEIGHT = 160
YOLO11_MASK_WIDTH = 160


@dataclass
class BoundingBox:
    """↔ struct BoundingBox (IEB:6-15)."""
    CenterX: float
    CenterY: float
    Width: float
    Height: float
    Label: str
    ClassName: str
    WorldPos: tuple | None = None


class IEBoxer:
    def __init__(self, labels_text: str):
        # ↔ IEBoxer.Start (IEB:31-35): Split on \n and \r, RemoveEmptyEntries
        self._labels = [s for s in labels_text.replace("\r", "\n").split("\n") if s != ""]

    def GetClassName(self, labelId: int) -> str:
        """↔ IEB:183-188."""
        if self._labels is None or labelId < 0 or labelId >= len(self._labels):
            return "unknown"
        return self._labels[labelId].replace(" ", "_")

    def DrawBoxes(self, executor: "IEExecutor", imageWidth: float, imageHeight: float) -> list[BoundingBox]:
        """↔ IEBoxer.DrawBoxes(output, labelIds, imageWidth, imageHeight) (IEB:37-81), first 200 boxes."""
        boxes, labels, _ = executor._runner.decode(imageWidth, imageHeight, _lib.BOX_DRAWBOXES)
        out = []
        for b, l in zip(boxes, labels):
            name = self._labels[l].replace(" ", "_")          # IEB:59: no bounds check, like the C#
            out.append(BoundingBox(float(b[0]), float(b[1]), float(b[2]), float(b[3]), name, name))
        return out


class IEMasker:
    def __init__(self, confidenceThreshold: float = 0.5):
        self._confidenceThreshold = confidenceThreshold   # ↔ IEMasker.Initialize(display, 0.5) (IEE:262)

    def _thr(self) -> float:
        """The threshold of `mask[i,y,x] > _confidenceThreshold` (IEM:104,176) in the C ABI's encoding (0 = the runner's
        default, negative = exactly 0)."""
        return -1.0 if self._confidenceThreshold == 0 else float(self._confidenceThreshold)

    def DrawMask(self, executor: "IEExecutor", imageWidth: int, imageHeight: int) -> np.ndarray:
        """↔ IEMasker.DrawMask(boundBoxes, mask, imageWidth, imageHeight) (IEM:82-119) with the DrawBoxes boxes:
        uint8 [n,160,160] in texture order, 1 where the C# writes the mask colour."""
        return executor._runner.masks(_lib.MASK_REFERENCE_160, _lib.BOX_DRAWBOXES, float(imageWidth), float(imageHeight),
                                      int(imageWidth), int(imageHeight), threshold=self._thr())

    def DrawSingleMask(self, executor: "IEExecutor", targetIndex: int, screenW: float, screenH: float, imageWidth: int,
                       imageHeight: int) -> np.ndarray | None:
        """↔ IEMasker.DrawSingleMask(targetIndex, box, mask, imageWidth, imageHeight) (IEM:124-196), fed -- like
        IEE:516 -- the ParseBoxes box of the target: uint8 [160,160] texture order, or None when targetIndex < 0."""
        if targetIndex < 0:
            return None
        m = executor._runner.masks(_lib.MASK_REFERENCE_160, _lib.BOX_PARSEBOXES, float(screenW), float(screenH),
                                   int(imageWidth), int(imageHeight), first=targetIndex, count=1, threshold=self._thr())
        return m[0]


class InferenceDownloadState(enum.IntEnum):
    """↔ IEE:17-25."""
    Running = 0
    RequestingOutputs = 1
    Success = 2
    Error = 3
    Cleanup = 4
    Completed = 5


class IEExecutor:
    """↔ the inference-runner part of IEExecutor (LoadModel 380-387, RunInference 363-376, UpdateInference 389-417,
    UpdateParallelReadbacks 419-456, ProcessInferenceResult 458-481 non-tracking branch, ParseBoxes 529-559)."""

    def __init__(self, sentisModel, labels_text: str, screen=(1920.0, 1080.0), backend=BackendType.GPUCompute,
                 layersPerFrame: int = 25, confidenceThreshold: float = 0.5, device: int = 0, **runner_kw):
        self._backend = backend
        self._layersPerFrame = layersPerFrame                # XRScene.unity:1223
        self._confidenceThreshold = confidenceThreshold      # IEE:32
        self.Screen = screen
        self._ieBoxer = IEBoxer(labels_text)
        self._ieMasker = IEMasker(confidenceThreshold)
        self._started = False
        self._downloadState = InferenceDownloadState.Completed
        self._readbacksInitiated = False
        self._outputBuffers = [None] * 4
        self._readbackComplete = [False] * 4
        self._outputs = [None] * 4
        self._input = None
        self._inputSize = (640, 640)
        self.CurrentFrameBoxes: list[BoundingBox] = []
        self.LastMasks = None
        self.IsModelLoaded = False
        self._LoadModel(sentisModel, device, runner_kw)

    # ↔ LoadModel (IEE:380-387): load, create the worker, warm-up run on a blank frame
    def _LoadModel(self, sentisModel, device, runner_kw):
        model = ModelLoader.Load(sentisModel)
        # _confidenceThreshold (IEE:32) is also the runner's mask threshold (depth extraction IEE:102, bit-packed masks)
        runner_kw.setdefault("mask_thr", -1.0 if self._confidenceThreshold == 0 else float(self._confidenceThreshold))
        self._inferenceEngineWorker = Worker(model, self._backend, device=device, **runner_kw)
        self._runner = self._inferenceEngineWorker._runner
        blank = TextureConverter.ToTensor(np.zeros((self._inputSize[1], self._inputSize[0], 3), np.uint8), 640, 640, 3)
        self._inferenceEngineWorker.Schedule(blank)
        self._runner.wait()
        self.IsModelLoaded = True

    def RunInference(self, inputTexture):
        """↔ IEE:363-376."""
        if not self._started:
            if self._input is not None:
                self._input.Dispose()
            if inputTexture is None:
                return
            self._inputSize = (inputTexture.shape[1], inputTexture.shape[0])
            self._input = TextureConverter.ToTensor(inputTexture, 640, 640, 3)
            self._schedule = self._inferenceEngineWorker.ScheduleIterable(self._input)
            self._downloadState = InferenceDownloadState.Running
            self._started = True
            self._readbacksInitiated = False

    def IsRunning(self) -> bool:
        return self._started

    def Update(self):
        self.UpdateInference()

    def UpdateInference(self):
        """↔ IEE:389-417."""
        if not self._started:
            return
        st = self._downloadState
        if st == InferenceDownloadState.Running:
            it = 0
            while self._schedule.MoveNext():
                it += 1
                if it % self._layersPerFrame == 0:
                    return
            self._downloadState = InferenceDownloadState.RequestingOutputs
        elif st == InferenceDownloadState.RequestingOutputs:
            self.UpdateParallelReadbacks()
        elif st == InferenceDownloadState.Success:
            self.ProcessInferenceResult()
            self._downloadState = InferenceDownloadState.Cleanup
        elif st in (InferenceDownloadState.Error, InferenceDownloadState.Cleanup):
            self.CleanupResources()
            self._downloadState = InferenceDownloadState.Completed
            self._started = False

    def UpdateParallelReadbacks(self):
        """↔ IEE:419-456."""
        if not self._readbacksInitiated:
            for i in range(4):
                self._readbackComplete[i] = False
                self._outputBuffers[i] = self._inferenceEngineWorker.PeekOutput(i)
                if self._outputBuffers[i].dataOnBackend is not None:
                    self._outputBuffers[i].ReadbackRequest()
                else:
                    self._downloadState = InferenceDownloadState.Error
                    return
            self._readbacksInitiated = True
            return
        allComplete = True
        for i in range(4):
            if not self._readbackComplete[i]:
                if self._outputBuffers[i].IsReadbackRequestDone():
                    self._readbackComplete[i] = True
                else:
                    allComplete = False
        if allComplete:
            self._outputs = [self._outputBuffers[i].ReadbackAndClone() for i in range(4)]
            for i in range(4):
                self._outputBuffers[i].Dispose()
                self._outputBuffers[i] = None
            ok = self._outputs[0] is not None and self._outputs[0].shape[0] > 0
            self._downloadState = InferenceDownloadState.Success if ok else InferenceDownloadState.Error

    def ProcessInferenceResult(self):
        """↔ IEE:458-481 (non-tracking branch): ParseBoxes, then DrawBoxes data; masks of all boxes on request."""
        screenW, screenH = self.Screen
        self.CurrentFrameBoxes = self.ParseBoxes(screenW, screenH)
        self.LastDrawBoxes = self._ieBoxer.DrawBoxes(self, screenW, screenH)

    def ParseBoxes(self, screenW: float, screenH: float) -> list[BoundingBox]:
        """↔ IEE:529-559 (first 50 boxes, centred Y-up screen coordinates)."""
        boxes, labels, _ = self._runner.decode(screenW, screenH, _lib.BOX_PARSEBOXES)
        out = []
        for b, l in zip(boxes, labels):
            name = self._ieBoxer.GetClassName(int(l))
            out.append(BoundingBox(float(b[0]), float(b[1]), float(b[2]), float(b[3]), name, name))
        return out

    def CleanupResources(self):
        """↔ IEE:693-701."""
        for t in self._outputs:
            if t is not None:
                t.Dispose()
        self._outputs = [None] * 4
        self._readbacksInitiated = False

    def OnDestroy(self):
        self._inferenceEngineWorker.Dispose()
