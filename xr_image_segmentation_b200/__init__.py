"""xr_image_segmentation_b200 -- B200-native (sm_100a) drop-in for the per-frame YOLO11-seg hot path of
netlab-dgist/xr-image-segmentation.

(The repository brief names the package `xr-image-segmentation_b200`; a Python package name cannot contain a hyphen,
hence the underscore.)

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (libxrseg.so, declared in include/xrseg.h)
  _lib.py          ctypes binding of the C ABI (the only way Python reaches the kernels)
  weights.py       XRSW weight-pack reader / writer, random-init generator
  inference.py     host-side mirror of the Unity Inference Engine surface the reference calls
                   (ModelLoader / Worker / Tensor / TextureConverter)
  executor.py      host-side mirror of IEExecutor / IEBoxer / IEMasker (state machine, box + mask post-processing)

There is no CPU fallback: every compute call goes through libxrseg.so and fails loudly without it.
"""
from ._lib import XrsegError, load_library, library_path  # noqa: F401

__all__ = ["XrsegError", "load_library", "library_path"]
