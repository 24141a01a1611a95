"""Host-side mirror of the reference's `omnistereo` package for the SOS front-end hot path.

Same module, class, method and keyword names as ubuntuslave/vo_single_camera_sos (NumPy in, NumPy out, the reference's
shapes and dtypes), with the arithmetic dispatched to the sm_100a kernels of libsosfront.so.  Only what the hot path of
demo_vo_sos.py / demo_vo_rgbd.py reaches is mirrored; calibration, GUIs, dense stereo and plotting are out
of scope.  The repository-root package `omnistereo` re-exports these modules under the reference's import paths, so
pickled `omnistereo.gum.GUMStereo` models resolve here.
"""
import torch

from .. import ops


def device_context() -> ops.Context:
    """Per-thread context on the current CUDA device (the reference's VO loop runs on its own thread)."""
    if not torch.cuda.is_available():
        raise RuntimeError("the omnistereo mirror runs its arithmetic on a B200; no CUDA device is visible "
                           "(there is no CPU fallback)")
    return ops.default_context()


def to_device(a, dtype=None):
    import numpy as np
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()
