"""Oracle for the ORB description of given keypoints.  TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

The primary checker is cv2.ORB.compute itself (the call the reference makes, camera_models.py:1683, 1766); this module
restates what that call does for keypoints of octave 0, as identified by scripts/derive_orb_pattern.py, so that the
stages can be tested separately."""
from __future__ import annotations

import os

import cv2
import numpy as np

_TABLE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "orb_pattern.npy")


def pattern() -> np.ndarray:
    return np.load(_TABLE)


def gaussian_kernel_7_2() -> np.ndarray:
    x = np.arange(-3, 4, dtype=np.float64)
    k = np.exp(-(x * x) / (2.0 * 2.0 * 2.0))
    return k / k.sum()


def blur(gray: np.ndarray) -> np.ndarray:
    """The blur inside cv2.ORB.compute: exact separable 7-tap Gaussian of sigma 2 (BORDER_REFLECT_101) rounded once."""
    k = gaussian_kernel_7_2().reshape(-1, 1)
    return np.rint(cv2.sepFilter2D(gray.astype(np.float64), -1, k, k, borderType=cv2.BORDER_REFLECT_101)).astype(np.uint8)


def describe(gray: np.ndarray, pts: np.ndarray, angles_deg: np.ndarray, table: np.ndarray | None = None):
    """-> (descriptors uint8 [n, 32], keep bool [n]); float32 rotation arithmetic as in OpenCV's computeOrbDescriptors."""
    t = (pattern() if table is None else table).astype(np.float32)
    h, w = gray.shape
    B = blur(gray)
    out = np.zeros((len(pts), 32), np.uint8)
    keep = np.zeros(len(pts), bool)
    for i, ((x, y), ang) in enumerate(zip(pts, angles_deg)):
        cx, cy = int(np.rint(np.float32(x))), int(np.rint(np.float32(y)))
        if not (31 <= cx < w - 31 and 31 <= cy < h - 31):
            continue
        keep[i] = True
        ang32 = np.float32(ang) * np.float32(np.pi / 180.0)
        a, b = np.float32(np.cos(ang32)), np.float32(np.sin(ang32))
        x0 = np.rint(t[:, 0] * a - t[:, 1] * b).astype(int); y0 = np.rint(t[:, 0] * b + t[:, 1] * a).astype(int)
        x1 = np.rint(t[:, 2] * a - t[:, 3] * b).astype(int); y1 = np.rint(t[:, 2] * b + t[:, 3] * a).astype(int)
        out[i] = np.packbits(B[cy + y0, cx + x0] < B[cy + y1, cx + x1], bitorder="little")
    return out, keep
