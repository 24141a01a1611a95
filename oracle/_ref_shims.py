"""Harness-side shims that let the UNMODIFIED reference package import under NumPy 2 / Python 3.12 / headless OpenCV
(SURVEY §8c).  Used only by oracle/gen_golden.py, which runs in the build container where /root/reference exists.
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import collections
import collections.abc
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def install():
    if not hasattr(np, "float"):
        np.float = float  # camera_models.py:337,347,3083-3084; pose_est_tools.py:237,241
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable  # common_tools.py:260
    if "pyopengv" not in sys.modules:  # pose_est_tools.py:42 imports it at module load
        m = types.ModuleType("pyopengv")

        def _absent(*a, **k):
            raise RuntimeError("pyopengv is not available in this container (OpenGV parity is unpinned)")

        for name in ("triangulation_triangulate", "triangulation_triangulate2", "absolute_pose_noncentral_ransac",
                     "absolute_pose_noncentral_optimize_nonlinear", "absolute_pose_ransac",
                     "absolute_pose_optimize_nonlinear", "relative_pose_ransac"):
            setattr(m, name, _absent)
        sys.modules["pyopengv"] = m
    if "omnistereo.common_plot" not in sys.modules:  # matplotlib / vispy are not installed
        m = types.ModuleType("omnistereo.common_plot")
        m.draw_matches_between_frames = lambda *a, **k: None
        m.DrawerVO = object
        m.replay_VO_visualization = lambda *a, **k: None
        sys.modules["omnistereo.common_plot"] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True
