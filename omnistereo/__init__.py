"""Import-path alias: `omnistereo.*` resolves to the B200-backed mirror in vo_single_camera_sos_b200.omnistereo, so that
code (and pickles of omnistereo.gum.GUMStereo, demo_vo_sos.py:109) written against the reference keeps importing."""
import importlib
import sys

_MIRROR = "vo_single_camera_sos_b200.omnistereo"
for _name in ("common_tools", "common_cv", "common_plot", "transformations", "panorama", "camera_models", "gum", "pose_est_tools"):
    _mod = importlib.import_module(f"{_MIRROR}.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod
