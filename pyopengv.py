"""Import-path alias: `import pyopengv` (pose_est_tools.py:42 of the reference) resolves to the RANSAC-kernel backed
replacement in vo_single_camera_sos_b200.pyopengv."""
from vo_single_camera_sos_b200.pyopengv import *  # noqa: F401,F403
from vo_single_camera_sos_b200.pyopengv import (absolute_pose_noncentral_optimize_nonlinear, absolute_pose_noncentral_ransac,  # noqa: F401
                                                absolute_pose_optimize_nonlinear, absolute_pose_ransac, hypothesis_list,
                                                relative_pose_ransac, triangulation_triangulate, triangulation_triangulate2)
