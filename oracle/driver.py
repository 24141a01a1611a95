"""CPU restatement of the reference's VO loop, one frame per iteration, as run_VO writes it
(pose_est_tools.py:1403-1628): every frame is tracked against the current keyframe, the keyframe decision tree, the
cumulative average of tracked correspondences, pose chaining through the keyframe list, and the TUM line format.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker of vo_single_camera_sos_b200/driver.py."""
from __future__ import annotations

import numpy as np

from . import pipeline


def rpe_translation_metric(T):
    """transformations.py:2078-2084"""
    return np.linalg.norm(T[:3, 3])


def rpe_rotation_metric(T):
    """transformations.py:2097-2106"""
    d = 0.5 * (np.trace(T[0:3, 0:3]) - 1.0)
    return np.arccos(min(1.0, max(-1.0, d)))


def quaternion_from_matrix(matrix):
    """transformations.quaternion_from_matrix(isprecise=False), transformations.py:1311-1332 -> [w, x, y, z]."""
    M = np.asarray(matrix, np.float64)[:4, :4]
    m00, m01, m02 = M[0, 0], M[0, 1], M[0, 2]
    m10, m11, m12 = M[1, 0], M[1, 1], M[1, 2]
    m20, m21, m22 = M[2, 0], M[2, 1], M[2, 2]
    K = np.array([[m00 - m11 - m22, 0.0, 0.0, 0.0],
                  [m01 + m10, m11 - m00 - m22, 0.0, 0.0],
                  [m02 + m20, m12 + m21, m22 - m00 - m11, 0.0],
                  [m21 - m12, m02 - m20, m10 - m01, m00 + m11 + m22]]) / 3.0
    w, V = np.linalg.eigh(K)
    q = V[[3, 0, 1, 2], np.argmax(w)]
    return -q if q[0] < 0.0 else q


def tum_line(idx, T):
    """pose_est_tools.py:1609-1612: print(idx, t[0], t[1], t[2], q[1], q[2], q[3], q[0], sep=' ')"""
    q = quaternion_from_matrix(T)
    t = T[:3, 3]
    return " ".join(str(v) for v in (idx, t[0], t[1], t[2], q[1], q[2], q[3], q[0]))


INDOOR = dict(pos_min=0.01, pos_max=0.20, ang_min=np.deg2rad(1.0), ang_max=np.deg2rad(10.0), tracked_ratio=0.10,
              keypoint_ratio=0.10)  # pose_est_tools.py:1303-1309


def run_vo(frames, stereo_args=None, track_args=None, thresholds=None, units_to_m=1.0, number_of_cams=2, refine="arun",
           frame_fn=None, track_fn=None):
    """frames: list of dicts (px_top, desc_top, boff_top, px_bot, desc_bot, boff_bot) of ONE frame each (trimmed to their
    valid rows); stereo_args = (pano_g, f_top, f_bot, cap); track_args = (hyp, mode, threshold, rig, max_du).
    frame_fn(f) / track_fn(reference, cur) replace the oracle pipeline (host-logic tests feed canned results).
    Returns dict(poses_wrt_S, poses_wrt_keyframe, keyframe_ids, tracked, decisions, status)."""
    th = dict(INDOOR if thresholds is None else thresholds)
    if frame_fn is None:
        pano_g, f_top, f_bot, cap = stereo_args
        frame_fn = lambda f: pipeline.stereo_frame(pano_g, f_top, f_bot, f["px_top"], f["desc_top"], f["boff_top"],
                                                   f["px_bot"], f["desc_bot"], f["boff_bot"], cap=cap)
    if track_fn is None:
        hyp, mode, threshold, rig, max_du = track_args
        track_fn = lambda ref, cur: pipeline.track_pair(ref, cur, hyp, mode, threshold, rig, max_du, refine=refine)
    T_key_list = []
    T_curr = np.eye(4)
    reference = None
    create_keyframe = True                       # :1406
    n_tracked_wrt_key = 0
    prev_avg = 0.0
    out = dict(poses_wrt_S=[], poses_wrt_keyframe=[], keyframe_ids=[], tracked=[], decisions=[], status="ok")
    for idx, f in enumerate(frames):
        cur = frame_fn(f)
        T_rel = np.eye(4)
        inl = 0
        if idx > 0:
            o = track_fn(reference, cur)
            if o is None or o["n_corr"] < 2 * 3 * (0.33 * number_of_cams) or o["best_hyp"] < 0:   # :779-781
                out["status"] = "tracking failed at frame %d" % idx
                break
            T_rel[:3] = o["refit"]
            T_rel[:3, 3] = T_rel[:3, 3] * units_to_m                                            # :833
            T_curr = T_key_list[-1] @ T_rel                                                       # :837
            n_tracked_wrt_key += 1                                                                # :1492
            inl = int(o["best_count"])
            dist = rpe_translation_metric(T_rel)
            ang = rpe_rotation_metric(T_rel)
            num_tracked = inl / float(number_of_cams)                                            # :1513
            M_K, M_F = len(reference["xyz"]), len(cur["xyz"])
            if (th["pos_min"] < dist < th["pos_max"]) or (th["ang_min"] < ang < th["ang_max"]):   # :1519
                if num_tracked > th["tracked_ratio"] * prev_avg and M_F > th["keypoint_ratio"] * M_K:
                    if th["pos_min"] < dist < th["pos_max"]:
                        if ang < th["ang_max"]:
                            create_keyframe = True
                    elif th["ang_min"] < ang:
                        if dist < th["pos_max"] < th["ang_max"]:                                  # :1531 (sic)
                            create_keyframe = True
            prev_avg = (num_tracked + (float(n_tracked_wrt_key) - 1.0) * prev_avg) / float(n_tracked_wrt_key)  # :1541
            out["decisions"].append((dist, ang, num_tracked, M_F, M_K))
        out["poses_wrt_keyframe"].append(T_rel.copy())
        out["tracked"].append(inl)
        if create_keyframe:                                                                       # :1551-1566
            n_tracked_wrt_key = 0
            prev_avg = 0.0
            reference = cur
            out["keyframe_ids"].append(idx)
            if T_key_list:
                T_key_list.append(T_key_list[-1] @ T_rel)
            else:
                T_key_list.append(T_curr.copy())
            create_keyframe = False
        out["poses_wrt_S"].append(T_curr.copy())
    return out
