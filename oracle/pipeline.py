"""CPU restatement of one SOS front-end frame / frame pair, chaining the oracle stages the way the reference's
StereoPanoramicFrame.establish_stereo_correspondences (pose_est_tools.py:320-402) and TrackerStereoSE3.track_frame
(pose_est_tools.py:736-847) chain them.  TEST INFRASTRUCTURE ONLY: used as the checker in tests and as the timed CPU
baseline of bench.py (`cpu_baseline`, `--impl reference`).  It makes the same OpenCV calls the reference makes
(cv2.bitwise_and, cv2.remap, cv2.BFMatcher.match + sorted) and NumPy float64 everywhere else; RANSAC is the float64
restatement of oracle/ransac.py because OpenGV is not available (parity unpinned, see oracle/__init__.py)."""
from __future__ import annotations

import cv2
import numpy as np

import time
from contextlib import contextmanager

from . import geometry, hamming, ransac


@contextmanager
def _stage(timer, name):
    """Accumulate the wall-clock time of one stage into timer[name] (bench.py's per-stage CPU split); no-op without a timer."""
    if timer is None:
        yield
        return
    t0 = time.perf_counter()
    try:
        yield
    finally:
        timer[name] = timer.get(name, 0.0) + time.perf_counter() - t0


def fully_masked_images(omni, masks, background_state, color_RGB=(0, 0, 0)):
    """OmniStereoModel.get_fully_masked_images as set_current_omni_image(apply_mask=True, mask_RGB=(0,0,0)) runs it per
    frame once the masks are cached (camera_models.py:2990-3010), statement by statement — including the two float64
    np.zeros(omni.shape) temporaries the reference allocates and never uses (cv2 ignores a dst of the wrong type) and the
    background repaint through the inverted masks.  Semantically: dst = mask ? src : background."""
    out = []
    if background_state.get("color") != color_RGB:                       # :2998-3002, first frame only
        background_state["color"] = color_RGB
        bg = np.zeros_like(omni)
        bg[:, :, :] += np.array((color_RGB[2], color_RGB[1], color_RGB[0]), dtype="uint8")
        background_state["img"] = bg
    bg = background_state["img"]
    for which in ("top", "bot"):
        masked = np.zeros(omni.shape)                                      # :2991 / :2994 (float64, 8 bytes per sample)
        masked = cv2.bitwise_and(src1=omni, src2=omni, dst=masked, mask=masks[which])   # :2992 / :2995
        inv = cv2.bitwise_not(src=masks[which])                            # :3004 / :3008
        masked = cv2.bitwise_and(src1=bg, src2=bg, dst=masked, mask=inv)   # :3006 / :3010
        out.append(masked)
    return out


def remap_views(omni, maps, masks, background_state=None, timer=None):
    """set_current_omni_image(apply_mask=True, mask_RGB=(0,0,0)): per view bitwise_and with the mirror mask
    (camera_models.py:2990-2995) then Panorama.get_panoramic_image: float64 -> float32 cast of both maps every frame and
    cv2.remap (panorama.py:291-298)."""
    out = []
    if background_state is not None:
        with _stage(timer, "mask"):
            masked_views = fully_masked_images(omni, masks, background_state)
    else:
        masked_views = [cv2.bitwise_and(omni, omni, mask=masks[which]) for which in ("top", "bot")]
    with _stage(timer, "remap"):
        for masked, which in zip(masked_views, ("top", "bot")):
            mx, my = maps[which]
            out.append(cv2.remap(masked, mx.astype("float32"), my.astype("float32"), cv2.INTER_LINEAR, None,
                                 cv2.BORDER_CONSTANT, (0, 0, 0)))
    return out


def _bf_sorted(q, t):
    """FeatureMatcher.match, k_best = 1 (camera_models.py:404-446): BFMatcher.match then Python sorted by distance."""
    m = cv2.BFMatcher(normType=cv2.NORM_HAMMING).match(queryDescriptors=q, trainDescriptors=t)
    m = sorted(m, key=lambda x: x.distance)
    return (np.fromiter((x.queryIdx for x in m), np.int64, len(m)), np.fromiter((x.trainIdx for x in m), np.int64, len(m)))


def stereo_frame(pano_g, f_top, f_bot, px_top, desc_top, boff_top, px_bot, desc_bot, boff_bot, max_du=2.5, min_dv=1.0,
                 min_range=0.5, max_range=7.0, cap=None, timer=None):
    """match_features_panoramic_top_bottom (camera_models.py:3027-3101) per bucket, then lifting, midpoint triangulation
    and the (homogeneous-norm) range gate of establish_stereo_correspondences (pose_est_tools.py:344-397)."""
    rq, rt = [], []
    with _stage(timer, "match_stereo"):
        for k in range(len(boff_top) - 1):
            q0, q1, t0, t1 = boff_bot[k], boff_bot[k + 1], boff_top[k], boff_top[k + 1]
            if q1 <= q0 or t1 <= t0:
                continue
            qi, ti = _bf_sorted(desc_bot[q0:q1], desc_top[t0:t1])
            rq.append(q0 + qi)
            rt.append(t0 + ti)
        rq = np.concatenate(rq) if rq else np.zeros(0, np.int64)
        rt = np.concatenate(rt) if rt else np.zeros(0, np.int64)
        ok = hamming.filter_pixel_correspondences(px_top[rt], px_bot[rq], min_dv, max_du)
        rq, rt = rq[ok], rt[ok]
    t_lift = time.perf_counter()
    az1, el1 = geometry.pano_pixel_to_angles(pano_g, px_top[rt])
    az2, el2 = geometry.pano_pixel_to_angles(pano_g, px_bot[rq])
    b_top = geometry.angles_to_sphere(az1, el1)
    b_bot = geometry.angles_to_sphere(az2, el2)
    xyz = geometry.triangulate_midpoint(az1, el1, az2, el2, f_top, f_bot)
    keep = geometry.range_filter(np.hstack([xyz, np.ones((len(xyz), 1))]), min_range, max_range)
    if cap is not None:
        keep &= np.cumsum(keep) <= cap
    if timer is not None:
        timer["lift_triangulate"] = timer.get("lift_triangulate", 0.0) + time.perf_counter() - t_lift
    return dict(uv_top=px_top[rt][keep], uv_bot=px_bot[rq][keep], b_top=b_top[keep], b_bot=b_bot[keep], xyz=xyz[keep],
                desc_top=desc_top[rt][keep], desc_bot=desc_bot[rq][keep])


def track_pair(ref, cur, hyp, mode, threshold, rig, max_du, hyp_limit=None, refine="arun", timer=None):
    """match_features_frame_to_frame for both views (pose_est_tools.py:741-749), stacking (:752-778), RANSAC + refit."""
    parts = []
    with _stage(timer, "match_temporal"):
        for view, (uvk, dk, bk) in enumerate((("uv_top", "desc_top", "b_top"), ("uv_bot", "desc_bot", "b_bot"))):
            if len(cur[dk]) == 0 or len(ref[dk]) == 0:
                continue
            qi, ti = _bf_sorted(cur[dk], ref[dk])
            ok = hamming.filter_pixel_correspondences(ref[uvk][ti], cur[uvk][qi], -1, max_du)
            qi, ti = qi[ok], ti[ok]
            parts.append((ref["xyz"][ti], cur["xyz"][qi], cur[bk][qi], np.full(len(qi), view, np.uint8)))
    if not parts:
        return None
    p_ref, p_cur, f_cur, cam = (np.concatenate(x) for x in zip(*parts))
    f32 = lambda a: a.astype(np.float32)
    h = hyp if hyp_limit is None else hyp[:hyp_limit]
    with _stage(timer, "ransac"):
        o = ransac.ransac_p3d(f32(p_ref), f32(p_cur), h, mode, threshold, f_cur=f32(f_cur), cam=cam, rig=rig)
    with _stage(timer, "refine"):
        if o["best_hyp"] >= 0:
            if refine == "lm":  # pose_est_tools.py:824-834: non-linear refinement on the inliers, started at the RANSAC pose
                o["refit"] = ransac.refine_pose_lm(f32(p_ref), f32(f_cur), o["pose"], cam, rig, o["mask"])[0]
            elif refine == "arun":
                o["refit"] = ransac.refit(p_ref, p_cur, o["mask"])
            else:
                o["refit"] = o["pose"]
    o["n_corr"] = len(p_ref)
    return o
