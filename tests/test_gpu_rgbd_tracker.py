"""GPU, BASELINE config 5: the RGB-D comparison path end to end (demo_vo_rgbd.py → RGBDFrame + TrackerRGBDSE3.track_frame,
pose_est_tools.py:404-623, 880-958) on a synthetic 640x480 depth map with 2000 keypoints: back-projection, the same Hamming
matcher + |Δu| gate, central-camera RANSAC on the seeded hypothesis list and the LM refinement — against the oracle
composition of the same stages."""
import cv2
import numpy as np
import pytest

from oracle import geometry, hamming, pipeline, ransac

pytestmark = pytest.mark.gpu

W, H = 640, 480


def synth_rgbd_pair(depth_is_Z, seed=0, n=2000):
    """Landmarks seen by two poses of a pinhole RGB-D camera; depth maps hold the landmark depth at the TRUNCATED pixel
    (the reference back-projects at the truncated pixel, pose_est_tools.py:612) over a 3 m background."""
    rng = np.random.default_rng(seed)
    fx = fy = 525.0 if depth_is_Z else 554.256258
    cx, cy = 319.5, 239.5
    ang = np.deg2rad(2.0)
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([0.04, -0.01, 0.03])
    # points in the reference camera; p_ref = R p_cur + t
    uv0 = np.stack([rng.uniform(40, W - 40, 3 * n), rng.uniform(40, H - 40, 3 * n)], 1)
    Z = rng.uniform(1.0, 6.0, 3 * n)
    P0 = np.stack([(uv0[:, 0] - cx) / fx * Z, (uv0[:, 1] - cy) / fy * Z, Z], 1)
    P1 = (P0 - t) @ R            # R^T (p - t)
    uv1 = np.stack([P1[:, 0] / P1[:, 2] * fx + cx, P1[:, 1] / P1[:, 2] * fy + cy], 1)
    ok = (uv1[:, 0] > 2) & (uv1[:, 0] < W - 2) & (uv1[:, 1] > 2) & (uv1[:, 1] < H - 2) & (P1[:, 2] > 0.9)
    # one landmark per truncated pixel in either image, so the sparse depth maps do not collide
    keep, seen0, seen1 = [], set(), set()
    for i in np.flatnonzero(ok):
        a, b = (int(uv0[i, 0]), int(uv0[i, 1])), (int(uv1[i, 0]), int(uv1[i, 1]))
        if a in seen0 or b in seen1:
            continue
        seen0.add(a); seen1.add(b); keep.append(i)
        if len(keep) == n:
            break
    keep = np.array(keep)
    desc = rng.integers(0, 256, (len(keep), 32), dtype=np.uint8)
    frames = []
    for uv, P in ((uv0[keep], P0[keep]), (uv1[keep], P1[keep])):
        depth = np.full((H, W), 3.0, np.float32)
        ui, vi = uv[:, 0].astype(int), uv[:, 1].astype(int)
        d = P[:, 2] if depth_is_Z else np.linalg.norm(
            np.stack([(ui - cx) / fx, (vi - cy) / fy, np.ones(len(ui))], 1), axis=1) * P[:, 2]
        depth[vi, ui] = d.astype(np.float32)
        flips = (rng.random((len(keep), 256)) < 0.04)
        dsc = np.packbits(np.unpackbits(desc, axis=1) ^ flips.astype(np.uint8), axis=1)
        order = rng.permutation(len(keep))
        kp = [cv2.KeyPoint(float(uv[j, 0]), float(uv[j, 1]), 7.0) for j in order]
        frames.append((depth, kp, dsc[order], uv[order].astype(np.float32)))
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    return dict(fx=fx, fy=fy, center_x=cx, center_y=cy, depth_is_Z=depth_is_Z, units="m"), frames, T


@pytest.mark.parametrize("depth_is_Z", [True, False], ids=["Z", "radial"])
def test_rgbd_tracker_pair_matches_oracle(ctx, depth_is_Z):
    import pyopengv
    from omnistereo.camera_models import RGBDCamModel
    from omnistereo import pose_est_tools
    kw, frames, T_true = synth_rgbd_pair(depth_is_Z, seed=3 if depth_is_Z else 4)
    cam = RGBDCamModel(**kw)
    tracker = pose_est_tools.TrackerRGBDSE3(cam)
    rgb = np.zeros((H, W, 3), np.uint8)
    fobj = [pose_est_tools.RGBDFrame(cam, rgb, d, i, features=(kp, dsc)) for i, (d, kp, dsc, _) in enumerate(frames)]
    ok, msg = tracker.track_frame(fobj[0], fobj[1])
    assert ok, msg
    T = fobj[1].T_frame_wrt_tracking_ref_frame

    # ---- oracle composition on the same inputs
    camd = dict(fx=kw["fx"], fy=kw["fy"], center_x=kw["center_x"], center_y=kw["center_y"], focal_length_m=1.0 / 1000.0,
                depth_is_Z=float(depth_is_Z))
    assert set(camd) == set(geometry.RGBD_FIELDS)
    st = []
    for depth, kp, dsc, uv in frames:
        u, v = uv[:, 0].astype(np.uint).astype(np.int32), uv[:, 1].astype(np.uint).astype(np.int32)
        xyz, bearing, valid = geometry.rgbd_backproject(camd, depth, u, v, 0.8, 7.0)
        st.append(dict(xyz=xyz[valid], b=bearing[valid], desc=dsc[valid], uv=uv[valid]))
        assert valid.sum() > 1500
    assert fobj[0].num_valid_keypoints == len(st[0]["xyz"]) and fobj[1].num_valid_keypoints == len(st[1]["xyz"])
    np.testing.assert_allclose(fobj[1].keypoints_3D_points, st[1]["xyz"], rtol=1e-4, atol=1e-6)   # fp32 kernel vs float64
    qi, ti = pipeline._bf_sorted(st[1]["desc"], st[0]["desc"])
    gate = hamming.filter_pixel_correspondences(st[0]["uv"][ti], st[1]["uv"][qi], -1, tracker.max_horizontal_diff_f2f_matches)
    qi, ti = qi[gate], ti[gate]
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    hyp = pyopengv.hypothesis_list(tracker.max_ransac_iterations_3D_to_2D, tracker.ransac_seed)
    # the tracker feeds the kernels float32 copies of ITS back-projected points; use those for the bit-exact part
    p_ref, p_cur = f32(fobj[0].keypoints_3D_points[ti]), f32(fobj[1].keypoints_3D_points[qi])
    b_cur = f32(fobj[1].bearing_vectors[qi])
    o = ransac.ransac_p3d(p_ref, p_cur, hyp, "bearing", tracker.backprojection_score_threshold_3D_to_2D, f_cur=b_cur)
    assert o["margin"] > 1e-9
    assert tracker.num_tracked_correspondences == o["best_count"]
    m = o["mask"]
    want, _, _ = ransac.refine_pose_lm(p_ref[m], b_cur[m], o["pose"])
    np.testing.assert_allclose(T[:3], want, rtol=0, atol=1e-6)
    # and the recovered motion is the planted one up to the pixel-truncation noise of the depth lookup
    assert np.allclose(T[:3, :3], T_true[:3, :3], atol=5e-3) and np.allclose(T[:3, 3], T_true[:3, 3], atol=2e-2)
