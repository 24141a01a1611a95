"""Host logic of the batched VO driver, no GPU needed: pose bookkeeping against the reference's own
functions (golden driver.npz), the keyframe decision tree and pose chaining against the sequential oracle loop."""
import io

import numpy as np

from conftest import load_golden
from oracle import driver as odriver
from vo_single_camera_sos_b200 import driver


def test_pose_bookkeeping_matches_reference_golden():
    g = load_golden("driver.npz")
    for i, T in enumerate(g["T"]):
        for impl in (driver.quaternion_wxyz, odriver.quaternion_from_matrix):
            q = impl(T)
            assert np.allclose(q, g["quat"][i], atol=1e-12) or np.allclose(-q, g["quat"][i], atol=1e-12)
            assert q[0] >= 0
        assert driver.translation_metric(T) == g["dist"][i] == odriver.rpe_translation_metric(T)
        assert driver.rotation_metric(T) == g["angle"][i] == odriver.rpe_rotation_metric(T)
        want = " ".join(str(v) for v in (i, *g["trans"][i], g["quat"][i][1], g["quat"][i][2], g["quat"][i][3], g["quat"][i][0]))
        got = driver.tum_line(i, T).split()
        assert got[0] == str(i) and np.allclose([float(x) for x in got[1:]], [float(x) for x in want.split()[1:]], atol=1e-12)
        assert odriver.tum_line(i, T).split()[0] == str(i)
    for i in range(39):
        assert np.allclose(g["T"][i] @ g["T"][i + 1], g["chain"][i], atol=1e-15)


def canned_sequence(rng, n):
    """Per-frame tracking results wrt whatever the current keyframe is: small motions that accumulate until a
    threshold trips.  Returned as a function of (keyframe index, frame index) so both drivers see the same numbers."""
    steps = [np.eye(4)]
    for _ in range(n):
        ang = rng.normal() * 0.004
        T = np.eye(4)
        T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
        T[:3, 3] = rng.normal(size=3) * 0.004 + [0.003, 0, 0]
        steps.append(steps[-1] @ T)
    inl = rng.integers(40, 400, n + 1)
    kps = rng.integers(100, 500, n + 1)
    return steps, inl, kps


def test_tracking_state_and_policy_match_sequential_oracle():
    rng = np.random.default_rng(3)
    for thresholds in (None, dict(odriver.INDOOR, pos_min=0.02), dict(odriver.INDOOR, pos_min=0.5, ang_min=0.005, ang_max=0.05, pos_max=0.01)):
        n = 60
        world, inl, kps = canned_sequence(rng, n)
        frames = list(range(n + 1))
        frame_fn = lambda f: dict(idx=f, xyz=np.zeros((kps[f], 3)))

        def track_fn(ref, cur):
            T = np.linalg.inv(world[ref["idx"]]) @ world[cur["idx"]]
            return dict(refit=T[:3], n_corr=int(inl[cur["idx"]]) + 10, best_hyp=0, best_count=int(inl[cur["idx"]]))

        want = odriver.run_vo(frames, thresholds=thresholds, frame_fn=frame_fn, track_fn=track_fn)
        th = odriver.INDOOR if thresholds is None else thresholds
        st = driver.TrackingState(driver.KeyframePolicy(**th))
        key = 0
        st.first_frame(0, int(kps[0]))
        for f in frames[1:]:
            T = np.linalg.inv(world[key]) @ world[f]
            if st.tracked_frame(f, T[:3], int(inl[f]), int(kps[f])):
                key = f
        r = st.result
        assert r.keyframe_ids == want["keyframe_ids"]
        assert len(r.keyframe_ids) > 1 or thresholds is not None
        assert np.allclose(np.array(r.poses_wrt_S), np.array(want["poses_wrt_S"]), atol=1e-12)
        assert np.allclose(np.array(r.poses_wrt_keyframe), np.array(want["poses_wrt_keyframe"]), atol=1e-12)
        assert np.allclose(np.array(r.poses_wrt_S), np.array(world), atol=1e-9)   # exact relative poses chain back
        assert r.tracked[1:] == [int(x) for x in inl[1:]]


def test_rotation_branch_quirk_is_reproduced():
    """pose_est_tools.py:1531 compares pos_max with ang_max; with the reference's thresholds the rotation-only branch
    never creates a keyframe.  Both implementations keep that behaviour."""
    p = driver.KeyframePolicy()
    assert not p.wants_keyframe(0.001, np.deg2rad(5.0), 100, 50, 300, 300)      # rotation only: no keyframe
    assert p.wants_keyframe(0.05, np.deg2rad(5.0), 100, 50, 300, 300)           # translation in range
    assert not p.wants_keyframe(0.05, np.deg2rad(11.0), 100, 50, 300, 300)      # "crazy" rotation
    assert not p.wants_keyframe(0.05, 0.0, 4, 50, 300, 300)                     # too few tracked correspondences
    assert not p.wants_keyframe(0.05, 0.0, 100, 50, 20, 300)                    # too few keypoints
    q = driver.KeyframePolicy(pos_max=0.01, ang_max=0.05, ang_min=0.005, pos_min=0.5)
    assert q.wants_keyframe(0.001, 0.01, 100, 50, 300, 300)                     # only when pos_max < ang_max
