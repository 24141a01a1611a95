"""Bearing-only RANSAC (sos_ransac_p3p) against oracle/p3p.py: the arguments the reference passes to OpenGV
(pose_est_tools.py:785, 915) — bearings of the current frame, 3D points of the reference frame — and nothing else."""
import numpy as np
import pytest
import torch

from oracle import p3p as op3p
from oracle.ransac import cayley_to_rot

pytestmark = pytest.mark.gpu

RIG = np.stack([np.hstack([np.eye(3), [[0.0], [0.0], [0.06]]]),
                np.hstack([cayley_to_rot([0.01, 0.02, 0.0]), [[0.01], [0.0], [-0.07]]])])


def _scene(rng, n, rig, outliers=0.3, noise=0.0):
    R, t = cayley_to_rot(rng.normal(0, 0.05, 3)), rng.normal(0, 0.1, 3)
    pb = rng.normal(0, 1, (n, 3))
    pb = pb / np.linalg.norm(pb, axis=1, keepdims=True) * rng.uniform(0.5, 7, (n, 1))       # body frame
    p_ref = (pb @ R.T + t).astype(np.float32)
    cam = rng.integers(0, 2, n).astype(np.uint8) if rig is not None else None
    x = pb if rig is None else np.einsum("nji,nj->ni", rig[cam][:, :, :3], pb - rig[cam][:, :, 3])
    x = x + rng.normal(0, noise, x.shape) * np.linalg.norm(x, axis=1, keepdims=True)
    f = x / np.linalg.norm(x, axis=1, keepdims=True)
    bad = rng.random(n) < outliers
    fb = rng.normal(0, 1, (n, 3))
    f[bad] = (fb / np.linalg.norm(fb, axis=1, keepdims=True))[bad]
    return p_ref, f.astype(np.float32), cam, np.hstack([R, t[:, None]]), ~bad


@pytest.mark.parametrize("noncentral", [False, True])
def test_p3p_ransac_matches_oracle(ctx, noncentral):
    rng = np.random.default_rng(11 + noncentral)
    rig = RIG if noncentral else None
    B, cap, H = 3, 600, 256
    thr = 1.0 - np.cos(np.radians(1.0))
    hyp = rng.integers(0, 2 ** 32, (H, 4), dtype=np.uint64).astype(np.uint32)
    ns = [600, 411, 3]                                                    # the last problem cannot be sampled
    p_ref = np.zeros((B, cap, 3), np.float32)
    f_cur = np.zeros((B, cap, 3), np.float32)
    cam = np.zeros((B, cap), np.uint8)
    truth = []
    for b, n in enumerate(ns):
        p, f, c, M, good = _scene(rng, n, rig, noise=2e-4)
        p_ref[b, :n], f_cur[b, :n] = p, f
        if c is not None:
            cam[b, :n] = c
        truth.append((M, good))
    dev = lambda a: torch.from_numpy(a).cuda()
    counts = torch.zeros((B, H), dtype=torch.int32, device="cuda")
    pose, best_hyp, best_count, mask, _ = ctx.ransac_p3p(
        dev(p_ref), dev(f_cur), torch.tensor(ns, dtype=torch.int32, device="cuda"), dev(hyp.view(np.int32)), thr,
        cam=dev(cam) if noncentral else None, rig=rig, n_cams=2 if noncentral else 0, all_counts=counts)
    pose, best_hyp, best_count = pose.cpu().numpy(), best_hyp.cpu().numpy(), best_count.cpu().numpy()
    mask, counts = mask.cpu().numpy().astype(bool), counts.cpu().numpy()
    for b, n in enumerate(ns):
        M, h, c, inl, cnt = op3p.ransac_p3p(p_ref[b, :n], f_cur[b, :n], cam[b, :n] if noncentral else None, rig, hyp, thr)
        if n < 4:
            assert best_count[b] < 0 and not mask[b].any()
            continue
        valid_o, valid_k = cnt >= 0, counts[b] >= 0
        assert (valid_o == valid_k).mean() > 0.99
        both = valid_o & valid_k
        assert both.sum() > H // 4
        assert (cnt[both] == counts[b][both]).mean() > 0.99            # independent quartic solvers: near-double roots may differ
        assert best_hyp[b] == h and best_count[b] == c
        np.testing.assert_array_equal(mask[b, :n], inl)
        assert not mask[b, n:].any()
        np.testing.assert_allclose(pose[b], M, atol=2e-6)
        # and it is the planted motion
        Mt, good = truth[b]
        np.testing.assert_allclose(pose[b], Mt, atol=5e-3)
        assert (mask[b, :n] & good).sum() >= 0.9 * good.sum()


def test_pyopengv_ransac_with_the_reference_arguments(ctx):
    """The reference's own call signatures (no points_cur): pose_est_tools.py:785 (non-central) and :915 (central)."""
    import pyopengv
    rng = np.random.default_rng(5)
    thr = 1.0 - np.cos(np.radians(1.0))
    p, f, c, M, good = _scene(rng, 500, RIG)
    T, inliers = pyopengv.absolute_pose_noncentral_ransac(f.astype(np.float64), c.astype(np.float64)[:, None], p.astype(np.float64),
                                                         RIG[:, :, 3], RIG[:, :, :3], thr, 210)
    assert T.shape == (3, 4) and abs(T - M).max() < 1e-4
    assert set(np.nonzero(good)[0]) <= set(inliers.tolist())
    p, f, _, M, good = _scene(rng, 500, None)
    T, inliers = pyopengv.absolute_pose_ransac(f.astype(np.float64), p.astype(np.float64), "KNEIP", thr, 210)
    assert abs(T - M).max() < 1e-4 and set(np.nonzero(good)[0]) <= set(inliers.tolist())
    T2 = pyopengv.absolute_pose_optimize_nonlinear(f[inliers].astype(np.float64), p[inliers].astype(np.float64), T[:, 3], T[:, :3])
    assert abs(T2 - M).max() < 2e-4                                  # float32 inputs, chance inliers among the outliers
