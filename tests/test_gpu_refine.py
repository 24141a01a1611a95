"""GPU parity: Levenberg-Marquardt pose refinement (sos_refine_pose) vs the scipy/MINPACK oracle on
OpenGV's Cayley parameters.  Both minimise sum (1 - f . reprojection)^2 in float64 from the same float32 inputs, so the
minimisers must agree far inside the 1e-4 relative pose tolerance; the tolerance used is written at each assert."""
import numpy as np
import pytest
import torch

from oracle import ransac
from test_gpu_ransac import dev, make_problem

pytestmark = pytest.mark.gpu

RIG = np.zeros((2, 3, 4))
RIG[:, :, :3] = np.eye(3)
RIG[0, 2, 3] = 0.06
RIG[1, 2, 3] = -0.07


def noisy_start(rng, p_ref, p_cur, mask):
    """A RANSAC-like start: Arun on three inlier rows."""
    rows = rng.choice(np.flatnonzero(mask), 3, replace=False)
    return ransac.superimposition(p_cur[rows].astype(np.float64).T, p_ref[rows].astype(np.float64).T)[:3]


def build(rng, sizes, cap, rig):
    B = len(sizes)
    P = np.zeros((B, cap, 3), np.float32)
    F = np.zeros((B, cap, 3), np.float32)
    C = np.zeros((B, cap), np.uint8)
    M = np.zeros((B, cap), np.uint8)
    pose0 = np.zeros((B, 3, 4), np.float32)
    for b, n in enumerate(sizes):
        p_ref, p_cur, f, cam = make_problem(rng, n, inlier_frac=0.6, sigma=0.004, rig=rig)
        f = f + rng.normal(0, 1e-3, f.shape).astype(np.float32)
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        if rig is None:
            cam[:] = 0
        truth = np.linalg.norm(p_ref.astype(np.float64) - p_cur, axis=1) < 0.5  # planted inliers (outliers are far away)
        mask = truth & (rng.random(n) < 0.9)
        P[b, :n], F[b, :n], C[b, :n], M[b, :n] = p_ref, f, cam, mask
        pose0[b] = noisy_start(rng, p_ref, p_cur, mask)
    return P, F, C, M, pose0


@pytest.mark.parametrize("rig", [RIG, None], ids=["noncentral", "central"])
def test_refine_matches_minpack(ctx, rig):
    rng = np.random.default_rng(11 if rig is None else 12)
    sizes = [40, 700, 5000]
    cap = 5120
    P, F, C, M, pose0 = build(rng, sizes, cap, rig)
    n = dev(np.array(sizes, np.int32))
    results = {}
    for S in (1, 4, 8):
        pose, pose64, stats = ctx.refine_pose(dev(P), dev(F), dev(C), dev(M), n, rig, dev(pose0), max_iters=60,
                                              cluster_size=S)
        results[S] = (pose.cpu().numpy(), pose64.cpu().numpy(), stats.cpu().numpy())
    _, pose64, stats = results[8]
    for b, nb in enumerate(sizes):
        want, c0, c1 = ransac.refine_pose_lm(P[b, :nb], F[b, :nb], pose0[b], C[b, :nb], rig, M[b, :nb])
        assert stats[b, 3] == M[b, :nb].sum()
        np.testing.assert_allclose(stats[b, 0], c0, rtol=1e-4)  # float32 pose_in is orthonormalised differently
        assert stats[b, 1] <= c1 * (1 + 1e-9), "LM stopped above MINPACK's minimum"
        assert stats[b, 1] < stats[b, 0]
        # minimiser: float64 on both sides; 1e-7 absolute on R entries and on t [model units]
        np.testing.assert_allclose(pose64[b], want, rtol=0, atol=1e-7)
        R = pose64[b, :, :3]
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
        np.testing.assert_allclose(results[8][0][b], want, rtol=0, atol=2e-7)  # float32 output
        for S in (1, 4):  # the cluster size only changes the summation order
            np.testing.assert_allclose(results[S][1][b], pose64[b], rtol=0, atol=1e-9)


def test_refine_without_mask_and_passthrough(ctx):
    rng = np.random.default_rng(5)
    sizes = [300, 4]  # second problem: fewer than 6 rows -> pose_in passed through
    cap = 512
    P, F, C, M, pose0 = build(rng, [300, 300], cap, RIG)
    n = dev(np.array(sizes, np.int32))
    pose, pose64, stats = ctx.refine_pose(dev(P), dev(F), dev(C), None, n, RIG, dev(pose0), max_iters=80)
    pose64 = pose64.cpu().numpy()
    # no mask: all 300 rows incl. the far outliers enter the cost, as OpenGV would do with them
    want, c0, c1 = ransac.refine_pose_lm(P[0, :300], F[0, :300], pose0[0], C[0, :300], RIG, None)
    s = stats.cpu().numpy()
    assert s[0, 3] == 300 and s[0, 1] <= c1 * (1 + 1e-9)
    np.testing.assert_allclose(pose64[0], want, rtol=0, atol=1e-6)
    np.testing.assert_array_equal(pose.cpu().numpy()[1], pose0[1])
    assert s[1, 3] == 4


def test_pyopengv_optimize_nonlinear_signature_and_result(ctx):
    """The reference's call (pose_est_tools.py:830, :937): float64 arrays in, T 3x4 float64 out."""
    import pyopengv
    rng = np.random.default_rng(21)
    P, F, C, M, pose0 = build(rng, [900], 1024, RIG)
    m = M[0, :900].astype(bool)
    p, f, cam = P[0, :900][m].astype(np.float64), F[0, :900][m].astype(np.float64), C[0, :900][m]
    t0, R0 = pose0[0][:, 3].astype(np.float64), pose0[0][:, :3].astype(np.float64)
    T = pyopengv.absolute_pose_noncentral_optimize_nonlinear(f, cam.reshape(-1, 1), p, RIG[:, :, 3], RIG[:, :, :3], t0, R0)
    assert T.shape == (3, 4) and T.dtype == np.float64
    want, _, _ = ransac.refine_pose_lm(p, f, pose0[0], cam, RIG, None)
    np.testing.assert_allclose(T, want, rtol=0, atol=1e-7)
    Tc = pyopengv.absolute_pose_optimize_nonlinear(f, p, t0, R0)
    want_c, _, _ = ransac.refine_pose_lm(p, f, pose0[0], None, None, None)
    np.testing.assert_allclose(Tc, want_c, rtol=0, atol=1e-7)
