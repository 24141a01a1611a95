"""`pyopengv`-compatible module backed by the batched RANSAC kernels (SURVEY §8b, row "L3 -> OpenGV").

The reference imports seven names from `pyopengv` (pose_est_tools.py:42,78,647-649).  OpenGV is neither vendored nor
installed and its outputs are pinned by no test of the reference, so parity with OpenGV itself is UNPINNED; the
semantics implemented here are the ones written down in include/sosfront.h:

  * hypotheses, with the arguments the reference passes (bearings of the current frame + 3D points of the reference frame):
    a three-point absolute-pose solver on 3 + 1 sampled correspondences (sos_ransac_p3p: Grunert's quartic for a central
    camera, Newton on the distance equations for the two-viewpoint rig, the 4th sample picks among the solutions) - the
    problem OpenGV's KNEIP / GP3P solvers solve, by other code.  "EPNP" (the reference's central default, a 6-point fit
    inside OpenGV's RANSAC) is served by the same minimal solver.
  * hypotheses, when the caller also passes the 3D points of the current frame (`points_cur=`, which the mirrored trackers in
    omnistereo.pose_est_tools do - the reference has them at the call site, pose_est_tools.py:753-756): Arun / Kabsch rigid
    registration on 3 sampled 3D-3D correspondences (transformations.py:874-1030), what north_star asks for.
  * score: OpenGV's bearing-angle score 1 - f . reprojection, threshold and iteration budget as given.
  * first maximum wins; the inlier indices are returned ascending, as the caller assumes (pose_est_tools.py:787).
  * *_optimize_nonlinear: OpenGV's published algorithm — Levenberg-Marquardt on one residual 1 - f . reprojection per
    correspondence — run by sos_refine_pose (csrc/refine.cu); checked against scipy's MINPACK driver in the tests.
"""
import numpy as np
import torch

from . import ops

_SEED = 0


def _ctx():
    return ops.default_context()


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def hypothesis_list(n_hyp: int, seed: int = _SEED, k: int = 3) -> np.ndarray:
    """The shared seeded hypothesis list: uint32 [n_hyp, k]; sample j of hypothesis h is row floor(u * n / 2^32)."""
    return np.random.default_rng(seed).integers(0, 2 ** 32, (int(n_hyp), int(k)), dtype=np.uint64).astype(np.uint32)


def _rig(cam_offsets, cam_rotations):
    off = np.asarray(cam_offsets, np.float64).reshape(-1, 3)
    rot = np.asarray(cam_rotations, np.float64).reshape(-1, 3, 3)
    if len(off) > 2:
        raise NotImplementedError("the SOS rig has two cameras")
    return np.concatenate([rot, off[:, :, None]], axis=2), len(off)


def _ransac(bearings, points, points_cur, threshold, max_iterations, cam=None, rig=None, n_cams=0, seed=_SEED):
    b = np.asarray(bearings, np.float32)[:, :3]
    p = np.asarray(points, np.float32)[:, :3]
    n = len(p)
    ctx = _ctx()
    n_dev = torch.tensor([n], dtype=torch.int32, device=ctx.device)
    cam_dev = None if cam is None else _dev(np.asarray(cam).reshape(-1).astype(np.uint8)[None])
    if points_cur is None:
        hyp = hypothesis_list(max_iterations, seed, 4)
        pose, best_hyp, best_count, mask, _ = ctx.ransac_p3p(_dev(p[None]), _dev(b[None]), n_dev, _dev(hyp.view(np.int32)),
                                                             float(threshold), cam=cam_dev, rig=rig, n_cams=n_cams)
    else:
        pc = np.asarray(points_cur, np.float32)[:, :3]
        hyp = hypothesis_list(max_iterations, seed)
        pose, best_hyp, best_count, mask, _ = ctx.ransac_p3d(
            _dev(p[None]), _dev(pc[None]), n_dev, _dev(hyp.view(np.int32)), ops.SCORE_BEARING, float(threshold),
            f_cur=_dev(b[None]), cam=cam_dev, rig=rig, n_cams=n_cams)
    T = pose.cpu().numpy()[0].astype(np.float64)
    inliers = np.nonzero(mask.cpu().numpy()[0])[0].astype(np.int64)
    return T, inliers


def absolute_pose_noncentral_ransac(bearing_vectors, cam_correspondences, points, cam_offsets, cam_rotations, threshold,
                                    max_iterations, points_cur=None, seed=_SEED):
    """-> (T 3x4 [R|t] of the current frame wrt the reference frame, ascending inlier indices): pose_est_tools.py:785."""
    rig, n_cams = _rig(cam_offsets, cam_rotations)
    return _ransac(bearing_vectors, points, points_cur, threshold, max_iterations, cam=cam_correspondences, rig=rig,
                   n_cams=n_cams, seed=seed)


def absolute_pose_ransac(bearing_vectors, points, algo_name, threshold, max_iterations, points_cur=None, seed=_SEED):
    """Central camera (RGB-D path, pose_est_tools.py:915); every `algo_name` ("EPNP", "KNEIP", "GAO") runs the same solver."""
    return _ransac(bearing_vectors, points, points_cur, threshold, max_iterations, seed=seed)


def _refine(bearings, points, t, R, cam=None, rig=None, max_iters=60):
    """Levenberg-Marquardt on sum (1 - f . reprojection)^2 over ALL rows given (the caller passes the inliers),
    started at (t, R): sos_refine_pose.  float64 result."""
    b = np.asarray(bearings, np.float32)[:, :3]
    p = np.asarray(points, np.float32)[:, :3]
    n = len(p)
    pose0 = np.hstack([np.asarray(R, np.float64).reshape(3, 3), np.asarray(t, np.float64).reshape(3, 1)])
    if n == 0:
        return pose0
    ctx = _ctx()
    _, pose64, _ = ctx.refine_pose(
        _dev(p[None]), _dev(b[None]), None if cam is None else _dev(np.asarray(cam).reshape(-1).astype(np.uint8)[None]),
        None, torch.tensor([n], dtype=torch.int32, device=ctx.device), rig, _dev(pose0.astype(np.float32)[None]),
        max_iters=max_iters)
    return pose64.cpu().numpy()[0]


def absolute_pose_noncentral_optimize_nonlinear(bearing_vectors, cam_correspondences, points, cam_offsets, cam_rotations, t, R,
                                                points_cur=None):
    """pose_est_tools.py:830 -> refined T 3x4.  OpenGV's algorithm (LM on the bearing residual over the rows given);
    `points_cur` is accepted for symmetry with the RANSAC call and not needed."""
    rig, _ = _rig(cam_offsets, cam_rotations)
    return _refine(bearing_vectors, points, t, R, cam=cam_correspondences, rig=rig)


def absolute_pose_optimize_nonlinear(bearing_vectors, points, t, R, points_cur=None):
    """Central camera (pose_est_tools.py:937)."""
    return _refine(bearing_vectors, points, t, R)


def relative_pose_ransac(*args, **kwargs):
    raise NotImplementedError("2D-2D relative pose is not on the SOS / RGB-D hot path (pose_est_tools.py:78 is a disabled branch)")


def triangulation_triangulate(*args, **kwargs):
    raise NotImplementedError("OpenGV triangulation is a disabled branch (use_opengv_triangulation = False, pose_est_tools.py:284)")


triangulation_triangulate2 = triangulation_triangulate
