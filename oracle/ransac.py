"""Oracle for step 5 (RANSAC over 3D-3D Arun hypotheses). TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

PARITY UNPINNED against OpenGV (not vendored, not installed, no RANSAC output pinned by the reference).  The solver
follows transformations.superimposition_matrix, which IS pinned (tests/golden/arun.npz); the score follows the
reference's own restatement of OpenGV's score (pose_est_tools.py:150-203); the loop semantics are the ones written
in include/sosfront.h (shared seeded hypothesis list, first maximum wins)."""
from __future__ import annotations

import numpy as np

DEGENERATE_SIN2 = 1e-12  # same constant as csrc/ransac.cu::triangle_degenerate


def superimposition(v0: np.ndarray, v1: np.ndarray) -> np.ndarray:
    """transformations.affine_matrix_from_points(v0, v1, shear=False, scale=False, usesvd=True)
    (transformations.py:914-980) for 3 x K arrays: returns the 4 x 4 matrix M with v1 ~ M v0."""
    v0 = np.array(v0, np.float64, copy=True)
    v1 = np.array(v1, np.float64, copy=True)
    t0 = -np.mean(v0, axis=1)
    t1 = -np.mean(v1, axis=1)
    v0 += t0.reshape(3, 1)
    v1 += t1.reshape(3, 1)
    u, s, vh = np.linalg.svd(np.dot(v1, v0.T))
    R = np.dot(u, vh)
    if np.linalg.det(R) < 0.0:
        R -= np.outer(u[:, 2], vh[2, :] * 2.0)
    M = np.identity(4)
    M[:3, :3] = R
    M0 = np.identity(4)
    M0[:3, 3] = t0
    M1 = np.identity(4)
    M1[:3, 3] = t1
    M = np.dot(np.linalg.inv(M1), np.dot(M, M0))
    return M / M[3, 3]


def arun_batch(v0: np.ndarray, v1: np.ndarray) -> np.ndarray:
    """Batched form of `superimposition` for [S, k, 3] point sets -> [S, 3, 4] ([R|t], v1 ~ R v0 + t)."""
    v0 = np.asarray(v0, np.float64)
    v1 = np.asarray(v1, np.float64)
    c0 = v0.mean(axis=1, keepdims=True)
    c1 = v1.mean(axis=1, keepdims=True)
    a, b = v0 - c0, v1 - c1
    Hm = np.einsum("ski,skj->sij", b, a)  # dot(v1, v0.T)
    u, s, vh = np.linalg.svd(Hm)
    R = u @ vh
    neg = np.linalg.det(R) < 0.0
    R[neg] -= 2.0 * u[neg][:, :, 2:3] * vh[neg][:, 2:3, :]
    t = c1[:, 0, :] - np.einsum("sij,sj->si", R, c0[:, 0, :])
    return np.concatenate([R, t[:, :, None]], axis=2)


def sample_rows(hyp: np.ndarray, n: int) -> np.ndarray:
    """Row indices of the hypothesis samples: floor(hyp * n / 2^32) (include/sosfront.h, sos_ransac_p3d)."""
    return ((hyp.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def triangle_degenerate(p: np.ndarray) -> np.ndarray:
    """[S,3,3] triangles -> bool[S]: sin^2 of the angle at vertex 0 <= 1e-12 (collinear / coincident samples)."""
    e1 = p[:, 1] - p[:, 0]
    e2 = p[:, 2] - p[:, 0]
    c = np.cross(e1, e2)
    a2 = np.sum(c * c, axis=1)
    return ~(a2 > DEGENERATE_SIN2 * np.sum(e1 * e1, axis=1) * np.sum(e2 * e2, axis=1))


def score_euclid(M: np.ndarray, p_ref: np.ndarray, p_cur: np.ndarray) -> np.ndarray:
    """|p_ref - (R p_cur + t)| per correspondence."""
    return np.linalg.norm(p_ref - (p_cur @ M[:3, :3].T + M[:3, 3]), axis=1)


def score_bearing(M: np.ndarray, p_ref: np.ndarray, f_cur: np.ndarray, cam: np.ndarray | None = None,
                  rig: np.ndarray | None = None) -> np.ndarray:
    """1 - f . normalize(Rc^T (R^T (p - t) - tc)): get_selected_distances_to_model's absolute-pose branch
    (pose_est_tools.py:150-203) with the non-central camera correction quoted in its comment (:181-185)."""
    R, t = M[:3, :3], M[:3, 3]
    body = (p_ref - t) @ R  # R^T (p - t), row-vector form
    if rig is not None:
        rig = np.asarray(rig, np.float64).reshape(-1, 3, 4)
        c = np.zeros(len(p_ref), np.int64) if cam is None else np.asarray(cam).astype(np.int64).reshape(-1)
        Rc = rig[c, :, :3]
        tc = rig[c, :, 3]
        body = np.einsum("nji,nj->ni", Rc, body - tc)
    with np.errstate(invalid="ignore", divide="ignore"):
        reproj = body / np.linalg.norm(body, axis=1, keepdims=True)
    return 1.0 - np.sum(f_cur * reproj, axis=1)


def ransac_p3d(p_ref, p_cur, hyp, mode: str, threshold: float, f_cur=None, cam=None, rig=None, hyp_chunk: int = 256):
    """One RANSAC problem over a shared seeded hypothesis list.

    Returns dict(pose [3,4], best_hyp, best_count, mask [n] bool, counts [H] int64 (-1 = rejected sample),
    margin = min |score - thr| / thr over all correspondences of the best hypothesis)."""
    p_ref = np.asarray(p_ref, np.float64)
    p_cur = np.asarray(p_cur, np.float64)
    n = len(p_ref)
    H = len(hyp)
    counts = np.full(H, -1, np.int64)
    poses = np.full((H, 3, 4), np.nan)
    if n >= 3 and H > 0:
        rows = sample_rows(np.asarray(hyp), n)
        ok = (rows[:, 0] != rows[:, 1]) & (rows[:, 0] != rows[:, 2]) & (rows[:, 1] != rows[:, 2])
        ok &= ~triangle_degenerate(p_cur[rows]) & ~triangle_degenerate(p_ref[rows])
        idx = np.nonzero(ok)[0]
        if len(idx):
            poses[idx] = arun_batch(p_cur[rows[idx]], p_ref[rows[idx]])
        for s in range(0, len(idx), hyp_chunk):
            hs = idx[s:s + hyp_chunk]
            for h in hs:
                sc = _score(poses[h], mode, p_ref, p_cur, f_cur, cam, rig)
                counts[h] = int(np.count_nonzero(sc < threshold))
    best = int(np.argmax(counts)) if H > 0 and counts.max() >= 0 else -1
    out = dict(counts=counts, best_hyp=best, best_count=int(counts[best]) if best >= 0 else -1)
    if best >= 0:
        sc = _score(poses[best], mode, p_ref, p_cur, f_cur, cam, rig)
        out["pose"] = poses[best]
        out["mask"] = sc < threshold
        with np.errstate(invalid="ignore"):
            out["margin"] = float(np.nanmin(np.abs(sc - threshold))) / threshold
    else:
        out["pose"] = np.full((3, 4), np.nan)
        out["mask"] = np.zeros(n, bool)
        out["margin"] = np.inf
    return out


def _score(M, mode, p_ref, p_cur, f_cur, cam, rig):
    if mode == "euclid":
        return score_euclid(M, p_ref, p_cur)
    if mode == "bearing":
        return score_bearing(M, p_ref, f_cur, cam, rig)
    raise ValueError(mode)


def refit(p_ref, p_cur, mask) -> np.ndarray:
    """Arun on the inlier set (approximates *_optimize_nonlinear)."""
    m = np.asarray(mask, bool)
    return superimposition(np.asarray(p_cur, np.float64)[m].T, np.asarray(p_ref, np.float64)[m].T)[:3]


def num_iterations(outlier_fraction: float = 0.65, n_points: int = 3, p: float = 0.998) -> int:
    """TrackerSE3.compute_num_of_iterations_RANSAC (pose_est_tools.py:709-720) -> 210 for the defaults."""
    from math import log10, sqrt
    w = 1.0 - outlier_fraction
    k = log10(1.0 - p) / log10(1.0 - w ** n_points)
    std = sqrt(1.0 - w ** n_points) / (w ** n_points)
    return int(k + 3 * std)


# ---------------------------------------------------------------------------------------------------------------
# Non-linear refinement.  OpenGV is absent (parity unpinned); this restates its published algorithm:
# absolute_pose::optimize_nonlinear = Levenberg-Marquardt (Eigen's MINPACK port) over x = (t, Cayley(R)) with one
# residual per correspondence, 1 - f . reprojection — the same residual as score_bearing above.  scipy's
# least_squares(method="lm") is the original MINPACK driver.
# ---------------------------------------------------------------------------------------------------------------
def cayley_to_rot(c: np.ndarray) -> np.ndarray:
    c = np.asarray(c, np.float64)
    x, y, z = c
    s = 1.0 + x * x + y * y + z * z
    return np.array([[1 + x * x - y * y - z * z, 2 * (x * y - z), 2 * (x * z + y)],
                     [2 * (x * y + z), 1 - x * x + y * y - z * z, 2 * (y * z - x)],
                     [2 * (x * z - y), 2 * (y * z + x), 1 - x * x - y * y + z * z]]) / s


def rot_to_cayley(R: np.ndarray) -> np.ndarray:
    """Inverse of cayley_to_rot: [c]x = (R - I)(R + I)^-1."""
    R = np.asarray(R, np.float64)
    C = (R - np.eye(3)) @ np.linalg.inv(R + np.eye(3))
    return np.array([C[2, 1] - C[1, 2], C[0, 2] - C[2, 0], C[1, 0] - C[0, 1]]) * 0.5


def refine_pose_lm(p_ref, f_cur, pose0, cam=None, rig=None, mask=None, tol: float = 1e-15):
    """Refined [3,4] pose, and (initial cost, final cost) with cost = sum r^2."""
    from scipy.optimize import least_squares
    p_ref = np.asarray(p_ref, np.float64)
    f_cur = np.asarray(f_cur, np.float64)
    if mask is not None:
        m = np.asarray(mask, bool)
        p_ref, f_cur = p_ref[m], f_cur[m]
        cam = None if cam is None else np.asarray(cam)[m]
    pose0 = np.asarray(pose0, np.float64).reshape(3, 4)

    def unpack(x):
        M = np.zeros((3, 4))
        M[:, :3] = cayley_to_rot(x[3:])
        M[:, 3] = x[:3]
        return M

    def fun(x):
        return score_bearing(unpack(x), p_ref, f_cur, cam, rig)

    x0 = np.concatenate([pose0[:, 3], rot_to_cayley(pose0[:, :3])])
    sol = least_squares(fun, x0, method="lm", xtol=tol, ftol=tol, gtol=tol, max_nfev=2000)
    r0 = fun(x0)
    return unpack(sol.x), float(r0 @ r0), float(sol.fun @ sol.fun)
