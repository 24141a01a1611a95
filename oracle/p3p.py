"""TEST INFRASTRUCTURE ONLY (never imported by the product).

Bearing-only RANSAC hypotheses: what pyopengv.absolute_pose_ransac("KNEIP"/"GAO") and
pyopengv.absolute_pose_noncentral_ransac (GP3P) hypothesise at the reference's call sites (pose_est_tools.py:785, 915), where
only the bearings of the current frame and the 3D points of the reference frame are passed.  OpenGV is absent from
/root/reference, so this restates the PUBLISHED problem, not OpenGV's code — parity with OpenGV is unpinned:

  * sample 4 correspondences (OpenGV's AbsolutePoseSacProblem draws 3 + 1 for the 3-point solvers: the 4th picks among
    the solutions);
  * find depths l_i > 0 along the rays  X_i = o_i + l_i d_i  (o_i, d_i: ray of correspondence i in the body frame, from the
    rig [Rc|tc]) with |X_i - X_j| = |P_i - P_j|: Grunert's three-point problem (Haralick et al., IJCV 1994) for a common
    origin - a quartic in v = s3/s1 - and, for a non-central rig, Newton's method on the three distance equations started at
    every central solution;
  * pose from the two congruent triangles; the solution whose reprojection of the 4th point has the smallest bearing
    residual 1 - f . x/|x| is the hypothesis.

This file solves the quartic with numpy.roots (companion matrix), the kernel with Ferrari's closed form: independent routes.
"""
from __future__ import annotations

import numpy as np

NEWTON_ITERS = 8


def sample_rows4(hyp_row, n):
    """4 rows of hypothesis h: floor(u * n / 2^32) (same rule as oracle.ransac.sample_rows)."""
    return [int((int(u) * int(n)) >> 32) for u in hyp_row]


def _unit(v):
    return v / np.linalg.norm(v)


def grunert_depths(d, P):
    """All (s1, s2, s3) > 0 with |s_i d_i - s_j d_j| = |P_i - P_j| for unit directions d [3,3] from one origin."""
    a2 = float(np.sum((P[1] - P[2]) ** 2))
    b2 = float(np.sum((P[0] - P[2]) ** 2))
    c2 = float(np.sum((P[0] - P[1]) ** 2))
    ca, cb, cg = float(d[1] @ d[2]), float(d[0] @ d[2]), float(d[0] @ d[1])
    ra, rc = a2 / b2, c2 / b2
    # with u = s2/s1, v = s3/s1:  u = N(v) / (2 D(v)),  N^2 - 4 cg N D + 4 Q D^2 = 0
    N = np.array([1 - ra + rc, 2 * (ra - rc) * cb, rc - ra - 1])          # highest power first
    D = np.array([ca, -cg])
    Q = np.array([-rc, 2 * rc * cb, 1 - rc])
    quartic = np.polyadd(np.polysub(np.polymul(N, N), 4 * cg * np.polymul(N, D)), 4 * np.polymul(Q, np.polymul(D, D)))
    out = []
    if not np.all(np.isfinite(quartic)) or abs(quartic[0]) < 1e-300:
        return out
    scale = np.max(np.abs(quartic))
    for root in np.roots(quartic):
        if abs(root.imag) > 1e-7 * max(1.0, abs(root.real)):
            continue
        v = float(root.real)
        for _ in range(2):                                                 # polish on the quartic itself
            f = np.polyval(quartic, v)
            df = np.polyval(np.polyder(quartic), v)
            if df != 0:
                v -= f / df
        if not (v > 0) or abs(np.polyval(quartic, v)) > 1e-6 * scale * max(1.0, v ** 4):
            continue
        den = 2 * np.polyval(D, v)
        if abs(den) < 1e-12:
            continue
        u = np.polyval(N, v) / den
        w = 1 + v * v - 2 * v * cb
        if not (u > 0) or not (w > 0):
            continue
        s1 = np.sqrt(b2 / w)
        out.append((s1, u * s1, v * s1))
    return out


def newton_noncentral(o, d, P, lam0):
    """Depths along rays with their own origins o [3,3]; None if it does not converge to positive depths."""
    a2 = float(np.sum((P[1] - P[2]) ** 2))
    b2 = float(np.sum((P[0] - P[2]) ** 2))
    c2 = float(np.sum((P[0] - P[1]) ** 2))
    lam = np.array(lam0, float)
    for _ in range(NEWTON_ITERS):
        X = o + lam[:, None] * d
        e12, e13, e23 = X[0] - X[1], X[0] - X[2], X[1] - X[2]
        g = np.array([e12 @ e12 - c2, e13 @ e13 - b2, e23 @ e23 - a2])
        J = 2 * np.array([[e12 @ d[0], -(e12 @ d[1]), 0.0],
                          [e13 @ d[0], 0.0, -(e13 @ d[2])],
                          [0.0, e23 @ d[1], -(e23 @ d[2])]])
        det = np.linalg.det(J)
        if not np.isfinite(det) or abs(det) < 1e-300:
            return None
        lam = lam - np.linalg.solve(J, g)
    X = o + lam[:, None] * d
    g = np.array([np.sum((X[0] - X[1]) ** 2) - c2, np.sum((X[0] - X[2]) ** 2) - b2, np.sum((X[1] - X[2]) ** 2) - a2])
    if not np.all(lam > 0) or not np.max(np.abs(g)) < 1e-9 * max(a2, b2, c2):
        return None
    return lam


def pose_from_triangles(X, P):
    """[R|t] with P_i = R X_i + t for two congruent triangles (orthonormal frames on the first two edges)."""
    def frame(T):
        e1 = _unit(T[1] - T[0])
        e3 = _unit(np.cross(e1, T[2] - T[0]))
        return np.stack([e1, np.cross(e3, e1), e3], axis=1)
    R = frame(P) @ frame(X).T
    t = P.mean(0) - R @ X.mean(0)
    return np.hstack([R, t[:, None]])


def bearing_residual(M, p, f, Rc, tc):
    x = Rc.T @ (M[:, :3].T @ (p - M[:, 3]) - tc)
    return 1.0 - float(f @ x) / float(np.linalg.norm(x))


def hypothesis(p_ref, f_cur, cam, rig, rows):
    """Pose [3,4] (p_ref ~ R p_body + t) of one 4-row sample, or None.  rig [n_cams,3,4] = [Rc|tc] or None (central)."""
    if len(set(rows)) < 4:
        return None
    P = np.asarray(p_ref, np.float64)[rows]
    F = np.asarray(f_cur, np.float64)[rows]
    if rig is None:
        Rcs, tcs = [np.eye(3)] * 4, [np.zeros(3)] * 4
    else:
        rig = np.asarray(rig, np.float64).reshape(-1, 3, 4)
        cc = [0] * 4 if cam is None else [int(cam[r]) for r in rows]
        Rcs, tcs = [rig[c][:, :3] for c in cc], [rig[c][:, 3] for c in cc]
    e1, e2 = P[1] - P[0], P[2] - P[0]
    cr = np.cross(e1, e2)
    if not (cr @ cr > 1e-12 * (e1 @ e1) * (e2 @ e2)):
        return None                                                       # collinear sample
    d = np.stack([Rcs[i] @ F[i] for i in range(3)])
    o = np.stack([tcs[i] for i in range(3)])
    o_mean = o.mean(0)
    best, best_res = None, np.inf
    for s in grunert_depths(d, P[:3]):
        lam0 = [s[i] - d[i] @ (o[i] - o_mean) for i in range(3)]
        lam = newton_noncentral(o, d, P[:3], lam0)
        if lam is None:
            continue
        M = pose_from_triangles(o + lam[:, None] * d, P[:3])
        res = bearing_residual(M, P[3], F[3], Rcs[3], tcs[3])
        if res < best_res:
            best, best_res = M, res
    return best


def ransac_p3p(p_ref, f_cur, cam, rig, hyp, threshold):
    """-> (best pose, best hypothesis, best count, inlier mask, all counts): first maximum wins (OpenGV's strict '>')."""
    from .ransac import score_bearing
    p_ref = np.asarray(p_ref, np.float64)
    f_cur = np.asarray(f_cur, np.float64)
    n = len(p_ref)
    counts = np.full(len(hyp), -(1 << 30), np.int64)
    best = (None, -1, -1, np.zeros(n, bool))
    rig_arr = None if rig is None else np.asarray(rig, np.float64).reshape(-1, 3, 4)
    for h, row in enumerate(hyp):
        if n < 4:
            break
        M = hypothesis(p_ref, f_cur, cam, rig_arr, sample_rows4(row, n))
        if M is None:
            continue
        r = score_bearing(M, p_ref, f_cur, cam, rig_arr)
        inl = r < threshold
        counts[h] = int(inl.sum())
        if counts[h] > best[2]:
            best = (M, h, int(counts[h]), inl)
    return best + (counts,)
