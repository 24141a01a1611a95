"""Oracle for step 2 (brute-force Hamming matching + pixel gates). TEST INFRASTRUCTURE ONLY — see oracle/__init__.py."""
from __future__ import annotations

import cv2
import numpy as np

_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.uint16)


def hamming_matrix(q: np.ndarray, t: np.ndarray, block: int = 512) -> np.ndarray:
    """All-pairs Hamming distance of [Nq,32] x [Nt,32] uint8 descriptors -> [Nq,Nt] uint16 (NORM_HAMMING)."""
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    out = np.empty((q.shape[0], t.shape[0]), np.uint16)
    for i in range(0, q.shape[0], block):
        x = q[i:i + block, None, :] ^ t[None, :, :]
        out[i:i + block] = _POP8[x].sum(axis=-1, dtype=np.uint16)
    return out


def knn2(q: np.ndarray, t: np.ndarray):
    """Two nearest train rows per query in (distance, train index) order — cv2.BFMatcher.knnMatch(k=2) semantics
    (camera_models.py:421), ties to the lowest train index.  Missing neighbours are -1."""
    nq, nt = q.shape[0], t.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), -1, np.int32)
    if nq == 0 or nt == 0:
        return idx[:, 0], dist[:, 0], idx[:, 1], dist[:, 1]
    d = hamming_matrix(q, t).astype(np.int32)
    key = d * (1 << 22) + np.arange(nt, dtype=np.int32)[None, :]
    k = min(2, nt)
    part = np.sort(key, axis=1)[:, :k]
    idx[:, :k] = part & ((1 << 22) - 1)
    dist[:, :k] = part >> 22
    return idx[:, 0], dist[:, 0], idx[:, 1], dist[:, 1]


def bf_match_reference(q: np.ndarray, t: np.ndarray):
    """The reference's own call: cv2.BFMatcher(NORM_HAMMING).match (camera_models.py:402,442) -> (query, train, distance)."""
    m = cv2.BFMatcher(normType=cv2.NORM_HAMMING).match(queryDescriptors=q, trainDescriptors=t)
    return (np.array([x.queryIdx for x in m], np.int32), np.array([x.trainIdx for x in m], np.int32),
            np.array([x.distance for x in m], np.float64))


def bf_knn2_reference(q: np.ndarray, t: np.ndarray):
    m = cv2.BFMatcher(normType=cv2.NORM_HAMMING).knnMatch(queryDescriptors=q, trainDescriptors=t, k=2)
    return m


def filter_pixel_correspondences(pts_top, pts_bot, min_rectified_disparity, max_horizontal_diff):
    """common_cv.py:167-188, restated (float64 arithmetic on the given coordinates)."""
    pts_top = np.asarray(pts_top, np.float64)
    pts_bot = np.asarray(pts_bot, np.float64)
    if max_horizontal_diff > 0:
        ok = np.abs(pts_top[..., 0] - pts_bot[..., 0]) <= max_horizontal_diff
    else:
        ok = np.ones(pts_top.shape[:-1], bool)
    if min_rectified_disparity >= 0:
        ok = ok & (pts_top[..., 1] - pts_bot[..., 1] >= min_rectified_disparity)
    return ok


def knn2_flat_sorted(q, t):
    """FeatureMatcher.match with k_best = 2 on binary descriptors (camera_models.py:417-444): knnMatch(k=2) flattened
    query by query — BOTH neighbours kept, the Lowe test applies to SIFT only — then sorted(key=distance), which is
    stable: ties keep the (query index, neighbour rank) order.  Returns (query_idx, train_idx, distance)."""
    i0, d0, i1, d1 = knn2(q, t)
    qi = np.repeat(np.arange(q.shape[0], dtype=np.int32), 2)
    ti = np.stack([i0, i1], 1).reshape(-1)
    dd = np.stack([d0, d1], 1).reshape(-1)
    have = ti >= 0
    qi, ti, dd = qi[have], ti[have], dd[have]
    order = np.argsort(dd, kind="stable")
    return qi[order], ti[order], dd[order]


def radius_match_flat_sorted(q, t, max_distance):
    """FeatureMatcher.match with use_radius_match (camera_models.py:409-412): BFMatcher.radiusMatch keeps every train row with
    distance <= maxDistance per query, each query's list sorted by distance (ties in train-index order); the lists are
    flattened query by query and sorted(key=distance) — stable.  Returns (query_idx, train_idx, distance)."""
    d = hamming_matrix(q, t).astype(np.int32)
    qi, ti = np.nonzero(d <= max_distance)
    dd = d[qi, ti]
    within = np.lexsort((ti, dd, qi))
    qi, ti, dd = qi[within], ti[within], dd[within]
    order = np.argsort(dd, kind="stable")
    return qi[order].astype(np.int32), ti[order].astype(np.int32), dd[order]


def match_select(q, t, mode="nn", ratio=0.75, px_q=None, px_t=None, max_du=-1.0, min_dv=-1.0):
    """FeatureMatcher.match (camera_models.py:404-446) + the gate its callers apply (camera_models.py:3086,
    pose_est_tools.py:245-247), for one (query, train) segment.

    mode 'nn': 1-NN; 'ratio': knn(k=2) + Lowe test d0 < ratio*d1 (camera_models.py:423); 'cross': mutual nearest
    neighbours (BFMatcher crossCheck).  The result is ordered like sorted(matches, key=distance) — stable, i.e. by
    (distance, query index) — and then gated.  Returns (query_idx, train_idx, distance) int32 arrays."""
    i0, d0, i1, d1 = knn2(q, t)
    keep = i0 >= 0
    if mode == "ratio":
        keep &= (i1 >= 0) & (d0.astype(np.float64) < d1.astype(np.float64) * ratio)
    elif mode == "cross":
        r0, _, _, _ = knn2(t, q)
        qi = np.arange(q.shape[0])
        keep &= r0[np.clip(i0, 0, None)] == qi
    elif mode != "nn":
        raise ValueError(mode)
    qi = np.nonzero(keep)[0].astype(np.int32)
    ti = i0[qi]
    dd = d0[qi]
    order = np.argsort(dd, kind="stable")
    qi, ti, dd = qi[order], ti[order], dd[order]
    if px_q is not None:
        ok = filter_pixel_correspondences(px_t[ti], px_q[qi], min_dv, max_du)
        qi, ti, dd = qi[ok], ti[ok], dd[ok]
    return qi, ti, dd


# ---- float descriptors, L2 norm (the SIFT / SURF branch of FeatureMatcher, camera_models.py:397-399, 417-442) -----------
def l2_matrix_sq(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    """Squared L2 distances of integer-valued float descriptors as exact int64 (cv2 sums the squared float32 differences;
    for integer values in [0, 255] and <= 128 dimensions that sum is an exact integer below 2^24)."""
    qi, ti = np.asarray(q, np.float64).astype(np.int64), np.asarray(t, np.float64).astype(np.int64)
    return (qi * qi).sum(1)[:, None] + (ti * ti).sum(1)[None, :] - 2 * (qi @ ti.T)


def l2_knn2(q: np.ndarray, t: np.ndarray):
    """Two nearest train rows per query by (distance, train index); distances float32 = sqrt(float32(sum)) like cv2."""
    nq, nt = len(q), len(t)
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), -1.0, np.float32)
    if nq == 0 or nt == 0:
        return idx[:, 0], dist[:, 0], idx[:, 1], dist[:, 1]
    d2 = l2_matrix_sq(q, t)
    key = d2 * (1 << 22) + np.arange(nt, dtype=np.int64)[None, :]
    k = min(2, nt)
    part = np.sort(key, axis=1)[:, :k]
    idx[:, :k] = part & ((1 << 22) - 1)
    dist[:, :k] = np.sqrt((part >> 22).astype(np.float32))
    return idx[:, 0], dist[:, 0], idx[:, 1], dist[:, 1]


def l2_match(q, t, method="SIFT", k_best=1):
    """FeatureMatcher(method, "BF", k_best).match for float descriptors -> (query_idx, train_idx, distance float32)."""
    i0, d0, i1, d1 = l2_knn2(q, t)
    qi = np.arange(len(q), dtype=np.int32)
    if k_best == 2 and method.upper() == "SIFT":
        keep = (i1 >= 0) & (d0.astype(np.float64) < d1.astype(np.float64) * 0.75)
        qi, ti, dd = qi[keep], i0[keep], d0[keep]
    elif k_best == 2:
        qi = np.repeat(qi, 2)
        ti, dd = np.stack([i0, i1], 1).reshape(-1), np.stack([d0, d1], 1).reshape(-1)
        have = ti >= 0
        qi, ti, dd = qi[have], ti[have], dd[have]
    else:
        ti, dd = i0, d0
    order = np.argsort(dd, kind="stable")
    return qi[order], ti[order], dd[order]
