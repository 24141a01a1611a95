"""GPU: the tensor-core engine of the bearing score (csrc/score_mma.cuh).

Its decisions are only trusted outside a guard band, so two things are pinned here: (1) what the tensor cores accumulate for
(s, n2) = (f . x, |x|^2) stays within a QUARTER of the band the kernel uses, measured against float64 on every pair through the
probe entry point, and (2) the counts equal the float64 oracle's on shapes that exercise ragged tiles, mixed-camera tiles,
unsorted camera lists, non-finite inputs and many pairs near the threshold."""
import numpy as np
import pytest
import torch

from oracle import ransac
from test_gpu_ransac import as_i32, dev, hyp_list, make_problem, random_rotation

pytestmark = pytest.mark.gpu

THR = 1.0 - np.cos(np.deg2rad(5.0))
BAND_REL = 13.0 * 2.0 ** -20 + 5e-7      # csrc/score_mma.cuh
B2_SHIFT = 1.52
BAND_EUCLID = 5e-6                       # csrc/score_mma.cuh


def rig_pair(rng):
    Rc = random_rotation(rng, 10.0)
    return np.stack([np.hstack([np.eye(3), [[0.0], [0.0], [0.12]]]), np.hstack([Rc, [[0.01], [-0.02], [0.0]]])])


def pack(probs, cap):
    B = len(probs)
    p_ref = np.zeros((B, cap, 3), np.float32); p_cur = np.zeros_like(p_ref); f_cur = np.zeros_like(p_ref)
    cam = np.zeros((B, cap), np.uint8)
    ns = np.zeros(B, np.int32)
    for b, (a, c, f, cm) in enumerate(probs):
        ns[b] = len(a)
        p_ref[b, :len(a)], p_cur[b, :len(a)], f_cur[b, :len(a)], cam[b, :len(a)] = a, c, f, cm
    return p_ref, p_cur, f_cur, cam, ns


@pytest.mark.parametrize("with_rig", [False, True])
def test_accumulators_within_a_quarter_of_the_band(ctx, with_rig):
    rng = np.random.default_rng(41 + with_rig)
    rig = rig_pair(rng) if with_rig else None
    n, H, B = 700, 300, 2
    probs = [make_problem(rng, n - 33 * b, rig=rig) for b in range(B)]
    p_ref, p_cur, f_cur, cam, ns = pack(probs, n)
    hyp = hyp_list(rng, H)
    counts, sn = ctx.ransac_score_probe(dev(p_ref), dev(p_cur), dev(f_cur), dev(ns), as_i32(hyp), THR,
                                        cam=dev(cam) if with_rig else None, rig=rig, n_cams=2 if with_rig else 0)
    counts, sn = counts.cpu().numpy(), sn.cpu().numpy().astype(np.float64)
    c2 = float(np.float32((1.0 - THR) ** 2))
    worst = 0.0
    for b, (a, c, f, cm) in enumerate(probs):
        a64, c64, f64 = a.astype(np.float64), c.astype(np.float64), f.astype(np.float64)
        o = ransac.ransac_p3d(a, c, hyp, "bearing", THR, f_cur=f, cam=cm if with_rig else None, rig=rig)
        assert np.array_equal(np.where(counts[b] < 0, -1, counts[b]), o["counts"])
        rows = ransac.sample_rows(hyp, len(a))
        good = np.nonzero(o["counts"] >= 0)[0]
        poses = ransac.arun_batch(c64[rows[good]], a64[rows[good]])
        r = np.asarray(rig, np.float64).reshape(-1, 3, 4) if with_rig else np.hstack([np.eye(3), np.zeros((3, 1))])[None]
        ci = cm.astype(np.int64) if with_rig else np.zeros(len(a), np.int64)
        for h, M in zip(good, poses):
            R, t = M[:, :3], M[:, 3]
            body = (a64 - t) @ R
            x = np.einsum("nji,nj->ni", r[ci, :, :3], body - r[ci, :, 3])
            s, n2 = np.sum(f64 * x, axis=1), np.sum(x * x, axis=1)
            bvec = np.einsum("nji,nj->ni", r[ci, :, :3], (-t @ R)[None, :] - r[ci, :, 3])
            scale = np.sum(a64 * a64, axis=1) + np.sum(bvec * bvec, axis=1)
            # the n2 accumulator carries the hypothesis' band constant: N' = n2 + B2_SHIFT max_cam |b|^2 (score_mma.cuh)
            ball = np.einsum("cji,j->ci", r[:, :, :3], -t @ R) - np.einsum("cji,cj->ci", r[:, :, :3], r[:, :, 3])
            shift = B2_SHIFT * float(np.max(np.sum(ball * ball, axis=1)))
            got_s, got_n = sn[b, h, :len(a), 0], sn[b, h, :len(a), 1] - shift
            D_true = s * np.abs(s) - c2 * n2
            D_got = got_s * np.abs(got_s) - c2 * got_n
            worst = max(worst, float(np.max(np.abs(D_got - D_true) / scale)))
        # padding correspondences are certain outliers
        pad = sn[b, good[0], len(a):((len(a) + 127) // 128) * 128]
        assert np.all(pad[:, 1] > 1e29) and np.all(pad[:, 0] == 0.0)
    print(f"worst |err D| / (|p|^2 + |b|^2) = {worst:.3e}  (band {BAND_REL:.3e})")
    assert worst < BAND_REL / 4.0, worst


def test_euclid_accumulators_within_a_quarter_of_the_band(ctx):
    """Euclidean score: r^2 = s' + n2 from the two accumulators against float64, scaled by |p|^2 + |q|^2 + |b|^2."""
    rng = np.random.default_rng(77)
    n, H, B = 700, 300, 2
    thr = 0.05
    probs = [make_problem(rng, n - 33 * b) for b in range(B)]
    p_ref, p_cur, f_cur, cam, ns = pack(probs, n)
    hyp = hyp_list(rng, H)
    counts, sn = ctx.ransac_score_probe(dev(p_ref), dev(p_cur), None, dev(ns), as_i32(hyp), thr, score_mode=0)
    counts, sn = counts.cpu().numpy(), sn.cpu().numpy().astype(np.float64)
    worst = 0.0
    for b, (a, c, f, cm) in enumerate(probs):
        a64, c64 = a.astype(np.float64), c.astype(np.float64)
        o = ransac.ransac_p3d(a, c, hyp, "euclid", thr)
        assert np.array_equal(np.where(counts[b] < 0, -1, counts[b]), o["counts"])
        rows = ransac.sample_rows(hyp, len(a))
        good = np.nonzero(o["counts"] >= 0)[0]
        poses = ransac.arun_batch(c64[rows[good]], a64[rows[good]])
        for h, M in zip(good, poses):
            R, t = M[:, :3], M[:, 3]
            x = (a64 - t) @ R                      # R^T (p_ref - t): the reference point in the current frame
            r2 = np.sum((x - c64) ** 2, axis=1)
            bvec = -t @ R
            scale = np.sum(a64 * a64, axis=1) + np.sum(c64 * c64, axis=1) + float(bvec @ bvec)
            got = sn[b, h, :len(a), 0] + sn[b, h, :len(a), 1]
            worst = max(worst, float(np.max(np.abs(got - r2) / scale)))
    print(f"euclid: worst |err r^2| / (|p|^2 + |q|^2 + |b|^2) = {worst:.3e}  (band {BAND_EUCLID:.3e})")
    assert worst < BAND_EUCLID / 4.0, worst


def counts_equal_oracle(ctx, probs, cap, hyp, rig, cams):
    p_ref, p_cur, f_cur, cam, ns = pack(probs, cap)
    B, H = len(probs), len(hyp)
    all_counts = torch.empty((B, H), dtype=torch.int32, device="cuda")
    pose, best_hyp, best_count, mask, key = ctx.ransac_p3d(dev(p_ref), dev(p_cur), dev(ns), as_i32(hyp), 1, THR, f_cur=dev(f_cur),
                                                           cam=dev(cam) if cams else None, rig=rig, n_cams=2 if cams else 0,
                                                           all_counts=all_counts)
    all_counts = all_counts.cpu().numpy()
    for b, (a, c, f, cm) in enumerate(probs):
        o = ransac.ransac_p3d(a, c, hyp, "bearing", THR, f_cur=f, cam=cm if cams else None, rig=rig)
        assert np.array_equal(np.where(all_counts[b] < 0, -1, all_counts[b]), o["counts"]), b
        assert int(best_hyp[b]) == o["best_hyp"] and int(best_count[b]) == o["best_count"]
        assert np.array_equal(mask[b, :len(a)].cpu().numpy().astype(bool), o["mask"])


@pytest.mark.parametrize("n,H", [(129, 33), (128, 128), (1000, 257), (5000, 1024)])
def test_ragged_tiles_and_unsorted_cameras(ctx, n, H):
    """Camera indices in random order: every correspondence tile mixes the two cameras (both MMAs into one accumulator)."""
    rng = np.random.default_rng(n + H)
    rig = rig_pair(rng)
    probs = []
    for b in range(2):
        a, c, f, cm = make_problem(rng, n - 5 * b, rig=rig)
        perm = rng.permutation(len(a))
        probs.append((a[perm], c[perm], f[perm], cm[perm]))
    counts_equal_oracle(ctx, probs, n, hyp_list(rng, H), rig, True)


def test_many_pairs_near_the_threshold(ctx):
    """Bearings rotated so that a large share of the residuals of the true motion sits within 1e-6 of the threshold: the
    deferred path decides them; counts still equal the float64 oracle's wherever its own margin allows a decision."""
    rng = np.random.default_rng(7)
    n, H = 3000, 256
    a, c, f, cm = make_problem(rng, n, inlier_frac=0.6, sigma=0.0)
    # tilt every other inlier bearing by 5 degrees +- 1e-5 about a random axis perpendicular to it
    f64 = f.astype(np.float64)
    for j in range(0, n, 2):
        ax = np.cross(f64[j], rng.normal(size=3)); ax /= np.linalg.norm(ax)
        ang = np.deg2rad(5.0) + rng.uniform(-1e-5, 1e-5)
        f64[j] = f64[j] * np.cos(ang) + np.cross(ax, f64[j]) * np.sin(ang)
    f = f64.astype(np.float32)
    hyp = hyp_list(rng, H)
    p_ref, p_cur, f_cur, cam, ns = pack([(a, c, f, cm)], n)
    all_counts = torch.empty((1, H), dtype=torch.int32, device="cuda")
    ctx.ransac_p3d(dev(p_ref), dev(p_cur), dev(ns), as_i32(hyp), 1, THR, f_cur=dev(f_cur), all_counts=all_counts)
    got = all_counts.cpu().numpy()[0]
    o = ransac.ransac_p3d(a, c, hyp, "bearing", THR, f_cur=f)
    assert np.array_equal(np.where(got < 0, -1, got), o["counts"])


def test_non_finite_correspondences_never_count(ctx):
    rng = np.random.default_rng(3)
    n, H = 600, 64
    a, c, f, cm = make_problem(rng, n)
    a2 = a.copy()
    a2[5] = np.nan
    a2[200, 1] = np.inf
    hyp = hyp_list(rng, H)
    # keep the poisoned rows out of the minimal samples
    rows = ransac.sample_rows(hyp, n)
    hyp = hyp[~np.isin(rows, [5, 200]).any(axis=1)]
    H = len(hyp)
    p_ref, p_cur, f_cur, cam, ns = pack([(a2, c, f, cm)], n)
    all_counts = torch.empty((1, H), dtype=torch.int32, device="cuda")
    ctx.ransac_p3d(dev(p_ref), dev(p_cur), dev(ns), as_i32(hyp), 1, THR, f_cur=dev(f_cur), all_counts=all_counts)
    got = all_counts.cpu().numpy()[0]
    with np.errstate(invalid="ignore"):
        o = ransac.ransac_p3d(a2, c, hyp, "bearing", THR, f_cur=f)
    assert np.array_equal(np.where(got < 0, -1, got), o["counts"])
