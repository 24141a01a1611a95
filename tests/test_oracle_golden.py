"""CPU: the oracle restatements against the golden vectors produced by the reference itself (oracle/gen_golden.py)."""
import numpy as np

from conftest import load_golden, sub
from oracle import geometry, hamming, ransac, remap


def test_remap_spec_matches_reference_panoramas():
    g = load_golden("remap.npz")
    for name in ("top", "bot"):
        mx, my, mask = g[f"map_x32_{name}"], g[f"map_y32_{name}"], g[f"mask_{name}"]
        # black mask background + black border (demo_vo_sos.py path: mask_RGB=(0,0,0))
        out = remap.remap_spec(g["img"], mx, my, border=(0, 0, 0), mask=mask, background=(0, 0, 0))
        assert np.array_equal(out, g[f"pano_{name}"])
        # same through OpenCV on the materialised masked image (what the reference literally does)
        ref = remap.remap_reference(remap.masked_image(g["img"], mask), mx, my)
        assert np.array_equal(ref, g[f"pano_{name}"])
        # coloured background (RGB (10,200,90) -> BGR) and coloured border (RGB (30,60,250) -> BGR)
        out = remap.remap_spec(g["img"], mx, my, border=(250, 60, 30), mask=mask, background=(90, 200, 10))
        assert np.array_equal(out, g[f"pano_colour_{name}"])
        # single-channel: the mirror mask itself remapped (panorama.py:534)
        out = remap.remap_spec(mask, mx, my, border=(0,))
        assert np.array_equal(out, g[f"pano_of_mask_{name}"])


def test_remap_spec_matches_cv2_on_adversarial_maps():
    rng = np.random.default_rng(0)
    H, W = 97, 131
    src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    R, C = 120, 203
    mx = rng.uniform(-3, W + 3, (R, C)).astype(np.float32)
    my = rng.uniform(-3, H + 3, (R, C)).astype(np.float32)
    mx[rng.random((R, C)) < 0.05] = np.nan
    my[rng.random((R, C)) < 0.05] = np.nan
    mx[0, :10] = [0, 0.5 / 32, 1.5 / 32, 2.5 / 32, W - 1, W - 1 + 0.5 / 32, W - 0.5, -1, -0.999, 1e20]
    my[0, :10] = [0, 0.5 / 32, 1.5 / 32, 2.5 / 32, H - 1, H - 1, H - 0.5, -1, -0.5, 3]
    ref = remap.remap_reference(src, mx, my, border=(7, 9, 11))
    assert np.array_equal(remap.remap_spec(src, mx, my, border=(7, 9, 11)), ref)


def test_lut_build_matches_reference_lut():
    g = load_golden("remap.npz")
    for name in ("top", "bot"):
        p = sub(g, f"gum_{name}_")
        pano = sub(g, f"pano_{name}_")
        lo, hi = g[f"elev_{name}"]
        u, v = geometry.lut_build(p, int(pano["rows"]), int(pano["cols"]), pano["cyl_height_max"], pano["cyl_height_min"], lo, hi)
        ru, rv = g[f"map_x64_sub_{name}"], g[f"map_y64_sub_{name}"]
        su, sv = u[::5, ::7], v[::5, ::7]
        assert np.array_equal(np.isnan(su), np.isnan(ru))
        ok = ~np.isnan(ru)
        # the reference evaluates parts of the chain in float32 (panorama.py:431,441); 1/64 px is the SURVEY bar
        assert np.max(np.abs(su[ok] - ru[ok])) < 1.0 / 64 and np.max(np.abs(sv[ok] - rv[ok])) < 1.0 / 64
        assert np.allclose(su[ok].astype("float32"), g[f"map_x32_{name}"][::5, ::7][ok], atol=1 / 64)


def test_hamming_oracle_matches_bfmatcher():
    g = load_golden("hamming.npz")
    q, t = g["q"], g["t"]
    qi, ti, dd = hamming.match_select(q, t, "nn")
    assert np.array_equal(qi, g["nn_q"]) and np.array_equal(ti, g["nn_t"]) and np.array_equal(dd, g["nn_d"].astype(np.int32))
    i0, d0, i1, d1 = hamming.knn2(q, t)
    assert np.array_equal(np.stack([i0, i1], 1), g["knn_t"])
    assert np.array_equal(np.stack([d0, d1], 1), g["knn_d"].astype(np.int32))
    # crossCheck: OpenCV returns mutual matches in query order
    cq, ct, cd = hamming.match_select(q, t, "cross")
    order = np.argsort(cq, kind="stable")
    assert np.array_equal(cq[order], g["cross_q"]) and np.array_equal(ct[order], g["cross_t"])
    assert np.array_equal(cd[order], g["cross_d"].astype(np.int32))
    # k_best = 2 on ORB descriptors: both neighbours, flattened and stably sorted (camera_models.py:417-444)
    kq, kt, kd = hamming.knn2_flat_sorted(q, t)
    assert len(kq) == 2 * len(q)
    assert np.array_equal(kq, g["k2_q"]) and np.array_equal(kt, g["k2_t"]) and np.array_equal(kd, g["k2_d"].astype(np.int32))
    # use_radius_match: radiusMatch flattened and sorted (camera_models.py:409-412), integer and fractional radius
    for tag in ("r40", "r70"):
        rq_, rt_, rd_ = hamming.radius_match_flat_sorted(q, t, float(g[f"{tag}_radius"]))
        assert len(rq_) > 50
        assert np.array_equal(rq_, g[f"{tag}_q"]) and np.array_equal(rt_, g[f"{tag}_t"]) and np.array_equal(rd_, g[f"{tag}_d"].astype(np.int32))
    # and against live OpenCV
    rq, rt, rd = hamming.bf_match_reference(q, t)
    assert np.array_equal(i0, rt) and np.array_equal(d0, rd.astype(np.int32))


def test_l2_oracle_matches_bfmatcher_on_float_descriptors():
    """The SIFT / SURF branch of FeatureMatcher (cv2.BFMatcher() = NORM_L2, camera_models.py:397-399, 417-442): the exact-integer
    restatement against goldens from the reference's class and against live OpenCV (distances bit-equal as float32)."""
    import cv2
    g = load_golden("l2.npz")
    for name, method, k, qk, tk in (("sift1", "SIFT", 1, "q", "t"), ("sift2", "SIFT", 2, "q", "t"), ("surf2", "SURF", 2, "q", "t"),
                                    ("surf1_64", "SURF", 1, "q64", "t64")):
        qi, ti, dd = hamming.l2_match(g[qk], g[tk], method, k)
        assert np.array_equal(qi, g[f"{name}_q"]) and np.array_equal(ti, g[f"{name}_t"]) and np.array_equal(dd, g[f"{name}_d"]), name
    assert 0 < len(g["sift2_q"]) < len(g["q"])          # the ratio test rejected some queries and kept others
    knn = cv2.BFMatcher().knnMatch(queryDescriptors=g["q"], trainDescriptors=g["t"], k=2)
    i0, d0, i1, d1 = hamming.l2_knn2(g["q"], g["t"])
    assert np.array_equal(np.array([[m.trainIdx for m in r] for r in knn]), np.stack([i0, i1], 1))
    assert np.array_equal(np.array([[m.distance for m in r] for r in knn], np.float32), np.stack([d0, d1], 1))


def test_stereo_and_temporal_matching_oracle():
    g = load_golden("matching_frames.npz")
    m_top, m_bot = [], []
    for b in range(int(g["n_buckets"])):
        q, t = g[f"b{b}_d_bot"], g[f"b{b}_d_top"]
        if len(q) == 0 or len(t) == 0:
            continue
        qi, ti, _ = hamming.match_select(q, t, "nn")
        m_top.append(g[f"b{b}_pt_top"][ti])
        m_bot.append(g[f"b{b}_pt_bot"][qi])
    m_top, m_bot = np.concatenate(m_top).astype(np.float64), np.concatenate(m_bot).astype(np.float64)
    ok = hamming.filter_pixel_correspondences(m_top, m_bot, 1, 2.5)
    assert np.array_equal(m_top[ok], g["stereo_m_top"][:, :2]) and np.array_equal(m_bot[ok], g["stereo_m_bot"][:, :2])
    assert 0 < ok.sum() < len(ok)
    qi, ti, _ = hamming.match_select(g["f2f_q"], g["f2f_t"], "nn", px_q=g["f2f_pq"][:, :2], px_t=g["f2f_pt"][:, :2],
                                     max_du=float(g["f2f_max_du"]), min_dv=-1)
    assert np.array_equal(qi, g["f2f_query_idx"]) and np.array_equal(ti, g["f2f_train_idx"])


def test_lifting_and_triangulation_oracle():
    g = load_golden("lifting.npz")
    pano = sub(g, "pano_top_")
    az, el = geometry.pano_pixel_to_angles(pano, g["pano_px"])
    assert np.array_equal(np.isnan(az), np.isnan(g["pano_az"])) and np.array_equal(np.isnan(el), np.isnan(g["pano_el"]))
    assert np.allclose(az, g["pano_az"], rtol=0, atol=1e-15, equal_nan=True)
    assert np.allclose(el, g["pano_el"], rtol=0, atol=1e-15, equal_nan=True)
    assert np.allclose(geometry.angles_to_sphere(az, el), g["pano_bearing"], atol=1e-15, equal_nan=True)
    xyz = geometry.triangulate_midpoint(g["tri_az1"], g["tri_el1"], g["tri_az2"], g["tri_el2"], g["tri_f1"], g["tri_f2"])
    ref = g["tri_xyz_homo"][:, :3]
    assert np.allclose(xyz, ref, rtol=1e-9, atol=1e-12)
    homo = np.hstack([xyz, np.ones((len(xyz), 1))])
    assert np.array_equal(geometry.range_filter(homo, 0.5, 7.0), g["tri_valid_homo"])
    assert np.array_equal(geometry.range_filter(xyz, 0.5, 7.0), g["tri_valid_xyz"])
    assert not np.array_equal(g["tri_valid_homo"], g["tri_valid_xyz"])  # the homogeneous-norm quirk is observable
    for tag in ("heik", "poly"):
        for name in ("top", "bot"):
            pre = f"{tag}_{name}_"
            p = sub(g, pre + "gum_")
            Ps, a, e = geometry.gum_lift(p, g[pre + "omni_uv"])
            assert np.allclose(Ps, g[pre + "sphere"], rtol=1e-12, atol=1e-14, equal_nan=True)
            assert np.allclose(a, g[pre + "omni_az"], atol=1e-13, equal_nan=True)
            assert np.allclose(e, g[pre + "omni_el"], atol=1e-13, equal_nan=True)
            u, v = geometry.gum_project(p, g[pre + "proj_pts"])
            assert np.allclose(u, g[pre + "proj_u"], rtol=1e-12) and np.allclose(v, g[pre + "proj_v"], rtol=1e-12)


def test_rgbd_oracle():
    g = load_golden("rgbd.npz")
    for tag in ("z", "radial"):
        cam = dict(zip(geometry.RGBD_FIELDS, g[f"{tag}_cam"]))
        z = geometry.rgbd_depth_to_z(cam, g["depth"])
        assert np.allclose(z, g[f"{tag}_depth_z"], rtol=1e-6)  # reference mixes float32/float64 here
        xyz, bearing, valid = geometry.rgbd_backproject(cam, g["depth"], g["u"], g["v"], 0.8, 7.0)
        assert np.allclose(xyz, g[f"{tag}_xyz"], rtol=1e-6, equal_nan=True)
        assert np.allclose(bearing, g[f"{tag}_bearing"], rtol=1e-6, atol=1e-9, equal_nan=True)
        assert np.array_equal(np.isnan(xyz[:, 2]), np.isnan(g[f"{tag}_xyz"][:, 2]))
        assert 0 < valid.sum() < len(valid)


def test_arun_and_score_oracle():
    g = load_golden("arun.npz")
    for k in (3, 4, 100):
        v0, v1, M = g[f"k{k}_v0"], g[f"k{k}_v1"], g[f"k{k}_M"]
        got = ransac.arun_batch(v0, v1)
        assert np.allclose(got, M, rtol=0, atol=1e-9)
        one = ransac.superimposition(v0[0].T, v1[0].T)[:3]
        assert np.allclose(one, M[0], atol=1e-12)
        assert np.allclose(np.linalg.det(got[:, :, :3]), 1.0)
    sc = ransac.score_bearing(g["score_M"], g["score_p_ref"], g["score_f"])
    assert np.allclose(sc, g["score_values"], rtol=0, atol=1e-14)
    assert np.array_equal(np.nonzero(sc < float(g["score_thr"]))[0], g["score_inliers"])
    assert ransac.num_iterations() == int(g["ransac_iters_default"]) == 210


def test_dense_triangulation_oracle():
    """the oracle's validity chain + geometry reproduce the reference's dense point list exactly."""
    g = load_golden("dense.npz")
    pt = {k[9:]: float(g[k]) for k in g if k.startswith("pano_top_")}
    pb = {k[9:]: float(g[k]) for k in g if k.startswith("pano_bot_")}
    for tag in ("all", "roi"):
        a = g[tag + "_args"]
        roi = None if a[2] < 0 else (int(a[2]), int(a[3]))
        xyz, valid = geometry.dense_triangulate(pt, pb, g["disparity"], g["f1"], g["f2"], a[0], a[1],
                                                float(g["lowest_reference_row"]), roi)
        vt = valid.T
        uu, vv = np.nonzero(vt)
        assert np.array_equal(np.stack([uu, vv], 1), g[tag + "_top_px"])
        assert np.allclose(xyz.transpose(1, 0, 2)[vt], g[tag + "_xyz"], rtol=1e-12, atol=1e-12)
        assert np.array_equal(g[tag + "_bot_px"][:, 1], vv - g["disparity"][vv, uu])


def test_orb_description_oracle_matches_cv2():
    """the restated descriptor model (blur inside ORB, test-point table, float32 rotation, border filter)
    equals cv2.ORB.compute bit for bit on random keypoints and angles."""
    import cv2
    from oracle import orb
    rng = np.random.default_rng(4)
    gray = cv2.GaussianBlur(rng.integers(0, 256, (200, 260), dtype=np.uint8), (0, 0), 1.5)
    pts = np.stack([rng.uniform(0, 260, 400), rng.uniform(0, 200, 400)], 1).astype(np.float32)
    ang = rng.uniform(0, 360, 400).astype(np.float32)
    ang[::4] = -1.0
    kps = [cv2.KeyPoint(float(p[0]), float(p[1]), 1.0, float(a), 1.0, 0, i) for i, (p, a) in enumerate(zip(pts, ang))]
    k2, want = cv2.ORB_create(nfeatures=10).compute(gray, kps)
    got, keep = orb.describe(gray, pts, ang)
    kept = np.array([k.class_id for k in k2])
    assert np.array_equal(np.sort(kept), np.flatnonzero(keep))
    assert np.array_equal(got[kept], want)
    assert orb.pattern().shape == (256, 4) and np.abs(orb.pattern()).max() == 13


def test_p3p_oracle_solves_the_published_problem():
    """oracle/p3p.py has no OpenGV output to pin against (parity unpinned): check that every returned depth triple satisfies
    the three distance constraints, that the planted depths are among them, and that RANSAC on planted data returns the motion
    — central and with the two-viewpoint rig."""
    from oracle import p3p
    from oracle.ransac import cayley_to_rot
    rng = np.random.default_rng(3)
    for _ in range(50):
        d = rng.normal(0, 1, (3, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        s_true = rng.uniform(0.5, 7, 3)
        X = d * s_true[:, None]
        R, t = cayley_to_rot(rng.normal(0, 0.3, 3)), rng.normal(0, 1, 3)
        P = X @ R.T + t
        sols = p3p.grunert_depths(d, P)
        assert any(np.allclose(s, s_true, rtol=1e-7) for s in sols)
        for s in sols:
            Y = d * np.array(s)[:, None]
            for i, j in ((0, 1), (0, 2), (1, 2)):
                assert abs(np.linalg.norm(Y[i] - Y[j]) - np.linalg.norm(P[i] - P[j])) < 1e-7
    rig = np.stack([np.hstack([np.eye(3), [[0.0], [0.0], [0.06]]]), np.hstack([cayley_to_rot([0.01, 0.02, 0.0]), [[0.01], [0.0], [-0.07]]])])
    for use_rig in (False, True):
        n = 200
        R, t = cayley_to_rot(rng.normal(0, 0.05, 3)), rng.normal(0, 0.1, 3)
        pb = rng.normal(0, 1, (n, 3)); pb = pb / np.linalg.norm(pb, axis=1, keepdims=True) * rng.uniform(0.5, 7, (n, 1))
        cam = rng.integers(0, 2, n) if use_rig else None
        x = np.einsum("nji,nj->ni", rig[cam][:, :, :3], pb - rig[cam][:, :, 3]) if use_rig else pb
        f = x / np.linalg.norm(x, axis=1, keepdims=True)
        bad = rng.random(n) < 0.3
        f[bad] = -f[bad]
        hyp = rng.integers(0, 2 ** 32, (100, 4), dtype=np.uint64).astype(np.uint32)
        M, h, c, inl, counts = p3p.ransac_p3p(pb @ R.T + t, f, cam, rig if use_rig else None, hyp, 1 - np.cos(np.radians(1.0)))
        assert c == (~bad).sum() and np.array_equal(inl, ~bad)
        assert np.allclose(M, np.hstack([R, t[:, None]]), atol=1e-9)
