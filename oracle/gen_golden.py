"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference under the shims of
oracle/_ref_shims.py) on seeded synthetic inputs.  Run it in the build container only:

    python -m oracle.gen_golden

The fixtures pin oracle/ (tests/test_oracle_golden.py) and, through it, the CUDA path.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import io
import contextlib
import os
import sys
import warnings

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

GUM_FIELDS = ("xi1", "xi2", "xi3", "k1", "k2", "k3", "gamma1", "gamma2", "alpha_c", "u_center", "v_center",
              "l1", "l2", "l3", "p1", "p2")


def gum_dict(g):
    p = g.precalib_params
    d = {k: float(getattr(p, k)) for k in GUM_FIELDS}
    d["plane_k"] = float(g.plane_k)
    d["use_distortion"] = float(bool(p.use_distortion))
    return d


def build_gums(W=320, H=240, pano_width=200, seed=0, l1=0.0):
    """Recipe of SURVEY §8c: a synthetic GUMS through the reference's own constructors."""
    from omnistereo.gum import GUM, GUMStereo
    rng = np.random.default_rng(seed)
    c = np.array([W / 2 - 0.5 + 1.3, H / 2 - 0.5 - 0.7])

    def mk(z_axis, xi3, gamma, F):
        g = GUM(precalib_filename="/nonexistent", new_method=True, z_axis=z_axis, image_size_pixels=(W, H), center_uv_point=c)
        g.precalib_params.set_gum_params(xi1=0.01, xi2=-0.012, xi3=xi3)
        g.precalib_params.set_generalized_cam_params(gamma1=gamma, gamma2=gamma * 1.01, alpha_c=0.002, u_center=c[0], v_center=c[1])
        g.precalib_params.k1, g.precalib_params.k2, g.precalib_params.k3 = 0.02, -0.003, 0.0005
        g.precalib_params.p1, g.precalib_params.p2 = 0.001, -0.0007
        g.precalib_params.l1, g.precalib_params.l2, g.precalib_params.l3 = l1, -l1 / 7.0, l1 / 31.0
        g.set_model_params()
        g.units = "m"
        g.set_pose(np.array(F), np.eye(3))
        return g

    top = mk(1.0, 0.85, 0.2 * W, [0.0, 0.0, 0.12])
    bot = mk(-1.0, -0.9, 0.17 * W, [0.0, 0.0, 0.0])
    gs = GUMStereo(top, bot, center_point_top=c, center_point_top_inner=c, center_point_top_outer=c,
                   center_point_bottom=c, center_point_bottom_inner=c, center_point_bottom_outer=c,
                   radius_top_outer=0.48 * H, radius_top_inner=0.30 * H, radius_bottom_outer=0.28 * H,
                   radius_bottom_inner=0.08 * H)
    img = rng.integers(0, 256, (H // 8, W // 8, 3), dtype=np.uint8).repeat(8, axis=0).repeat(8, axis=1)
    img = (img.astype(np.int16) + rng.integers(-20, 21, img.shape)).clip(0, 255).astype(np.uint8)
    gs.set_current_omni_image(img, pano_width_in_pixels=pano_width, generate_panoramas=True, view=False, apply_mask=True,
                              mask_RGB=(0, 0, 0))
    return gs, img


def pano_dict(p):
    return dict(cols=float(p.cols), rows=float(p.rows), pixel_size=float(p.pixel_size), cyl_height_max=float(p.cyl_height_max),
                cyl_height_min=float(p.z_height_min), cyl_circumference=float(p.cyl_circumference), cyl_radius=float(p.cyl_radius))


def flat(prefix, d):
    return {f"{prefix}{k}": np.float64(v) for k, v in d.items()}


def gen_remap():
    gs, img = build_gums()
    out = dict(img=img)
    for name, m in (("top", gs.top_model), ("bot", gs.bot_model)):
        p = m.panorama
        out[f"mask_{name}"] = m.mask
        out[f"map_x32_{name}"] = p.world2cam_LUT_map_x.astype("float32")
        out[f"map_y32_{name}"] = p.world2cam_LUT_map_y.astype("float32")
        # sparse float64 samples pin the LUT generator (F3) without shipping the whole table
        out[f"map_x64_sub_{name}"] = p.world2cam_LUT_map_x[::5, ::7]
        out[f"map_y64_sub_{name}"] = p.world2cam_LUT_map_y[::5, ::7]
        out[f"pano_{name}"] = p.panoramic_img.copy()  # later calls reuse this buffer as cv2.remap dst (panorama.py:295)
        out.update(flat(f"gum_{name}_", gum_dict(m)))
        out.update(flat(f"pano_{name}_", pano_dict(p)))
        out[f"elev_{name}"] = np.array([m.lowest_elevation_angle, m.highest_elevation_angle])
        # single-channel remap of the mirror mask, as generate_azimuthal_masks does (panorama.py:534)
        out[f"pano_of_mask_{name}"] = p.get_panoramic_image(input_omni_img=m.mask, set_own=False, border_RGB_color=(0, 0, 0)).copy()
        # a non-black border / background variant
        masked = gs.get_fully_masked_images(omni_img=img, view=False, color_RGB=(10, 200, 90))
        out[f"pano_colour_{name}"] = p.get_panoramic_image(masked[0 if name == "top" else 1], set_own=False,
                                                            border_RGB_color=(30, 60, 250)).copy()
    np.savez_compressed(os.path.join(OUT, "remap.npz"), **out)


def make_descriptors(rng, n_land, n_q, n_t, flip=0.08, n_ties=8):
    land = rng.integers(0, 256, (n_land, 32), dtype=np.uint8)

    def view(n):
        ids = rng.permutation(n_land)[:n]
        d = land[ids].copy()
        noise = (rng.random((n, 256)) < flip)
        d ^= np.packbits(noise, axis=1)
        return d, ids

    q, qi = view(n_q)
    t, ti = view(n_t)
    # distractors: a quarter of each view is replaced by unrelated descriptors
    q[: n_q // 4] = rng.integers(0, 256, (n_q // 4, 32), dtype=np.uint8)
    t[: n_t // 4] = rng.integers(0, 256, (n_t // 4, 32), dtype=np.uint8)
    # planted exact ties: duplicate train rows (lowest index must win), and duplicate query rows
    for k in range(n_ties):
        a, b = rng.integers(n_t // 4, n_t, 2)
        t[b] = t[a]
        a, b = rng.integers(n_q // 4, n_q, 2)
        q[b] = q[a]
    return q, t


def gen_hamming():
    import cv2
    from omnistereo.camera_models import FeatureMatcher
    rng = np.random.default_rng(1)
    out = {}
    q, t = make_descriptors(rng, 400, 300, 280)
    out["q"], out["t"] = q, t
    fm = FeatureMatcher("ORB", "BF", 1, percentage_good_matches=1.0)
    m = fm.match(query_descriptors=q, train_descriptors=t)
    out["nn_q"] = np.array([x.queryIdx for x in m], np.int32)
    out["nn_t"] = np.array([x.trainIdx for x in m], np.int32)
    out["nn_d"] = np.array([x.distance for x in m], np.float64)
    knn = cv2.BFMatcher(normType=cv2.NORM_HAMMING).knnMatch(queryDescriptors=q, trainDescriptors=t, k=2)
    out["knn_t"] = np.array([[x.trainIdx for x in r] for r in knn], np.int32)
    out["knn_d"] = np.array([[x.distance for x in r] for r in knn], np.float64)
    cc = cv2.BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=True).match(queryDescriptors=q, trainDescriptors=t)
    out["cross_q"] = np.array([x.queryIdx for x in cc], np.int32)
    out["cross_t"] = np.array([x.trainIdx for x in cc], np.int32)
    out["cross_d"] = np.array([x.distance for x in cc], np.float64)
    # k_best = 2 branch of FeatureMatcher.match for ORB (flattened knn, then sorted): camera_models.py:418-444
    fm2 = FeatureMatcher("ORB", "BF", 2)
    m2 = fm2.match(query_descriptors=q, train_descriptors=t)
    out["k2_q"] = np.array([x.queryIdx for x in m2], np.int32)
    out["k2_t"] = np.array([x.trainIdx for x in m2], np.int32)
    out["k2_d"] = np.array([x.distance for x in m2], np.float64)
    # use_radius_match branch (camera_models.py:409-412): radiusMatch at a descriptor distance, flattened, sorted
    fmr = FeatureMatcher("ORB", "BF", 1, use_radius_match=True)
    for radius in (40, 70.5):
        mr = fmr.match(query_descriptors=q, train_descriptors=t, max_descriptor_distance_radius=radius)
        tag = f"r{int(radius)}"
        out[f"{tag}_radius"] = np.float64(radius)
        out[f"{tag}_q"] = np.array([x.queryIdx for x in mr], np.int32)
        out[f"{tag}_t"] = np.array([x.trainIdx for x in mr], np.int32)
        out[f"{tag}_d"] = np.array([x.distance for x in mr], np.float64)
    np.savez_compressed(os.path.join(OUT, "hamming.npz"), **out)


def make_sift_like(rng, n_land, n_q, n_t, dim=128, n_ties=6):
    """Integer-valued float32 descriptors with SIFT's look: sparse-ish gradient histograms in [0, 255] (cv2's SIFT scales
    the unit vector by 512, clips at 255 and rounds), each view a noisy copy of a landmark subset, plus unrelated rows and
    exact duplicates (ties)."""
    land = np.clip(rng.gamma(0.7, 28.0, (n_land, dim)), 0, 255).round()

    def view(n):
        ids = rng.permutation(n_land)[:n]
        d = np.clip(land[ids] + rng.normal(0, 6.0, (n, dim)).round(), 0, 255)
        return d.astype(np.float32)

    q, t = view(n_q), view(n_t)
    q[: n_q // 4] = np.clip(rng.gamma(0.7, 28.0, (n_q // 4, dim)), 0, 255).round()
    for _ in range(n_ties):
        a, b = rng.integers(0, n_t, 2)
        t[b] = t[a]
    return q, t


def gen_l2():
    """The SIFT branch of FeatureMatcher (camera_models.py:397-399, 417-442): cv2.BFMatcher() = NORM_L2."""
    from omnistereo.camera_models import FeatureMatcher
    rng = np.random.default_rng(12)
    out = {}
    q, t = make_sift_like(rng, 500, 380, 450)
    out["q"], out["t"] = q, t
    for name, method, k in (("sift1", "SIFT", 1), ("sift2", "SIFT", 2), ("surf2", "SURF", 2)):
        m = FeatureMatcher(method, "BF", k).match(query_descriptors=q, train_descriptors=t)
        out[f"{name}_q"] = np.array([x.queryIdx for x in m], np.int32)
        out[f"{name}_t"] = np.array([x.trainIdx for x in m], np.int32)
        out[f"{name}_d"] = np.array([x.distance for x in m], np.float32)
    q64, t64 = make_sift_like(rng, 300, 200, 190, dim=64)
    out["q64"], out["t64"] = q64, t64
    m = FeatureMatcher("SURF", "BF", 1).match(query_descriptors=q64, train_descriptors=t64)
    out["surf1_64_q"] = np.array([x.queryIdx for x in m], np.int32)
    out["surf1_64_t"] = np.array([x.trainIdx for x in m], np.int32)
    out["surf1_64_d"] = np.array([x.distance for x in m], np.float32)
    np.savez_compressed(os.path.join(OUT, "l2.npz"), **out)


def gen_matching_frames():
    """match_features_panoramic_top_bottom (camera_models.py:3027-3101) and match_features_frame_to_frame
    (pose_est_tools.py:211-269) on synthetic bucketed keypoints."""
    import cv2
    from omnistereo.camera_models import FeatureMatcher
    from omnistereo import pose_est_tools
    gs, _ = build_gums()
    rng = np.random.default_rng(2)
    gs.feature_matcher_for_static_stereo = FeatureMatcher("ORB", "BF", 1, percentage_good_matches=1.0)
    gs.feature_matcher_for_motion = FeatureMatcher("ORB", "BF", 1, percentage_good_matches=1.0)
    cols, rows = gs.top_model.panorama.cols, gs.top_model.panorama.rows
    n_buckets = 5
    kp_top, kp_bot, d_top, d_bot = [], [], [], []
    out = {}
    for b in range(n_buckets):
        n = [37, 0, 52, 18, 44][b]
        q, t = make_descriptors(rng, max(n, 1) + 20, n, max(n - 5, 0) if n else 0, n_ties=2 if n > 10 else 0) if n else \
            (np.zeros((0, 32), np.uint8), np.zeros((0, 32), np.uint8))
        # bottom = query, top = train (camera_models.py:3042)
        u0, u1 = b * cols / n_buckets, (b + 1) * cols / n_buckets
        pt_bot = np.stack([rng.uniform(u0, u1, len(q)), rng.uniform(0, rows * 0.6, len(q))], 1).astype(np.float32)
        pt_top = np.stack([rng.uniform(u0, u1, len(t)), rng.uniform(0, rows, len(t))], 1).astype(np.float32)
        # make most true pairs geometrically consistent: same column (+-2 px), top below bottom in v
        k = min(len(q), len(t))
        m = fm_pairs(q, t)
        for qi, ti in m[: int(0.7 * len(m))]:
            pt_top[ti, 0] = pt_bot[qi, 0] + np.float32(rng.uniform(-3.2, 3.2))
            pt_top[ti, 1] = pt_bot[qi, 1] + np.float32(rng.uniform(0.2, 30))
        kp_bot.append([cv2.KeyPoint(float(x), float(y), 7.0) for x, y in pt_bot])
        kp_top.append([cv2.KeyPoint(float(x), float(y), 7.0) for x, y in pt_top])
        d_bot.append(q)
        d_top.append(t)
        out[f"b{b}_pt_bot"], out[f"b{b}_pt_top"], out[f"b{b}_d_bot"], out[f"b{b}_d_top"] = pt_bot, pt_top, q, t
    np.random.seed(0)
    (m_top, k_top, dd_top), (m_bot, k_bot, dd_bot), _ = gs.match_features_panoramic_top_bottom(
        keypts_list_top=kp_top, desc_list_top=d_top, keypts_list_bot=kp_bot, desc_list_bot=d_bot,
        min_rectified_disparity=1, max_horizontal_diff=2.5, show_matches=False)
    out["stereo_m_top"], out["stereo_m_bot"] = m_top, m_bot
    out["stereo_desc_top"], out["stereo_desc_bot"] = dd_top, dd_bot
    out["n_buckets"] = np.int64(n_buckets)

    # temporal matching: train = reference frame, query = current frame (pose_est_tools.py:215)
    q, t = make_descriptors(rng, 260, 200, 190, n_ties=4)
    pq = np.stack([rng.uniform(0, cols, len(q)), rng.uniform(0, rows, len(q))], 1)
    pt = np.stack([rng.uniform(0, cols, len(t)), rng.uniform(0, rows, len(t))], 1)
    for qi, ti in fm_pairs(q, t)[:120]:
        pt[ti] = pq[qi] + rng.uniform(-40, 40, 2)
    pq = np.hstack([pq.astype(np.float32).astype(np.float64), np.ones((len(q), 1))])
    pt = np.hstack([pt.astype(np.float32).astype(np.float64), np.ones((len(t), 1))])
    kq = np.array([cv2.KeyPoint(float(x), float(y), 7.0) for x, y, _ in pq])
    kt = np.array([cv2.KeyPoint(float(x), float(y), 7.0) for x, y, _ in pt])
    colors = rng.integers(0, 256, (len(t), 3), dtype=np.uint8)
    max_du = 0.125 * 0.5 * cols  # pose_est_tools.py:866
    (ti_, _, _), (qi_, _, _), _ = pose_est_tools.match_features_frame_to_frame(
        cam_model=gs, train_kpts=kt, train_desc=t, query_kpts=kq, query_desc=q, random_colors_RGB=colors,
        max_horizontal_diff=max_du, keypts_as_points_train=pt, keypts_as_points_query=pq)
    out.update(f2f_q=q, f2f_t=t, f2f_pq=pq, f2f_pt=pt, f2f_max_du=np.float64(max_du),
               f2f_train_idx=np.asarray(ti_, np.int32), f2f_query_idx=np.asarray(qi_, np.int32))
    np.savez_compressed(os.path.join(OUT, "matching_frames.npz"), **out)


def fm_pairs(q, t):
    import cv2
    if len(q) == 0 or len(t) == 0:
        return []
    return [(m.queryIdx, m.trainIdx) for m in cv2.BFMatcher(normType=cv2.NORM_HAMMING).match(q, t)]


def gen_lifting():
    rng = np.random.default_rng(3)
    out = {}
    for tag, l1 in (("heik", 0.0), ("poly", 0.013)):
        gs, _ = build_gums(l1=l1)
        for name, m in (("top", gs.top_model), ("bot", gs.bot_model)):
            pre = f"{tag}_{name}_"
            out.update(flat(pre + "gum_", gum_dict(m)))
            # F9: omni pixel -> sphere; and the derived angles (camera_models.py:1183-1194)
            W, H = 320, 240
            uv = np.stack([rng.uniform(0, W, 257), rng.uniform(0, H, 257)], 1)[None]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                Ps = m.lift_pixel_to_unit_sphere_wrt_focus(uv)
                az, el = m.get_direction_angles_from_pixel(uv)
            out[pre + "omni_uv"], out[pre + "sphere"], out[pre + "omni_az"], out[pre + "omni_el"] = uv[0], Ps[0], az[0], el[0]
            # F3: 3D point -> pixel
            P = rng.normal(size=(1, 193, 4))
            P[..., 3] = 1.0
            u, v, _ = m.get_pixel_from_3D_point_wrt_M(P)
            out[pre + "proj_pts"], out[pre + "proj_u"], out[pre + "proj_v"] = P[0, :, :3], u[0], v[0]
    gs, _ = build_gums()
    top, bot = gs.top_model, gs.bot_model
    out.update(flat("pano_top_", pano_dict(top.panorama)))
    out.update(flat("pano_bot_", pano_dict(bot.panorama)))
    cols, rows = top.panorama.cols, top.panorama.rows
    # F7/F8: panorama pixels (float32-valued, including out-of-range ones) -> angles -> bearings
    m_pano = np.stack([rng.uniform(-5, cols + 5, 300), rng.uniform(-5, rows + 5, 300), np.ones(300)], 1)
    m_pano[:4, :2] = [[0, 0], [cols - 1e-3, rows - 1e-3], [cols, 3], [3, rows]]
    m_pano = m_pano.astype(np.float32).astype(np.float64)
    az, el = top.panorama.get_direction_angles_from_pixel_pano(m_pano, use_LUTs=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b = top.get_3D_point_from_angles_wrt_focus(azimuth=az, elevation=el)
    out["pano_px"], out["pano_az"], out["pano_el"], out["pano_bearing"] = m_pano, az, el, b[0, :, :3]
    # F10/F11: angles of true 3D points as seen from both foci (+ pixel noise), midpoint triangulation, range gate
    f1, f2 = top.F[:3, 0].copy(), bot.F[:3, 0].copy()
    n = 400
    dirs = rng.normal(size=(n, 3))
    dirs[:, 2] *= 0.3
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    P = dirs * rng.uniform(0.2, 12.0, (n, 1))

    def angles(P, f):
        d = P - f
        return np.arctan2(d[:, 1], d[:, 0]), np.arctan2(d[:, 2], np.hypot(d[:, 0], d[:, 1]))

    az1, el1 = angles(P, f1)
    az2, el2 = angles(P, f2)
    az1 += rng.normal(0, 2e-3, n); el1 += rng.normal(0, 2e-3, n); az2 += rng.normal(0, 2e-3, n); el2 += rng.normal(0, 2e-3, n)
    az1, el1, az2, el2 = (a.astype(np.float32).astype(np.float64) for a in (az1, el1, az2, el2))
    xyz_h = gs.get_triangulated_point_from_direction_angles(dir_angs_top=(az1, el1), dir_angs_bot=(az2, el2),
                                                            use_midpoint_triangulation=True)[0]
    good_h = gs.filter_panoramic_points_due_to_range(xyz_h, min_3D_range=0.5, max_3D_range=7.0)  # as pose_est_tools.py:372
    good_3 = gs.filter_panoramic_points_due_to_range(xyz_h[:, :3], min_3D_range=0.5, max_3D_range=7.0)
    out.update(tri_az1=az1, tri_el1=el1, tri_az2=az2, tri_el2=el2, tri_f1=f1, tri_f2=f2, tri_xyz_homo=xyz_h,
               tri_valid_homo=good_h, tri_valid_xyz=good_3)
    np.savez_compressed(os.path.join(OUT, "lifting.npz"), **out)


def gen_rgbd():
    from omnistereo.camera_models import RGBDCamModel, get_normalized_points
    rng = np.random.default_rng(4)
    out = {}
    h, w = 48, 64
    depth = rng.uniform(0.3, 9.0, (h, w)).astype(np.float32)
    depth[rng.random((h, w)) < 0.1] = 0.0
    u = rng.integers(0, w, 150)
    v = rng.integers(0, h, 150)
    out.update(depth=depth, u=u.astype(np.int32), v=v.astype(np.int32))
    for tag, kw in (("z", dict(fx=52.5, fy=52.5, center_x=31.5, center_y=23.5, depth_is_Z=True)),
                    ("radial", dict(fx=55.4256258, fy=55.4256258, center_x=31.5, center_y=23.5, depth_is_Z=False,
                                    focal_length_m=1.0 / 1000.0))):
        cam = RGBDCamModel(**kw)
        out[f"{tag}_cam"] = np.array([cam.fx, cam.fy, cam.center_x, cam.center_y, cam.focal_length_m, float(cam.depth_is_Z)])
        out[f"{tag}_depth_z"] = np.asarray(cam.get_depth_Z(depth=depth, uv_coords=None), np.float64)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            xyz = cam.get_XYZ(depth=depth, u_coords=u.astype(np.uint), v_coords=v.astype(np.uint))
            out[f"{tag}_xyz"] = xyz[0]
            out[f"{tag}_bearing"] = get_normalized_points(xyz)[0]
    np.savez_compressed(os.path.join(OUT, "rgbd.npz"), **out)


def gen_arun():
    from omnistereo import transformations as tr
    from omnistereo import pose_est_tools
    rng = np.random.default_rng(5)
    out = {}
    for k, n_sets in ((3, 64), (4, 8), (100, 4)):
        v0s, v1s, Ms = [], [], []
        for s in range(n_sets):
            R = tr.random_rotation_matrix(rng.random(3))[:3, :3]
            t = rng.normal(size=3)
            v0 = rng.normal(size=(3, k)) * rng.uniform(0.5, 5.0)
            v1 = R @ v0 + t[:, None] + rng.normal(0, 5e-3, (3, k))
            if k == 3 and s % 16 == 0:  # near-mirror configurations exercise the det < 0 branch
                v1[:, 0] += rng.normal(0, 0.5, 3)
            M = tr.superimposition_matrix(np.ascontiguousarray(v0), np.ascontiguousarray(v1), scale=False, usesvd=True)
            v0s.append(v0.T); v1s.append(v1.T); Ms.append(M[:3])
        out[f"k{k}_v0"], out[f"k{k}_v1"], out[f"k{k}_M"] = np.array(v0s), np.array(v1s), np.array(Ms)
    # the reference's own restatement of OpenGV's absolute-pose score (pose_est_tools.py:150-203)
    n = 120
    R = tr.random_rotation_matrix(rng.random(3))[:3, :3]
    t = rng.normal(size=3) * 0.1
    M = np.identity(4); M[:3, :3] = R; M[:3, 3] = t
    p_ref = rng.normal(size=(n, 3)) * 3
    f = (p_ref - t) @ R
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    f += rng.normal(0, 0.02, f.shape)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    idx = np.arange(n)
    scores = pose_est_tools.get_selected_distances_to_model(M, idx, p_ref, f, False)
    thr = 1.0 - np.cos(np.deg2rad(5.0))
    inl, outl = pose_est_tools.select_inliers_within_distance(M, idx, thr, p_ref, f, False)
    out.update(score_M=M[:3], score_p_ref=p_ref, score_f=f, score_values=np.array(scores), score_thr=np.float64(thr),
               score_inliers=np.asarray(inl, np.int64))

    class _T:  # compute_num_of_iterations_RANSAC needs no state
        pass
    out["ransac_iters_default"] = np.int64(pose_est_tools.TrackerSE3.compute_num_of_iterations_RANSAC(_T(), 3, 0.65))
    np.savez_compressed(os.path.join(OUT, "arun.npz"), **out)


def gen_driver():
    """Pose bookkeeping of run_VO (pose_est_tools.py:1510-1512, 1609-1612): the reference's metrics, quaternion and
    matrix chaining on random rigid transforms (exact-rotation and slightly non-orthonormal, as float32 poses are)."""
    from omnistereo import transformations as tr
    rng = np.random.default_rng(6)
    T = np.zeros((40, 4, 4))
    for i in range(40):
        M = tr.random_rotation_matrix(rng.random(3))
        M[:3, 3] = rng.normal(size=3) * (0.001 if i % 4 == 0 else 0.3)
        if i % 3 == 0:  # small rotation, like a frame-to-keyframe motion
            M[:3, :3] = tr.rotation_matrix(rng.normal() * 0.02, rng.normal(size=3))[:3, :3]
        if i % 5 == 0:
            M[:3, :3] = M[:3, :3].astype(np.float32).astype(np.float64)
        T[i] = M
    out = dict(T=T,
               quat=np.array([tr.quaternion_from_matrix(matrix=M, isprecise=False) for M in T]),
               trans=np.array([tr.translation_from_matrix(matrix=M) for M in T]),
               dist=np.array([tr.rpe_translation_metric(M) for M in T]),
               angle=np.array([tr.rpe_rotation_metric(M) for M in T]),
               chain=np.array([tr.concatenate_matrices(T[i], T[i + 1]) for i in range(39)]))
    np.savez_compressed(os.path.join(OUT, "driver.npz"), **out)


def gen_dense():
    """Dense triangulation of a panoramic disparity map: resolve_pano_correspondences_from_disparity_map
    (camera_models.py:2492-2538) + the lifting / midpoint triangulation of triangulate_from_depth_map (:2567-2685,
    use_opengv_triangulation=False) on a synthetic disparity map with zeros, out-of-range values and an ROI."""
    gs, _ = build_gums()
    top, bot = gs.top_model, gs.bot_model
    rows, cols = top.panorama.rows, top.panorama.cols
    rng = np.random.default_rng(8)
    disp = rng.uniform(0.0, 14.0, (rows, cols)).astype(np.float32)
    disp[rng.random((rows, cols)) < 0.3] = 0.0
    disp[:, ::7] = np.round(disp[:, ::7])          # integer disparities, as a block matcher produces
    disp = np.minimum(disp, np.arange(rows, dtype=np.float32)[:, None])   # the match stays inside the bottom panorama
    out = dict(disparity=disp)
    out.update(flat("pano_top_", pano_dict(top.panorama)))
    out.update(flat("pano_bot_", pano_dict(bot.panorama)))
    out["f1"], out["f2"] = top.F[:3, 0].copy(), bot.F[:3, 0].copy()
    out["lowest_reference_row"] = np.float64(bot.panorama.get_panorama_row_from_elevation(bot.lowest_elevation_angle))
    ref_uv = np.transpose(np.indices(disp.shape[::-1]), (1, 2, 0))
    for tag, kw in (("all", dict(min_disparity=1, max_disparity=0, roi_cols=None)),
                    ("roi", dict(min_disparity=2, max_disparity=9, roi_cols=(30, 150)))):
        gs.disparity_map = disp
        tp, bp, dd = gs.resolve_pano_correspondences_from_disparity_map(ref_uv, **kw)
        az1, el1 = top.panorama.get_direction_angles_from_pixel_pano(tp, use_LUTs=False)
        az2, el2 = bot.panorama.get_direction_angles_from_pixel_pano(bp, use_LUTs=False)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            xyz = gs.get_triangulated_point_from_direction_angles(dir_angs_top=(az1, el1), dir_angs_bot=(az2, el2),
                                                                  use_midpoint_triangulation=True)
        out[tag + "_top_px"], out[tag + "_bot_px"], out[tag + "_disp"], out[tag + "_xyz"] = tp[0], bp[0], dd, xyz[0, :, :3]
        out[tag + "_args"] = np.array([kw["min_disparity"], kw["max_disparity"], *(kw["roi_cols"] or (-1, -1))], np.float64)
    np.savez_compressed(os.path.join(OUT, "dense.npz"), **out)


def main():
    if not os.path.isdir("/root/reference/omnistereo"):
        sys.exit("gen_golden needs /root/reference (build container only)")
    from . import _ref_shims
    _ref_shims.install()
    os.makedirs(OUT, exist_ok=True)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):  # the reference prints progress
        gen_remap()
        gen_hamming()
        gen_l2()
        gen_matching_frames()
        gen_lifting()
        gen_rgbd()
        gen_arun()
        gen_driver()
        gen_dense()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
