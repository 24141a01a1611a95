"""Oracle for steps 3+4 (lifting, triangulation), the GUM projection / LUT generation and the RGB-D path.
TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  Everything is float64 NumPy, following the reference line by line
in meaning (not in text)."""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# Panorama geometry
# ---------------------------------------------------------------------------------------------------------------
def pano_geometry(cols: int, elev_hi: float, elev_lo: float, cyl_radius: float = 1.0) -> dict:
    """Panorama._set_cylinder_height + _resolve_dimensions_pixel_sizing (panorama.py:142-172) for a given width."""
    h_max = cyl_radius * np.tan(elev_hi)
    h_min = cyl_radius * np.tan(elev_lo)
    circ = 2 * np.pi * cyl_radius
    pixel_size = circ / float(cols)
    rows = int(np.ceil((h_max - h_min) / pixel_size))
    return dict(cols=int(cols), rows=rows, pixel_size=pixel_size, cyl_height_max=h_max, cyl_height_min=h_min,
                cyl_circumference=circ, cyl_radius=cyl_radius)


def pano_vector(g: dict) -> np.ndarray:
    return np.array([g["cols"], g["rows"], g["pixel_size"], g["cyl_height_max"], g["cyl_circumference"], g["cyl_radius"]],
                    np.float64)


def pano_pixel_to_angles(g: dict, m_pano: np.ndarray):
    """Panorama.get_direction_angles_from_pixel_pano(use_LUTs=False) (panorama.py:616-642, 650-666)."""
    u = np.asarray(m_pano[..., 0], np.float64)
    v = np.asarray(m_pano[..., 1], np.float64)
    az = np.where((0.0 <= u) & (u < g["cols"]), g["cyl_circumference"] - g["pixel_size"] * u, np.nan)
    el = np.where((0.0 <= v) & (v < g["rows"]), np.arctan2(g["cyl_height_max"] - g["pixel_size"] * v, g["cyl_radius"]), np.nan)
    return az, el


def angles_to_sphere(az: np.ndarray, el: np.ndarray) -> np.ndarray:
    """OmniCamModel.map_angles_to_unit_sphere (camera_models.py:1031-1065) without the homogeneous 1."""
    ok = ~np.isnan(el)
    b = np.where(ok, np.cos(el), np.nan)
    z = np.where(ok, np.sin(el), np.nan)
    return np.stack([b * np.cos(az), b * np.sin(az), z], axis=-1)


def triangulate_midpoint(az1, el1, az2, el2, f1, f2) -> np.ndarray:
    """get_triangulated_point_from_direction_angles(use_midpoint_triangulation=True) (camera_models.py:3323-3364) ->
    get_triangulated_midpoint (:2420-2479) -> triangulate_for_skew_rays (:2481-2490)."""
    v1 = np.stack([np.cos(az1), np.sin(az1), np.tan(el1)], axis=-1)
    v2 = np.stack([np.cos(az2), np.sin(az2), np.tan(el2)], axis=-1)
    perp = np.cross(v1, v2)
    with np.errstate(invalid="ignore", divide="ignore"):
        unit = perp / np.linalg.norm(perp, axis=-1, keepdims=True)
    M = np.stack([v1, -v2, unit], axis=-1)  # columns
    b = np.broadcast_to((np.asarray(f2, np.float64) - np.asarray(f1, np.float64))[:, None], v1.shape + (1,))
    out = np.full(v1.shape, np.nan)
    good = np.all(np.isfinite(M), axis=(-1, -2))
    good &= np.abs(np.linalg.det(np.where(good[..., None, None], M, np.eye(3)))) > 0
    if good.any():
        lam = np.linalg.solve(M[good], b[good])[..., 0]
        g1 = np.asarray(f1, np.float64) + lam[:, 0:1] * v1[good]
        out[good] = g1 + lam[:, 2:3] / 2.0 * unit[good]
    return out


def range_filter(xyz: np.ndarray, min_range: float = 0.0, max_range: float = 0.0) -> np.ndarray:
    """OmniStereoModel.filter_panoramic_points_due_to_range (camera_models.py:3299-3321).  NOTE: the reference's frame
    code passes the HOMOGENEOUS N x 4 array (pose_est_tools.py:365-372), so the norm there includes the trailing 1."""
    ok = np.ones(xyz.shape[:-1], bool)
    if min_range > 0 or max_range > 0:
        with np.errstate(invalid="ignore"):
            nrm = np.linalg.norm(xyz, axis=-1)
            if min_range > 0:
                ok &= nrm >= min_range
            if max_range > 0:
                ok &= nrm <= max_range
    return ok


def dense_triangulate(pano_top: dict, pano_bot: dict, disparity: np.ndarray, f1, f2, min_disparity: float = 1.0,
                      max_disparity: float = 0.0, lowest_reference_row: float = np.inf, roi_cols=None):
    """Dense triangulation of a panoramic disparity map: the validity chain of
    resolve_pano_correspondences_from_disparity_map (camera_models.py:2492-2538: ROI columns, d != 0, min <= d <= max with
    max = 0 meaning the map's maximum, v - d <= lowest reference row), target pixel (u, v - d) in the bottom panorama,
    then the lifting + midpoint triangulation of triangulate_from_depth_map (:2567-2685, own midpoint method).
    -> xyz [rows, cols, 3] float64 (NaN where invalid), valid [rows, cols] bool."""
    d = np.asarray(disparity, np.float64)
    rows, cols = d.shape
    if roi_cols is not None:
        dd = np.zeros_like(d)
        dd[:, roi_cols[0]:roi_cols[1]] = d[:, roi_cols[0]:roi_cols[1]]
        d = dd
    if max_disparity == 0:
        max_disparity = d.max()
    v, u = np.mgrid[:rows, :cols].astype(np.float64)
    valid = (d != 0) & (min_disparity <= d) & (d <= max_disparity) & (v - d <= lowest_reference_row)
    az1, el1 = pano_pixel_to_angles(pano_top, np.stack([u, v], -1))
    az2, el2 = pano_pixel_to_angles(pano_bot, np.stack([u, v - d], -1))
    xyz = triangulate_midpoint(az1, el1, az2, el2, f1, f2)
    xyz[~valid] = np.nan
    return xyz, valid


# ---------------------------------------------------------------------------------------------------------------
# GUM
# ---------------------------------------------------------------------------------------------------------------
GUM_FIELDS = ("xi1", "xi2", "xi3", "k1", "k2", "k3", "gamma1", "gamma2", "alpha_c", "u_center", "v_center",
              "l1", "l2", "l3", "p1", "p2", "plane_k", "use_distortion")


def gum_vector(p: dict) -> np.ndarray:
    return np.array([float(p[k]) for k in GUM_FIELDS], np.float64)


def gum_project(p: dict, pts: np.ndarray):
    """GUM.get_pixel_from_3D_point_wrt_M (gum.py:2512-2551): normalise, shift by Cp (gum.py:1368), divide by |z|
    (gum.py:1378-1381), radial distortion (gum.py:1383-1385, 2942-2959), generalised camera matrix (gum.py:2554-2562)."""
    P = np.asarray(pts, np.float64)
    Ps = P[..., :3] / np.linalg.norm(P[..., :3], axis=-1, keepdims=True)
    q = Ps - np.array([p["xi1"], p["xi2"], p["xi3"]])
    pu = q[..., :2] / np.abs(q[..., 2:3])
    if p["use_distortion"]:
        r2 = pu[..., 0] ** 2 + pu[..., 1] ** 2
        f = 1.0 + p["k1"] * r2 + p["k2"] * r2 ** 2 + p["k3"] * r2 ** 3
        pu = pu * f[..., None]
    u = p["gamma1"] * pu[..., 0] + p["gamma1"] * p["alpha_c"] * pu[..., 1] + p["u_center"]
    v = p["gamma2"] * pu[..., 1] + p["v_center"]
    return u, v


def lut_build(p: dict, rows: int, cols: int, h_max: float, h_min: float, elev_lo: float, elev_hi: float):
    """Panorama._generate_LUTs (panorama.py:414-478): psi/theta grids rounded to float32 as the reference does."""
    psi = np.linspace(0, 2 * np.pi, num=cols, endpoint=False)[::-1].astype("float32").astype(np.float64)
    h = np.linspace(h_max, h_min, num=rows, endpoint=False)
    theta = np.arctan2(h, 1.0)
    theta = np.where((elev_lo <= theta) & (theta <= elev_hi), theta, np.nan).astype("float32").astype(np.float64)
    psi2, theta2 = np.meshgrid(psi, theta)
    with np.errstate(invalid="ignore"):
        return gum_project(p, angles_to_sphere(psi2, theta2))


def gum_lift(p: dict, m: np.ndarray):
    """GUM.lift_pixel_to_unit_sphere_wrt_focus, new_method branch (gum.py:2653-2762) with the line/sphere intersection
    of camera_models.py:135-187 (first root), and get_direction_angles_from_pixel (camera_models.py:1183-1194)."""
    u = np.asarray(m[..., 0], np.float64)
    v = np.asarray(m[..., 1], np.float64)
    g1, g2, ac, u0, v0 = p["gamma1"], p["gamma2"], p["alpha_c"], p["u_center"], p["v_center"]
    xd = (1 / g1) * u + (-ac / g2) * v + (ac * v0 / g2 - u0 / g1)
    yd = (1 / g2) * v + (-v0 / g2)
    xu, yu = xd, yd
    if p["use_distortion"]:
        if p["l1"] != 0:
            r2 = xd ** 2 + yd ** 2
            f = 1.0 + p["l1"] * r2 + p["l2"] * r2 ** 2 + p["l3"] * r2 ** 3
            xu, yu = xd * f, yd * f
        else:
            k1, k2, p1, p2 = p["k1"], p["k2"], p["p1"], p["p2"]
            x2, y2, xy = xd * xd, yd * yd, xd * yd
            r2 = x2 + y2
            r4 = r2 ** 2
            rad = k1 * r2 + k2 * r4
            dx = xd * rad + p2 * (r2 + 2 * x2) + 2 * p1 * xy
            dy = yd * rad + p1 * (r2 + 2 * y2) + 2 * p2 * xy
            inv = 1 / (1 + 4 * k1 * r2 + 6 * k2 * r4 + 8 * p1 * yd + 8 * p2 * xd)
            xu, yu = xd - inv * dx, yd - inv * dy
    cp = np.array([p["xi1"], p["xi2"], p["xi3"]])
    pt = np.stack([cp[0] + xu, cp[1] + yu, np.zeros_like(xu) + p["plane_k"]], axis=-1)
    d = pt - cp
    a = np.sum(d ** 2, axis=-1)
    b = 2 * np.sum(d * pt, axis=-1)
    c = np.sum(pt ** 2, axis=-1) - 1.0
    with np.errstate(invalid="ignore"):
        t = (-b + np.sqrt(b ** 2 - 4 * a * c)) / (2 * a)
    Ps = pt + t[..., None] * d
    with np.errstate(invalid="ignore"):
        az = np.arctan2(Ps[..., 1], Ps[..., 0])
        el = np.arcsin(Ps[..., 2])
    return Ps, az, el


# ---------------------------------------------------------------------------------------------------------------
# RGB-D (camera_models.py:750-860, pose_est_tools.py:570-623)
# ---------------------------------------------------------------------------------------------------------------
RGBD_FIELDS = ("fx", "fy", "center_x", "center_y", "focal_length_m", "depth_is_Z")


def rgbd_vector(c: dict) -> np.ndarray:
    return np.array([float(c[k]) for k in RGBD_FIELDS], np.float64)


def rgbd_depth_to_z(c: dict, depth: np.ndarray) -> np.ndarray:
    """RGBDCamModel.get_depth_Z (camera_models.py:781-799)."""
    depth = np.asarray(depth, np.float64)
    if c["depth_is_Z"]:
        return depth
    h, w = depth.shape[-2:]
    uu, vv = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    f = c["focal_length_m"]
    xi = (f / c["fx"]) * (uu - c["center_x"])
    yi = (f / c["fy"]) * (vv - c["center_y"])
    return f * depth / np.sqrt(xi ** 2 + yi ** 2 + f ** 2)


def rgbd_backproject(c: dict, depth: np.ndarray, u: np.ndarray, v: np.ndarray, zmin: float = 0.0, zmax: float = 0.0):
    """RGBDCamModel.get_XYZ at keypoints (camera_models.py:835-860), get_normalized_points (:203-212) and the NaN +
    |Z|-range gate of RGBDFrame.establish_keypoints (pose_est_tools.py:612-620, 570-592)."""
    z_map = rgbd_depth_to_z(c, depth)
    z_map = np.where(z_map != 0, z_map, np.nan)
    Z = z_map[v, u]
    X = (u - c["center_x"]) * Z / c["fx"]
    Y = (v - c["center_y"]) * Z / c["fy"]
    xyz = np.stack([X, Y, Z], axis=-1)
    with np.errstate(invalid="ignore"):
        bearing = xyz / np.linalg.norm(xyz, axis=-1, keepdims=True)
    valid = ~np.isnan(Z)
    az = np.nan_to_num(np.abs(Z))
    if zmin > 0:
        valid &= az >= zmin
    if zmax > 0:
        valid &= az <= zmax
    return xyz, bearing, valid
