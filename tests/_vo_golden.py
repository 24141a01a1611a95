"""Shared helpers for the tests that consume tests/golden/vo_sequence.npz (the reference's own run_VO on a rendered
sequence, oracle/gen_vo_golden.py)."""
import numpy as np

from conftest import load_golden


def load():
    g = load_golden("vo_sequence.npz")
    n = int(g["n_frames"])
    frames = [dict(px_top=g[f"f{i}_px_top"], desc_top=g[f"f{i}_desc_top"], boff_top=g[f"f{i}_boff_top"],
                   px_bot=g[f"f{i}_px_bot"], desc_bot=g[f"f{i}_desc_bot"], boff_bot=g[f"f{i}_boff_bot"]) for i in range(n)]
    return g, frames


def rig_matrix(g):
    """[Rc | tc] per camera as the reference's tracker hands them to OpenGV (cam_rotations, cam_offsets; pose_est_tools.py:852-859)."""
    return np.concatenate([np.asarray(g["cam_rotations"], np.float64).reshape(-1, 3, 3),
                           np.asarray(g["cam_offsets"], np.float64).reshape(-1, 3, 1)], axis=2)


def hypothesis_list(g, k=4):
    return np.random.default_rng(int(g["ransac_seed"])).integers(0, 2 ** 32, (int(g["max_iterations"]), k), dtype=np.uint64).astype(np.uint32)


def tum_to_matrix(row):
    """TUM row [id tx ty tz qx qy qz qw] -> 4x4."""
    x, y, z, w = row[4:8]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, row[1:4]
    return T


PANO_KEYS = ("cols", "rows", "pixel_size", "cyl_height_max", "cyl_circumference", "cyl_radius")


def pano_geometry(g, view="top"):
    """The reference's panorama geometry as the dict oracle/geometry.py takes."""
    d = dict(zip(PANO_KEYS, [float(x) for x in g[f"pano_{view}"]]))
    d["cols"], d["rows"] = int(d["cols"]), int(d["rows"])
    d["cyl_height_min"] = float(g[f"pano_{view}_cyl_height_min"])
    return d
