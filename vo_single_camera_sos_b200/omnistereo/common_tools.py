"""Host helpers of omnistereo/common_tools.py that the demos and the VO driver use: paths, pickles,
unit factors and TUM pose files.  Plain host code, no kernels."""
import os
import pickle

import numpy as np


def make_sure_path_exists(path):
    """common_tools.py: create the directory (and parents) if it is missing."""
    os.makedirs(path, exist_ok=True)


def str2bool(v):
    """argparse helper of the demos (demo_vo_sos.py:31)."""
    return str(v).lower() in ("yes", "true", "t", "1", "y")


def load_obj_from_pickle(filename):
    with open(filename, "rb") as f:
        return pickle.load(f)


def save_obj_in_pickle(obj_instance, filename, locals=None):
    with open(filename, "wb") as f:
        pickle.dump(obj_instance, f, protocol=pickle.HIGHEST_PROTOCOL)


def get_length_units_conversion_factor(input_units, output_units):
    """common_tools.py:580-597."""
    scale = {"mm": 1e-3, "cm": 1e-2, "m": 1.0}
    return scale[input_units] / scale[output_units]


def _transform_from_tum(p):
    """[tx, ty, tz, qx, qy, qz, qw] -> 4x4 (transformations.transform44_from_TUM_entry)."""
    tx, ty, tz, x, y, z, w = (float(v) for v in p)
    n = x * x + y * y + z * z + w * w
    T = np.identity(4)
    if n > np.finfo(float).eps * 4.0:
        s = 2.0 / n
        T[:3, :3] = [[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                     [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                     [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]]
    T[:3, 3] = [tx, ty, tz]
    return T


def get_poses_from_file(poses_filename, input_units="m", output_working_units="m", indices=None, pose_format="tum",
                        zero_up_wrt_origin=False, initial_T=None, delimiter=None):
    """TUM pose files (`stamp tx ty tz qx qy qz qw` per line) -> (list of 7-vectors in TUM order, list of 4x4 transforms),
    common_tools.py:624-736; `zero_up_wrt_origin` re-expresses every pose wrt the first valid one.  Only the "tum" format is
    mirrored (the POV-Ray format belongs to the synthetic-data tooling)."""
    if pose_format.lower() != "tum":
        raise NotImplementedError("only TUM pose files are supported")
    grid = np.loadtxt(poses_filename, delimiter=delimiter or " ", usecols=range(8), comments="#", ndmin=2)
    f = get_length_units_conversion_factor(input_units, output_working_units)
    if indices is None or len(indices) == 0:
        indices = range(len(grid))
    poses7, Ts = [], []
    T0_inv = None
    for i in indices:
        row = grid[i]
        if np.any(np.isnan(row)):
            poses7.append(7 * [np.nan])
            Ts.append(np.full((4, 4), np.nan))
            continue
        p = [f * row[1], f * row[2], f * row[3], row[4], row[5], row[6], row[7]]
        T = _transform_from_tum(p)
        if initial_T is not None:
            T = np.asarray(initial_T) @ T
        if zero_up_wrt_origin:
            if T0_inv is None:
                T0_inv = np.linalg.inv(T)
            T = T0_inv @ T
        poses7.append(p)
        Ts.append(T)
    return poses7, Ts
