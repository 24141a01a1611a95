"""Mirror of the hot-path part of omnistereo/transformations.py: the Arun / Kabsch / Umeyama registration."""
import numpy as np

from . import device_context, to_device


def superimposition_matrix(v0, v1, scale=False, usesvd=True):
    """4 x 4 matrix M with v1 ~ M v0 for (3 or 4) x K point arrays, as transformations.superimposition_matrix
    (transformations.py:982-1030) -> affine_matrix_from_points(shear=False, usesvd=True) (transformations.py:942-980).
    Solved by sos_arun_batch on the device (float64 Kabsch; Umeyama scale when scale=True)."""
    if not usesvd:
        raise NotImplementedError("only the SVD (Kabsch) branch of the reference is mirrored")
    v0 = np.array(v0, dtype=np.float64)[:3]
    v1 = np.array(v1, dtype=np.float64)[:3]
    if v0.shape != v1.shape or v0.shape[1] < 3:
        raise ValueError("input arrays are of wrong shape or type")
    ctx = device_context()
    M, ok = ctx.arun_batch(to_device(v0.T[None]), to_device(v1.T[None]), with_scale=bool(scale))
    if not bool(ok.cpu().numpy()[0]):
        raise ValueError("degenerate point configuration (rank < 2)")
    out = np.identity(4)
    out[:3] = M.cpu().numpy()[0]
    return out


def concatenate_matrices(*matrices):
    """transformations.py:1803: product of the given matrices (host glue, 4 x 4)."""
    M = np.identity(4)
    for m in matrices:
        M = np.dot(M, m)
    return M


def identity_matrix():
    return np.identity(4)


def inverse_matrix(matrix):
    return np.linalg.inv(matrix)


def rotation_matrix(angle, direction, point=None):
    """4 x 4 rotation by `angle` [rad] about the axis `direction` (through `point`), transformations.py:297-338 (Rodrigues)."""
    d = np.asarray(direction, np.float64)[:3]
    d = d / np.linalg.norm(d)
    s, c = np.sin(angle), np.cos(angle)
    K = np.array([[0.0, -d[2], d[1]], [d[2], 0.0, -d[0]], [-d[1], d[0], 0.0]])
    M = np.identity(4)
    M[:3, :3] = c * np.identity(3) + (1.0 - c) * np.outer(d, d) + s * K
    if point is not None:
        p = np.asarray(point, np.float64)[:3]
        M[:3, 3] = p - M[:3, :3] @ p
    return M


def translation_from_matrix(matrix):
    return np.array(matrix, copy=True)[:3, 3]


def quaternion_from_matrix(matrix, isprecise=False):
    """[w, x, y, z], w >= 0 (transformations.py:1258-1332, the isprecise=False branch for both settings)."""
    from ..driver import quaternion_wxyz
    return quaternion_wxyz(np.asarray(matrix, np.float64))


def rpe_translation_metric(T):
    from ..driver import translation_metric
    return translation_metric(np.asarray(T))


def rpe_rotation_metric(T):
    from ..driver import rotation_metric
    return rotation_metric(np.asarray(T))
