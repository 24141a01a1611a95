"""Mirror of the hot-path part of omnistereo/common_cv.py."""
import numpy as np
import torch

from . import device_context, to_device


def filter_pixel_correspondences(matched_points_top, matched_points_bot, min_rectified_disparity, max_horizontal_diff):
    """Same contract as the reference (common_cv.py:167-188): boolean validity of each matched pixel pair,
    `|u_top - u_bot| <= max_horizontal_diff` (when > 0) and `v_top - v_bot >= min_rectified_disparity` (when >= 0).
    Evaluated by sos_pixel_gate on the device in float64."""
    top = np.asarray(matched_points_top, np.float64)
    bot = np.asarray(matched_points_bot, np.float64)
    shape = top.shape[:-1]
    if top.size == 0:
        return np.ones(shape, dtype=bool)
    ctx = device_context()
    valid = ctx.pixel_gate(to_device(top[..., :2].reshape(-1, 2)), to_device(bot[..., :2].reshape(-1, 2)),
                           float(max_horizontal_diff), float(min_rectified_disparity))
    return valid.cpu().numpy().astype(bool).reshape(shape)


def rgb2bgr_color(rgb_color):
    return (rgb_color[2], rgb_color[1], rgb_color[0])


def clean_up(wait_key_time=0):
    """The reference closes HighGUI windows here (common_cv.py:79-81); headless OpenCV has none."""
    try:
        import cv2
        cv2.destroyAllWindows()
    except Exception:
        pass


def get_images(filename_template, indices_list=[], show_images=False, return_names_only=False):
    """common_cv.py:1293-1340: the files matching the template (fnmatch on the directory listing); here the names are sorted
    so that the order does not depend on the file system.  With return_names_only=False the images are read with cv2."""
    import fnmatch
    from os import listdir
    from os.path import join, split
    import cv2
    path, pattern = split(filename_template)
    names = sorted(fnmatch.filter(listdir(path), pattern))
    if indices_list is None or len(indices_list) == 0:
        indices_list = range(len(names))
    chosen = [join(path, names[i]) for i in indices_list]
    if return_names_only:
        return chosen
    return [cv2.imread(fn) for fn in chosen]


def get_depthmap_float32_from_png(depth_img_filename, conversion_factor=1. / 1000.0):
    """common_cv.py: 16-bit depth PNG -> float32 map in metres (0 where there is no measurement)."""
    import cv2
    d = cv2.imread(depth_img_filename, cv2.IMREAD_UNCHANGED)
    return d.astype(np.float32) * np.float32(conversion_factor)
