"""Mirror of the names the demos and the VO driver import from omnistereo/common_plot.py (3 285 lines of vispy / matplotlib
drawing in the reference — out of scope).  Importing this module never needs a display; what draws says so
instead of dying with ModuleNotFoundError (demo_vo_sos.py:88-90 reaches replay_VO_visualization with
only_visualize_frames_from_existing_VO_poses_file=True)."""
import numpy as np


class VisualisationUnavailable(RuntimeError):
    pass


_MSG = ("the B200 drop-in of the SOS front-end does not include the reference's vispy / matplotlib visualisation "
        "(omnistereo/common_plot.py); the VO results are written to estimated_frame_poses_TUM.txt / keyframe_ids.txt")


def replay_VO_visualization(scene_path_vo_results, first_image_index=0, last_image_index=-1, step_for_poses=1, vis_name=""):
    """common_plot.py:506-517 replays a finished run in a vispy window.  Here: raise a clear error naming the result files."""
    raise VisualisationUnavailable(f"replay_VO_visualization({scene_path_vo_results!r}): {_MSG}")


class DrawerVO(object):
    """common_plot.py:206-442.  run_VO only touches it when a visualiser is passed (visualize_VO=True)."""

    def __init__(self, *args, **kwargs):
        raise VisualisationUnavailable(f"DrawerVO: {_MSG}")


def draw_matches_between_frames(pano_img_train, pano_img_query, matched_kpts_train, matched_kpts_query, random_colors=None,
                                win_name="Matches", show_window=False):
    """common_plot.py:1152-1200 paints the two panoramas on top of each other with lines between the matched keypoints and
    returns the image (track_frame saves it when save_correspondence_images is on).  Same layout, drawn with cv2 only."""
    import cv2
    top, bot = np.asarray(pano_img_train), np.asarray(pano_img_query)
    if top.ndim == 2:
        top, bot = cv2.cvtColor(top, cv2.COLOR_GRAY2BGR), cv2.cvtColor(bot, cv2.COLOR_GRAY2BGR)
    out = np.vstack([top, bot]).copy()
    dy = top.shape[0]
    for i, (a, b) in enumerate(zip(matched_kpts_train, matched_kpts_query)):
        pa = a.pt if hasattr(a, "pt") else a[:2]
        pb = b.pt if hasattr(b, "pt") else b[:2]
        col = (0, 255, 0) if random_colors is None or len(random_colors) <= i else tuple(int(c) for c in random_colors[i][::-1])
        cv2.line(out, (int(round(pa[0])), int(round(pa[1]))), (int(round(pb[0])), int(round(pb[1])) + dy), col, 1)
    return out
