"""Mirror of the frame front-end and trackers of omnistereo/pose_est_tools.py, same names and defaults.

Differences from the reference that are visible to a caller:
  * the RANSAC behind `pyopengv.absolute_pose_*_ransac` hypothesises with 3-point Arun registration on 3D-3D
    correspondences under a seeded hypothesis list (see vo_single_camera_sos_b200/pyopengv.py); the trackers therefore
    also hand the current frame's triangulated points to it;
  * `*_optimize_nonlinear` is Levenberg-Marquardt on OpenGV's bearing residual over the inliers (sos_refine_pose);
  * frames accept pre-computed features (`features=`) because feature detection is upstream of the hot path.
"""
from math import log10, sqrt

import cv2
import numpy as np

from .. import pyopengv
from ..pyopengv import (absolute_pose_noncentral_optimize_nonlinear, absolute_pose_noncentral_ransac,
                        absolute_pose_optimize_nonlinear, absolute_pose_ransac)
from . import transformations as tr
from .camera_models import FeatureMatcher, KeyPointAndDescriptor, PanoramicCorrespondences, get_normalized_points


def get_length_units_conversion_factor(input_units, output_units):
    """common_tools.py:580-597."""
    scale = {"mm": 1e-3, "cm": 1e-2, "m": 1.0}
    return scale[input_units] / scale[output_units]


def match_features_frame_to_frame(cam_model, train_kpts, train_desc, query_kpts, query_desc, random_colors_RGB,
                                  max_horizontal_diff=-1, max_descriptor_distance_radius=-1, keypts_as_points_train=None,
                                  keypts_as_points_query=None, pano_img_train=None, pano_img_query=None, show_matches=False,
                                  win_name="Matches (Frame-to-Frame)"):
    """Temporal matching, query = current frame, train = reference frame (pose_est_tools.py:211-269); the |du| gate of
    filter_pixel_correspondences is fused into the match-select kernel."""
    fm = cam_model.feature_matcher_for_motion
    train_kpts, query_kpts = np.asarray(train_kpts), np.asarray(query_kpts)
    if keypts_as_points_train is None:
        keypts_as_points_train = cv2.KeyPoint_convert(list(train_kpts)).astype(float).reshape(-1, 2)
    if keypts_as_points_query is None:
        keypts_as_points_query = cv2.KeyPoint_convert(list(query_kpts)).astype(float).reshape(-1, 2)
    gate = max_horizontal_diff >= 0
    qi, ti, _ = fm.match_arrays(query_desc, train_desc, max_descriptor_distance_radius,
                                px_query=np.asarray(keypts_as_points_query)[:, :2] if gate else None,
                                px_train=np.asarray(keypts_as_points_train)[:, :2] if gate else None,
                                max_horizontal_diff=float(max_horizontal_diff), min_rectified_disparity=-1.0)
    n_good = int(fm.percentage_good_matches * len(qi)) if not gate else len(qi)
    qi, ti = qi[:n_good], ti[:n_good]
    if len(qi) == 0:
        return ([], [], []), ([], [], []), []
    colors = np.asarray(random_colors_RGB)[ti] if len(random_colors_RGB) else []
    return ((ti, train_kpts[ti], np.asarray(train_desc)[ti]), (qi, query_kpts[qi], np.asarray(query_desc)[qi]), colors)


class StereoPanoramicFrame(object):
    """pose_est_tools.py:271-402."""

    def __init__(self, stereo_camera_model, frame_id, **kwargs):
        self.frame_id = frame_id
        self.parent_id = kwargs.get("parent_id", -1)
        self.T_frame_wrt_tracking_ref_frame = np.identity(4)
        top_p, bot_p = stereo_camera_model.top_model.panorama, stereo_camera_model.bot_model.panorama
        self.panoramic_image_top = None if top_p is None or top_p.panoramic_img is None else top_p.panoramic_img.copy()
        self.panoramic_image_bottom = None if bot_p is None or bot_p.panoramic_img is None else bot_p.panoramic_img.copy()
        self.use_midpoint_triangulation = True
        self.use_opengv_triangulation = False
        self.conversion_factor_length_to_m = get_length_units_conversion_factor(stereo_camera_model.units, "m")
        self.median_win_size = 11
        self.min_disp = 1
        self.max_u_dist = 2.5
        f = get_length_units_conversion_factor("m", stereo_camera_model.units)
        self.min_range, self.max_range = 0.5 * f, 7.0 * f
        self.pano_correspondences = None
        self.num_valid_keypoints = 0
        self.establish_stereo_correspondences(omnistereo_model=stereo_camera_model, features=kwargs.get("features"))

    def establish_stereo_correspondences(self, omnistereo_model, collect_time_statistics=False, features=None):
        m = omnistereo_model
        if features is None:
            fm = m.feature_matcher_for_static_stereo
            kt, dt = m.top_model.detect_sparse_features_on_panorama(feature_detection_method=fm.feature_detection_method,
                                                                    num_of_features=fm.num_of_features,
                                                                    median_win_size=self.median_win_size, show=False)
            kb, db = m.bot_model.detect_sparse_features_on_panorama(feature_detection_method=fm.feature_detection_method,
                                                                    num_of_features=fm.num_of_features,
                                                                    median_win_size=self.median_win_size, show=False)
        else:
            kt, dt, kb, db = features
        (m_top, k_top, d_top), (m_bot, k_bot, d_bot), colors = m.match_features_panoramic_top_bottom(
            keypts_list_top=kt, desc_list_top=dt, keypts_list_bot=kb, desc_list_bot=db,
            min_rectified_disparity=self.min_disp, max_horizontal_diff=self.max_u_dist, show_matches=False)
        az1, el1 = m.top_model.panorama.get_direction_angles_from_pixel_pano(m_top, use_LUTs=False)
        az2, el2 = m.bot_model.panorama.get_direction_angles_from_pixel_pano(m_bot, use_LUTs=False)
        b_top = m.top_model.get_3D_point_from_angles_wrt_focus(azimuth=az1, elevation=el1)[0, ..., :3]
        b_bot = m.bot_model.get_3D_point_from_angles_wrt_focus(azimuth=az2, elevation=el2)[0, ..., :3]
        xyz = m.get_triangulated_point_from_direction_angles(dir_angs_top=(az1, el1), dir_angs_bot=(az2, el2),
                                                             use_midpoint_triangulation=self.use_midpoint_triangulation)[0]
        good = m.filter_panoramic_points_due_to_range(xyz, min_3D_range=self.min_range, max_3D_range=self.max_range)
        self.num_valid_keypoints = int(np.count_nonzero(good))
        self.bearing_vectors_top_stereo_triangulated = b_top[good]
        self.bearing_vectors_bottom_stereo_triangulated = b_bot[good]
        self.pano_correspondences = PanoramicCorrespondences(
            kpts_top_list=k_top[good], desc_top_list=d_top[good], kpts_bot_list=k_bot[good], desc_bot_list=d_bot[good],
            points_3D=xyz[good], m_top_array=m_top[good], m_bot_array=m_bot[good], random_colors_RGB_list=colors[good],
            do_flattening=False)


class StereoPanoramicKeyFrame(StereoPanoramicFrame):
    def __init__(self, frame):
        self.__dict__.update(frame.__dict__)


class TrackerSE3(object):
    """pose_est_tools.py:626-720."""

    def __init__(self, camera_model, show_3D_points=False, **kwargs):
        self.camera_model = camera_model
        self.show_3D_points = show_3D_points
        self.T_C_wrt_S_init = tr.identity_matrix()
        self.T_C_curr_frame_wrt_S_est = tr.identity_matrix()
        self.num_tracked_correspondences = 0
        self.inlier_tracked_correspondences_ratio = 0.
        self.number_of_cams = 1
        self.T_Ckey_wrt_S_est_list = []
        self.ransac_seed = kwargs.get("ransac_seed", 0)
        self.set_global_parameters_for_tracking()

    def set_global_parameters_for_tracking(self):
        self.backprojection_score_threshold_3D_to_2D_in_degrees = 5.
        self.backprojection_score_threshold_3D_to_2D = 1.0 - np.cos(np.deg2rad(self.backprojection_score_threshold_3D_to_2D_in_degrees))
        self.detection_method = "GFT"
        self.matching_type = "BF"
        self.k_best_matches = 1
        self.percentage_good_matches = 1.0
        self.num_features_detection_for_motion = 1000
        self.use_descriptor_radius_match_for_motion = False
        self.max_horizontal_search_ratio = 0.5
        self.pose_est_algorithm = "EPNP"
        self.n_points_for_RANSAC_model = 3
        self.correspondences_outliers_fraction = 0.65
        self.max_ransac_iterations_3D_to_2D = self.compute_num_of_iterations_RANSAC(
            n_points_for_model=self.n_points_for_RANSAC_model, correspondences_outliers_fraction=self.correspondences_outliers_fraction)

    def compute_num_of_iterations_RANSAC(self, n_points_for_model, correspondences_outliers_fraction):
        """pose_est_tools.py:709-720 -> 210 for the defaults."""
        w = 1.0 - correspondences_outliers_fraction
        num_of_iters = log10(1.0 - 0.998) / log10(1.0 - w ** n_points_for_model)
        std_of_k = sqrt(1.0 - w ** n_points_for_model) / (w ** n_points_for_model)
        return int(num_of_iters + 3 * std_of_k)


class TrackerStereoSE3(TrackerSE3):
    """pose_est_tools.py:722-878."""

    def __init__(self, camera_model, show_3D_points=False, **kwargs):
        TrackerSE3.__init__(self, camera_model, show_3D_points, **kwargs)
        self.omnistereo_model = self.camera_model
        self.number_of_cams = 2
        self.bootstrap_tracker()

    def bootstrap_tracker(self):
        m = self.omnistereo_model
        self.cam_offsets = np.array([m.top_model.T_model_wrt_C[:3, 3], m.bot_model.T_model_wrt_C[:3, 3]])
        self.cam_rotations = np.array([m.top_model.T_model_wrt_C[:3, :3], m.bot_model.T_model_wrt_C[:3, :3]])
        self.num_features_detection_for_static_stereo = 1000
        m.feature_matcher_for_static_stereo = FeatureMatcher(
            method=self.detection_method, matcher_type=self.matching_type, k_best=self.k_best_matches,
            percentage_good_matches=self.percentage_good_matches, num_of_features=self.num_features_detection_for_static_stereo)
        self.max_horizontal_diff_f2f_matches = 0.125 * self.max_horizontal_search_ratio * m.top_model.panorama.cols
        m.feature_matcher_for_motion = FeatureMatcher(
            method=self.detection_method, matcher_type=self.matching_type, k_best=self.k_best_matches,
            percentage_good_matches=self.percentage_good_matches, num_of_features=self.num_features_detection_for_motion)
        self.omni_mask_extra_padding = 10
        for mm in (m.top_model, m.bot_model):
            mm.panorama.generate_azimuthal_masks(azimuth_mask_degrees=30, overlap_degrees=0,
                                                 elev_mask_padding=self.omni_mask_extra_padding,
                                                 stand_masks_azimuth_coord_in_degrees_list=[50, 170, 290],
                                                 stand_masks_width_in_degrees=10)

    def track_frame(self, reference_frame, current_frame):
        self.num_tracked_correspondences = 0
        self.inlier_tracked_correspondences_ratio = 0.
        ref, cur = reference_frame.pano_correspondences, current_frame.pano_correspondences
        parts = []
        for cam_idx, (kpts, desc, m_attr, b_cur) in enumerate((
                ("kpts_top", "desc_top", "m_top", current_frame.bearing_vectors_top_stereo_triangulated),
                ("kpts_bot", "desc_bot", "m_bot", current_frame.bearing_vectors_bottom_stereo_triangulated))):
            (ti, _, _), (qi, _, _), _ = match_features_frame_to_frame(
                cam_model=self.omnistereo_model, train_kpts=getattr(ref, kpts), train_desc=getattr(ref, desc),
                query_kpts=getattr(cur, kpts), query_desc=getattr(cur, desc), random_colors_RGB=ref.random_colors_RGB,
                keypts_as_points_train=getattr(ref, m_attr), keypts_as_points_query=getattr(cur, m_attr),
                max_horizontal_diff=self.max_horizontal_diff_f2f_matches)
            ti, qi = np.asarray(ti, int), np.asarray(qi, int)
            parts.append((b_cur[qi], np.zeros((len(qi), 1)) + float(cam_idx), ref.points_3D_coords_homo[ti][:, :3],
                          cur.points_3D_coords_homo[qi][:, :3]))
        bearings = np.vstack([p[0] for p in parts])
        cam_idx_all = np.vstack([p[1] for p in parts])
        points_ref = np.vstack([p[2] for p in parts])
        points_cur = np.vstack([p[3] for p in parts])
        n = len(points_ref)
        if n < 2 * self.n_points_for_RANSAC_model * (0.33 * self.number_of_cams):
            return False, "Cannot track on only %d point correspondences" % (n)
        T, inliers = absolute_pose_noncentral_ransac(bearings, cam_idx_all, points_ref, self.cam_offsets, self.cam_rotations,
                                                     self.backprojection_score_threshold_3D_to_2D,
                                                     self.max_ransac_iterations_3D_to_2D, points_cur=points_cur,
                                                     seed=self.ransac_seed)
        if len(inliers) < 3 or not np.all(np.isfinite(T)):
            return False, "RANSAC found no model on %d point correspondences" % (n)
        self.indices_inliers_combined = inliers
        self.num_tracked_correspondences = len(inliers)
        self.inlier_tracked_correspondences_ratio = float(len(inliers)) / float(n)
        T_nl = absolute_pose_noncentral_optimize_nonlinear(bearings[inliers], cam_idx_all[inliers], points_ref[inliers],
                                                           self.cam_offsets, self.cam_rotations, T[:3, 3], T[:3, :3],
                                                           points_cur=points_cur[inliers])
        T_homo = np.identity(4)
        T_homo[:3] = T_nl
        T_homo[:3, 3] = T_nl[:3, 3] * current_frame.conversion_factor_length_to_m
        current_frame.T_frame_wrt_tracking_ref_frame = T_homo
        key = self.T_Ckey_wrt_S_est_list[-1] if self.T_Ckey_wrt_S_est_list else tr.identity_matrix()
        self.T_C_curr_frame_wrt_S_est = tr.concatenate_matrices(key, T_homo)
        return True, "tracking used %d inlier point correspondences" % (self.num_tracked_correspondences)


class RGBDFrame(object):
    """pose_est_tools.py:404-623; keypoints + descriptors come in through `features=(kpts, desc)` or cv2 ORB."""

    def __init__(self, rgbd_camera_model, rgb, depth, frame_id, **kwargs):
        self.frame_id = frame_id
        self.rgbd_camera_model = rgbd_camera_model
        self.rgb_img = rgb
        self.T_frame_wrt_tracking_ref_frame = np.identity(4)
        self.conversion_factor_length_to_m = get_length_units_conversion_factor(rgbd_camera_model.units, "m")
        f = get_length_units_conversion_factor("m", rgbd_camera_model.units)
        self.min_range, self.max_range = 0.8 * f, 7.0 * f
        self.establish_keypoints(rgb, depth, features=kwargs.get("features"))

    def establish_keypoints(self, rgb, depth, features=None):
        from .. import omnistereo as _o
        self.current_depth = depth
        if features is None:
            orb = cv2.ORB_create(nfeatures=self.rgbd_camera_model.feature_matcher_for_motion.num_of_features)
            gray = cv2.cvtColor(rgb, cv2.COLOR_BGR2GRAY) if rgb.ndim == 3 else rgb
            kpts, desc = orb.detectAndCompute(gray, None)
            kpts, desc = np.array(kpts), (desc if desc is not None else np.zeros((0, 32), np.uint8))
        else:
            kpts, desc = np.array(features[0]), np.asarray(features[1], np.uint8)
        all_kd = KeyPointAndDescriptor(kpts_list=kpts, desc_list=desc, do_flattening=False)
        u = all_kd.pixel_coords[..., 0].astype(np.uint).ravel().astype(np.int32)   # truncation, pose_est_tools.py:612
        v = all_kd.pixel_coords[..., 1].astype(np.uint).ravel().astype(np.int32)
        ctx = _o.device_context()
        cam = self.rgbd_camera_model
        xyz, bearing, valid = ctx.rgbd_backproject(cam.cam_vector(), _o.to_device(np.ascontiguousarray(depth, np.float32)),
                                                   _o.to_device(u), _o.to_device(v), self.min_range, self.max_range)
        valid = valid.cpu().numpy().astype(bool)
        self.keypoints_3D_points = xyz.cpu().numpy().astype(np.float64)[valid]
        self.bearing_vectors = bearing.cpu().numpy().astype(np.float64)[valid]
        self.keypoints_and_descriptors = KeyPointAndDescriptor(
            kpts_list=kpts[valid], desc_list=desc[valid], coords_array=all_kd.pixel_coords[0][valid],
            random_colors_RGB_list=all_kd.random_colors_RGB[valid], do_flattening=False)
        self.num_valid_keypoints = len(self.keypoints_3D_points)


class TrackerRGBDSE3(TrackerSE3):
    """pose_est_tools.py:880-958."""

    def __init__(self, camera_model, show_3D_points=False, **kwargs):
        TrackerSE3.__init__(self, camera_model, show_3D_points, **kwargs)
        self.number_of_cams = 1
        self.bootstrap_tracker()

    def bootstrap_tracker(self):
        self.camera_model.feature_matcher_for_motion = FeatureMatcher(
            method=self.detection_method, matcher_type=self.matching_type, k_best=self.k_best_matches,
            percentage_good_matches=self.percentage_good_matches, num_of_features=self.num_features_detection_for_motion)
        self.max_horizontal_diff_f2f_matches = self.max_horizontal_search_ratio * (self.camera_model.center_x * 2.)

    def track_frame(self, reference_frame, current_frame):
        self.num_tracked_correspondences = 0
        ref, cur = reference_frame.keypoints_and_descriptors, current_frame.keypoints_and_descriptors
        (ti, _, _), (qi, _, _), _ = match_features_frame_to_frame(
            cam_model=self.camera_model, train_kpts=ref.keypoints, train_desc=ref.descriptors, query_kpts=cur.keypoints,
            query_desc=cur.descriptors, random_colors_RGB=ref.random_colors_RGB, keypts_as_points_train=ref.pixel_coords,
            keypts_as_points_query=cur.pixel_coords, max_horizontal_diff=self.max_horizontal_diff_f2f_matches)
        ti, qi = np.asarray(ti, int), np.asarray(qi, int)
        n = len(ti)
        if n < 2 * self.n_points_for_RANSAC_model * self.number_of_cams:
            return False, "Cannot track on only %d point correspondences" % (n)
        bearings = current_frame.bearing_vectors[qi]
        p_ref = reference_frame.keypoints_3D_points[ti]
        p_cur = current_frame.keypoints_3D_points[qi]
        T, inliers = absolute_pose_ransac(bearings[..., :3], p_ref[..., :3], self.pose_est_algorithm,
                                          self.backprojection_score_threshold_3D_to_2D, self.max_ransac_iterations_3D_to_2D,
                                          points_cur=p_cur, seed=self.ransac_seed)
        if len(inliers) < 3 or not np.all(np.isfinite(T)):
            return False, "RANSAC found no model on %d point correspondences" % (n)
        self.num_tracked_correspondences = len(inliers)
        self.inlier_tracked_correspondences_ratio = float(len(inliers)) / float(n)
        T_nl = absolute_pose_optimize_nonlinear(bearings[inliers], p_ref[inliers], T[:3, 3], T[:3, :3], points_cur=p_cur[inliers])
        T_homo = np.identity(4)
        T_homo[:3] = T_nl
        T_homo[:3, 3] = T_nl[:3, 3] * current_frame.conversion_factor_length_to_m
        current_frame.T_frame_wrt_tracking_ref_frame = T_homo
        key = self.T_Ckey_wrt_S_est_list[-1] if self.T_Ckey_wrt_S_est_list else tr.identity_matrix()
        self.T_C_curr_frame_wrt_S_est = tr.concatenate_matrices(key, T_homo)
        return True, "tracking used %d inlier point correspondences" % (self.num_tracked_correspondences)


class RGBDKeyFrame(RGBDFrame):
    def __init__(self, frame):
        self.__dict__.update(frame.__dict__)


# ---------------------------------------------------------------------------------------------------------------------
# The VO loop of the demos (pose_est_tools.py:1264-1741): run_VO / driver_VO with the reference's signatures, headless.
# One frame per iteration, like the reference; the batched form of the same loop is vo_single_camera_sos_b200.driver.
# ---------------------------------------------------------------------------------------------------------------------
def run_VO(visualizer_3D_VO, camera_model, gt_poses_filename=None, est_poses_filename="estimated_frame_poses_TUM.txt",
           img_filename_template=None, depth_filename_template=None, img_indices=[], results_path="~/temp", thread_name=""):
    """pose_est_tools.py:1264-1678 without the 3D visualisation (visualizer_3D_VO must be None: vispy / matplotlib drawing
    is out of scope).  Writes, like the reference: <results>/estimated_frame_poses_TUM.txt, gt_associated_frame_poses_TUM.txt,
    keyframe_ids.txt, printed_messages.log.  Returns the VOResult of the run."""
    from os.path import expanduser, join, realpath
    from .camera_models import OmniStereoModel, RGBDCamModel
    from .common_cv import get_depthmap_float32_from_png, get_images
    from .common_tools import get_poses_from_file, make_sure_path_exists
    from ..driver import KeyframePolicy, TrackingState, tum_line
    if visualizer_3D_VO is not None:
        raise NotImplementedError("3D visualisation is not part of this build: pass visualizer_3D_VO=None")
    prefix = thread_name + ": " if len(thread_name) > 0 else ""
    is_sos = isinstance(camera_model, OmniStereoModel)
    tracker_class, keyframe_class = (TrackerStereoSE3, StereoPanoramicKeyFrame) if is_sos else (TrackerRGBDSE3, RGBDKeyFrame)
    results_path = realpath(expanduser(results_path))
    make_sure_path_exists(results_path)
    log = open(join(results_path, "printed_messages.log"), "w")
    image_names = get_images(img_filename_template, indices_list=img_indices, show_images=False, return_names_only=True)
    if not is_sos:
        depth_names = get_images(depth_filename_template, indices_list=img_indices, show_images=False, return_names_only=True)
    if img_indices is None or len(img_indices) == 0:
        img_indices = list(range(len(image_names)))
    tracker = tracker_class(camera_model=camera_model, show_3D_points=False, save_correspondence_images=False,
                            results_path=results_path)
    if gt_poses_filename is None:
        gt_T = {idx: np.identity(4) for idx in img_indices}                    # pose_est_tools.py:1332-1352
    else:
        _, mats = get_poses_from_file(poses_filename=gt_poses_filename, input_units="m", output_working_units="m", indices=[],
                                      pose_format="tum", zero_up_wrt_origin=True)
        gt_T = {idx: mats[idx] for idx in img_indices if idx < len(mats)}
        hand_eye = getattr(camera_model, "T_Cest_wrt_Rgt", None)
        if hand_eye is not None:                                                # pose_est_tools.py:1361-1368, 1466-1470
            T_Rgt_wrt_S = tracker.T_C_wrt_S_init @ np.linalg.inv(np.asarray(hand_eye, float))
            T_S_wrt_Rgt = np.linalg.inv(T_Rgt_wrt_S)
            gt_T = {k: T_Rgt_wrt_S @ v @ T_S_wrt_Rgt for k, v in gt_T.items()}
    est_file = open(join(results_path, est_poses_filename), "w")
    gt_file = open(join(results_path, est_poses_filename.replace("estimated", "gt_associated")), "w")
    key_file = open(join(results_path, "keyframe_ids.txt"), "w")
    state = TrackingState(KeyframePolicy(), tracker.number_of_cams, 1.0)        # T_frame_wrt_tracking_ref_frame is in [m] already
    reference_frame = None
    for n, idx in enumerate(img_indices):
        if is_sos:
            omni = cv2.imread(image_names[n])
            camera_model.set_current_omni_image(omni, generate_panoramas=False, view=False, apply_mask=True, mask_RGB=(0, 0, 0))
            frame = StereoPanoramicFrame(stereo_camera_model=camera_model, frame_id=idx,
                                         parent_id=state.keyframe_id if state.keyframe_id is not None else idx)
        else:
            rgb = cv2.cvtColor(cv2.imread(image_names[n]), cv2.COLOR_BGR2RGB)
            depth = get_depthmap_float32_from_png(depth_img_filename=depth_names[n], conversion_factor=camera_model.scaling_factor)
            frame = RGBDFrame(camera_model, rgb, depth, idx)
        if n == 0:
            state.first_frame(idx, frame.num_valid_keypoints)
            became_key = True
        else:
            ok, msg = tracker.track_frame(reference_frame=reference_frame, current_frame=frame)
            if not ok:                                                          # pose_est_tools.py:1493-1497
                print(msg, file=log)
                print("%sWarning failed: %s" % (prefix, msg))
                state.result.status = msg
                break
            T = frame.T_frame_wrt_tracking_ref_frame
            became_key = state.tracked_frame(idx, T[:3], tracker.num_tracked_correspondences, frame.num_valid_keypoints)
        if became_key:                                                          # pose_est_tools.py:1551-1566
            reference_frame = keyframe_class(frame=frame)
            tracker.T_Ckey_wrt_S_est_list.append(state.T_key_wrt_S[-1].copy())
            print(idx, file=key_file)
        T_est = state.T_curr_wrt_S
        print(tum_line(idx, T_est), file=est_file)                              # pose_est_tools.py:1609-1612
        T_gt = gt_T.get(idx, np.full((4, 4), np.nan))
        print(tum_line(idx, T_gt) if not np.any(np.isnan(T_gt)) else " ".join([str(idx)] + 7 * ["nan"]), file=gt_file)
        done = "%sDONE with F[%d] (Parent K[%d])" % (prefix, idx, state.result.parent_ids[-1])
        print(done, file=log)
    msg = "%sVO done with %d keyframes" % (prefix, len(state.result.keyframe_ids))
    print(msg)
    print(msg, file=log)
    for f in (est_file, gt_file, key_file, log):
        f.close()
    return state.result


def driver_VO(camera_model, scene_path, scene_path_vo_results, scene_img_filename_template, depth_filename_template,
              num_scene_images, visualize_VO=False, use_multithreads_for_VO=True, step_for_scene_images=1, first_image_index=0,
              last_image_index=-1, thread_name=""):
    """pose_est_tools.py:1680-1741.  Visualisation is not available here: visualize_VO=True only prints a notice and the VO
    runs headless; with use_multithreads_for_VO the loop runs on its own thread like the reference's."""
    import os
    import threading
    from datetime import datetime
    if use_multithreads_for_VO:
        est_poses_filename = "estimated_frame_poses_TUM.txt"
    else:
        now = datetime.now()
        est_poses_filename = "estimated_frame_poses_TUM-%d-%d-%d-%d-%d-%d.txt" % (now.year, now.month, now.day, now.hour, now.minute,
                                                                                 now.second)
    static = any(k in scene_path.lower() for k in ("static", "park", "grand")) or "GCT" in scene_path.upper() or "CCNY" in scene_path.upper()
    gt_poses_filename = None if static else os.path.join(scene_path, "gt_TUM.txt")
    if gt_poses_filename is not None and not os.path.exists(gt_poses_filename):
        gt_poses_filename = None   # the reference would stop here; a sequence without ground truth still runs
    last = min(last_image_index, num_scene_images) if last_image_index > 0 else num_scene_images
    indices = list(range(first_image_index, last, step_for_scene_images))
    if visualize_VO:
        print("%s: 3D visualisation is not part of this build, running headless" % thread_name)
    kwargs = dict(visualizer_3D_VO=None, camera_model=camera_model, gt_poses_filename=gt_poses_filename,
                  est_poses_filename=est_poses_filename, img_filename_template=scene_img_filename_template,
                  depth_filename_template=depth_filename_template, img_indices=indices, results_path=scene_path_vo_results,
                  thread_name=thread_name)
    if use_multithreads_for_VO:
        t = threading.Thread(target=run_VO, kwargs=kwargs)
        t.start()
        t.join()
    else:
        run_VO(**kwargs)
    print("%s Done with VO for %s!" % (thread_name, scene_path))
    return "NOTHING"
