"""GPU: the demo entry point (demo_vo_sos.py -> omnistereo.pose_est_tools.driver_VO -> run_VO, pose_est_tools.py:1264-1741)
on top of the mirrored classes: a pickled GUMStereo, a folder of omni images and a TUM ground-truth file in, the reference's
result files out (estimated / gt-associated TUM poses, keyframe ids, message log)."""
import os

import cv2
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def mirrored_gums_from_rig(rig, pano_cols, first_image):
    from omnistereo.gum import GUM, GUMStereo
    models = []
    for g, z_axis, F in ((rig.gum_top, 1.0, rig.f_top), (rig.gum_bot, -1.0, rig.f_bot)):
        m = GUM(precalib_filename="/nonexistent", new_method=True, z_axis=z_axis, image_size_pixels=(rig.width, rig.height),
                center_uv_point=(g["u_center"], g["v_center"]))
        m.precalib_params.set_gum_params(xi1=g["xi1"], xi2=g["xi2"], xi3=g["xi3"])
        m.precalib_params.set_generalized_cam_params(gamma1=g["gamma1"], gamma2=g["gamma2"], alpha_c=g["alpha_c"],
                                                     u_center=g["u_center"], v_center=g["v_center"])
        for k in ("k1", "k2", "k3", "p1", "p2", "l1", "l2", "l3"):
            setattr(m.precalib_params, k, g[k])
        m.set_model_params()
        m.units = "m"
        m.set_pose(np.asarray(F, float), np.eye(3))
        models.append(m)
    c = models[0].precalib_params.center_point
    gs = GUMStereo(models[0], models[1], center_point_top=c, center_point_top_inner=c, center_point_top_outer=c,
                   center_point_bottom=c, center_point_bottom_inner=c, center_point_bottom_outer=c,
                   radius_top_outer=rig.radii_top[1], radius_top_inner=rig.radii_top[0], radius_bottom_outer=rig.radii_bot[1],
                   radius_bottom_inner=rig.radii_bot[0])
    gs.set_current_omni_image(first_image, pano_width_in_pixels=pano_cols, generate_panoramas=True, view=False, apply_mask=True,
                              mask_RGB=(0, 0, 0))
    return gs


def test_driver_vo_runs_like_the_demo(ctx, tmp_path):
    from omnistereo.common_tools import load_obj_from_pickle, make_sure_path_exists, save_obj_in_pickle
    from omnistereo.pose_est_tools import driver_VO
    from vo_single_camera_sos_b200 import synth
    from vo_single_camera_sos_b200.driver import quaternion_wxyz
    n, cols = 5, 800
    rig = synth.make_rig(640, 480, cols, seed=3)
    scene = synth.make_scene(500, seed=3)
    traj = synth.make_trajectory(n, seed=3)
    scene_path = str(tmp_path / "lab_sequence")
    omni_dir = os.path.join(scene_path, "omni")
    make_sure_path_exists(omni_dir)
    images = [synth.render_omni(rig, scene, traj[i]) for i in range(n)]
    for i, img in enumerate(images):
        cv2.imwrite(os.path.join(omni_dir, "image-%06d.png" % i), img)
    with open(os.path.join(omni_dir, "gt_TUM.txt"), "w") as f:     # driver_VO looks for it next to the images (:1704)
        for i in range(n):
            T = traj[i]
            q = quaternion_wxyz(T)
            print(i, T[0, 3], T[1, 3], T[2, 3], q[1], q[2], q[3], q[0], file=f)
    pkl = str(tmp_path / "gums-calibrated.pkl")
    save_obj_in_pickle(mirrored_gums_from_rig(rig, cols, images[0]), pkl)
    gums = load_obj_from_pickle(pkl)                               # demo_vo_sos.py:109
    results = os.path.join(scene_path, "results-omni")
    out = driver_VO(camera_model=gums, scene_path=omni_dir, scene_path_vo_results=results,
                    scene_img_filename_template=os.path.join(omni_dir, "image-*.png"), depth_filename_template=None,
                    num_scene_images=n, visualize_VO=False, use_multithreads_for_VO=True, step_for_scene_images=1,
                    first_image_index=0, last_image_index=-1, thread_name="lab_sequence-SOS")
    assert out == "NOTHING"
    est = np.loadtxt(os.path.join(results, "estimated_frame_poses_TUM.txt"), ndmin=2)
    gt = np.loadtxt(os.path.join(results, "gt_associated_frame_poses_TUM.txt"), ndmin=2)
    keys = np.loadtxt(os.path.join(results, "keyframe_ids.txt"), ndmin=1)
    assert est.shape == (n, 8) and gt.shape == (n, 8) and list(est[:, 0]) == list(range(n))
    assert keys[0] == 0 and os.path.getsize(os.path.join(results, "printed_messages.log")) > 0
    # estimated trajectory (wrt the first frame) follows the ground truth, which run_VO zeroes up wrt its first pose
    T0_inv = np.linalg.inv(traj[0])
    for i in range(n):
        t_gt = (T0_inv @ traj[i])[:3, 3]
        assert np.allclose(gt[i, 1:4], t_gt, atol=1e-9)
        assert np.linalg.norm(est[i, 1:4] - t_gt) < 0.06, (i, est[i, 1:4], t_gt)
        assert abs(np.linalg.norm(est[i, 4:8]) - 1.0) < 1e-9


def test_driver_vo_rgbd_like_the_demo(ctx, tmp_path):
    """demo_vo_rgbd.py's call sequence (real-data branch: fx = fy = 525, depth_is_Z, 16-bit depth PNGs in mm): a textured
    plane seen by a moving pinhole camera; RGB images by homography, depth maps from the plane equation."""
    from omnistereo.camera_models import RGBDCamModel
    from omnistereo.common_tools import make_sure_path_exists
    from omnistereo.pose_est_tools import driver_VO
    from omnistereo.transformations import rotation_matrix
    rng = np.random.default_rng(7)
    W, H, fx, cx, cy = 640, 480, 525.0, 319.5, 239.5
    K = np.array([[fx, 0, cx], [0, fx, cy], [0, 0, 1.0]])
    tex = cv2.GaussianBlur(cv2.resize(rng.integers(0, 256, (60, 80, 3), dtype=np.uint8), (1600, 1200), interpolation=cv2.INTER_NEAREST),
                           (0, 0), 1.2)
    # world plane z = 3 m with texture coordinates (X, Y) = ((u - 800) / 250, (v - 600) / 250) [m]
    n_w, d_w = np.array([0.0, 0.0, 1.0]), 3.0
    A = np.array([[1 / 250.0, 0, -800 / 250.0], [0, 1 / 250.0, -600 / 250.0], [0, 0, 1.0]])   # texture px -> (X, Y, 1)
    poses = [np.eye(4)]
    for i in range(1, 5):
        T = rotation_matrix(0.01 * i, [0.2, 1.0, 0.1])
        T[:3, 3] = [0.03 * i, -0.01 * i, 0.02 * i]
        poses.append(T)                                          # camera wrt world
    scene_path = str(tmp_path / "rgbd_lab")
    rgb_dir, depth_dir = os.path.join(scene_path, "rgbd", "rgb"), os.path.join(scene_path, "rgbd", "depth")
    make_sure_path_exists(rgb_dir); make_sure_path_exists(depth_dir)
    yy, xx = np.mgrid[:H, :W]
    rays = np.stack([(xx - cx) / fx, (yy - cy) / fx, np.ones_like(xx, float)], -1)
    for i, T in enumerate(poses):
        R, t = T[:3, :3], T[:3, 3]
        # a texture pixel -> world point (X, Y, 3) -> camera -> image:  x ~ K R^T ([X, Y, 3] - t)
        Mw = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 0.0]]) @ A + np.outer([0, 0, d_w], [0, 0, 1.0])   # (u, v, 1) -> (X, Y, 3)
        Hm = K @ R.T @ (Mw - np.outer(t, [0, 0, 1.0]))
        cv2.imwrite(os.path.join(rgb_dir, "%04d.png" % i), cv2.warpPerspective(tex, Hm, (W, H), flags=cv2.INTER_LINEAR))
        # depth: ray r in the camera, world point t + s R r on the plane n.X = d  ->  s = (d - n.t) / (n.R r); Z = s (r_z = 1)
        s = (d_w - n_w @ t) / (rays @ (R.T @ n_w))
        cv2.imwrite(os.path.join(depth_dir, "%04d.png" % i), np.clip(np.round(s * 1000.0), 0, 65535).astype(np.uint16))
    with open(os.path.join(scene_path, "rgbd", "rgb", "gt_TUM.txt"), "w") as f:
        from vo_single_camera_sos_b200.driver import quaternion_wxyz
        for i, T in enumerate(poses):
            q = quaternion_wxyz(T)
            print(i, T[0, 3], T[1, 3], T[2, 3], q[1], q[2], q[3], q[0], file=f)
    cam = RGBDCamModel(fx=fx, fy=fx, center_x=cx, center_y=cy, scaling_factor=1. / 1000.0, do_undistortion=False, depth_is_Z=True,
                       focal_length_m=1. / 1000.0)
    cam.T_Cest_wrt_Rgt = None
    results = os.path.join(scene_path, "results-rgbd")
    driver_VO(camera_model=cam, scene_path=rgb_dir, scene_path_vo_results=results,
              scene_img_filename_template=os.path.join(rgb_dir, "*.png"), depth_filename_template=os.path.join(depth_dir, "*.png"),
              num_scene_images=len(poses), visualize_VO=False, use_multithreads_for_VO=False, thread_name="rgbd_lab-RGB-D")
    est_name = [n for n in os.listdir(results) if n.startswith("estimated_frame_poses_TUM")][0]
    est = np.loadtxt(os.path.join(results, est_name), ndmin=2)
    assert est.shape == (len(poses), 8)
    # The reference's objective (sum of squared 1 - cos over everything inside the 5 degree RANSAC gate, pose_est_tools.py:830)
    # is dominated by the few-pixel localisation errors of coarse-level ORB keypoints, and on a fronto-parallel plane a rotation
    # about y trades against a translation along x: the scipy LM restatement on cv2's own matches of this pair lands 0.15 m off
    # in t.  What the data pin down well is where the plane centre falls in the camera, so its lateral position is the tight check.
    from vo_single_camera_sos_b200.omnistereo.common_tools import _transform_from_tum
    X = np.array([0.0, 0.0, d_w])
    lateral, rng_err = [], []
    for i, T in enumerate(poses):
        Te = _transform_from_tum(est[i, 1:8])
        pe, pt = Te[:3, :3].T @ (X - Te[:3, 3]), T[:3, :3].T @ (X - T[:3, 3])
        lateral.append(float(np.linalg.norm((pe - pt)[:2])))
        rng_err.append(float(abs(pe[2] - pt[2])))                  # range is the weak direction
    assert lateral[0] == 0 and max(lateral) < 0.08 and max(rng_err) < 0.25, (lateral, rng_err, est[:, 1:4])
