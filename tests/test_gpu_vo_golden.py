"""GPU: the batched VO driver (front-end in keyframe mode + host policy) against the reference's OWN run_VO
(pose_est_tools.py:1264-1678; golden tests/golden/vo_sequence.npz from oracle/gen_vo_golden.py).  Fed the features the
reference detected, with the reference's argument list for the pose solver (bearings + reference 3D points: the bearing-only
three-point solver, then Levenberg-Marquardt), the device path must create the same keyframes, track the same number of
inliers per frame and reproduce estimated_frame_poses_TUM.txt to 1e-4."""
import numpy as np
import pytest
import torch

import _vo_golden

pytestmark = pytest.mark.gpu


def test_batched_vo_reproduces_the_reference_run(ctx):
    from vo_single_camera_sos_b200 import ops, synth
    from vo_single_camera_sos_b200.driver import BatchedVO, INPUT_KEYS, KeyframePolicy
    from vo_single_camera_sos_b200.frontend import FrontendConfig
    g, frames = _vo_golden.load()
    n = len(frames)
    W, H = int(g["width"]), int(g["height"])
    rig = synth.make_rig(W, H, int(g["pano_cols"]), seed=int(g["seed"]))
    pano = _vo_golden.pano_geometry(g)
    rows, cols = pano["rows"], pano["cols"]
    luts = []
    for which in ("top", "bot"):
        lo, hi = (float(x) for x in g[f"elev_{which}"])
        mx, my = ctx.lut_build(rig.gum_vector(which), rows, cols, pano["cyl_height_max"], pano["cyl_height_min"], lo, hi)
        luts.append(ctx.lut_pack(mx, my, (H, W), mask=torch.from_numpy(rig.mask(which)).to(ctx.device)))
    lut = torch.stack(luts).contiguous()
    F, B = 2048, 4
    cfg = FrontendConfig(batch=B, src_h=H, src_w=W, pano_rows=rows, pano_cols=cols, pano_top=g["pano_top"], pano_bot=g["pano_bot"],
                         f_top=rig.f_top, f_bot=rig.f_bot, max_feat_per_view=F, max_feat_per_bucket=512, cap=1024,
                         n_hyp=int(g["max_iterations"]), score_mode=ops.SCORE_BEARING, ransac_threshold=float(g["threshold"]),
                         refit=ops.REFINE_LM, refine_iters=60, keyframe_mode=True, solver=ops.SOLVER_P3P,
                         temporal_max_du=float(g["max_horizontal_diff_f2f"]), rig=_vo_golden.rig_matrix(g))
    hyp = torch.from_numpy(_vo_golden.hypothesis_list(g).view(np.int32)).to(ctx.device)
    per_frame = []
    for f in frames:
        d = dict(omni=np.zeros((H, W, 3), np.uint8))          # the panoramas do not enter the pose (features are inputs)
        for view in ("top", "bot"):
            k = len(f[f"px_{view}"])
            assert k <= F and np.diff(f[f"boff_{view}"]).max() <= 512
            px = np.zeros((F, 2), np.float32); px[:k] = f[f"px_{view}"]
            de = np.zeros((F, 32), np.uint8); de[:k] = f[f"desc_{view}"]
            d[f"px_{view}"], d[f"desc_{view}"], d[f"boff_{view}"] = px, de, f[f"boff_{view}"].astype(np.int32)
        per_frame.append({k: d[k] for k in INPUT_KEYS})
    vo = BatchedVO(ctx, cfg, lut, hyp, KeyframePolicy())
    res = vo.run(per_frame)
    assert res.status == "ok"
    assert res.keyframe_ids == g["keyframe_ids"].tolist()
    assert res.tracked[1:] == [len(g[f"f{i}_ransac_inliers"]) for i in range(1, n)]
    T_ref = [_vo_golden.tum_to_matrix(g["est_tum"][i]) for i in range(n)]
    parent = 0
    for i in range(1, n):
        rel_ref = np.linalg.inv(T_ref[parent]) @ T_ref[i]
        assert np.allclose(res.poses_wrt_keyframe[i], rel_ref, atol=1e-4), (i, np.abs(res.poses_wrt_keyframe[i] - rel_ref).max())
        assert np.allclose(res.poses_wrt_S[i], T_ref[i], atol=5e-4), i
        if i in res.keyframe_ids:
            parent = i
    vo.close()
