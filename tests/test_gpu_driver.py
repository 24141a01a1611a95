"""GPU: the batched keyframe-mode VO driver against the sequential oracle loop on the same synthetic
frames — identical keyframe sequence, poses within the fp32 pose tolerance; plus the keyframe-mode primitives
(reference slots, promote, retrack) against the consecutive-mode chain."""
import io

import numpy as np
import pytest
import torch

from oracle import driver as odriver

pytestmark = pytest.mark.gpu


def host(buf):
    return {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in buf.items()}


def per_frame(fr, n):
    from vo_single_camera_sos_b200.driver import INPUT_KEYS
    return [{k: fr[k][i] for k in INPUT_KEYS} for i in range(n)]


def trimmed(fr, i):
    nt, nb = int(fr["boff_top"][i][-1]), int(fr["boff_bot"][i][-1])
    return dict(px_top=fr["px_top"][i][:nt], desc_top=fr["desc_top"][i][:nt], boff_top=fr["boff_top"][i],
                px_bot=fr["px_bot"][i][:nb], desc_bot=fr["desc_bot"][i][:nb], boff_bot=fr["boff_bot"][i])


@pytest.mark.parametrize("refine,pos_min", [("arun", 0.06), ("lm", 0.06), ("arun", 0.004)])
def test_batched_vo_matches_sequential_oracle(ctx, refine, pos_min):
    """pos_min = 0.06: a keyframe every 2-3 frames (pairs predicted against the keyframe slot); pos_min = 0.004: every frame
    becomes a keyframe, so from the second batch on the driver predicts the chain (previous frame = keyframe)."""
    from vo_single_camera_sos_b200 import ops, workload
    from vo_single_camera_sos_b200.driver import BatchedVO, KeyframePolicy, tum_line
    B, n_frames = 4, 15
    w = workload.build(ctx, "tiny", batch=B, n_frames=n_frames, seed=6, score_mode=ops.SCORE_BEARING)
    w.cfg.keyframe_mode = True
    w.cfg.refit = ops.REFINE_LM if refine == "lm" else ops.REFINE_ARUN
    w.cfg.refine_iters = 60
    fr = workload.make_frames(w, 0, n_frames, render=False)
    th = dict(odriver.INDOOR, pos_min=pos_min)   # the synthetic trajectory moves 1-5 cm per frame
    rig = np.zeros((2, 3, 4)); rig[:, :, :3] = np.eye(3); rig[0, :, 3] = w.rig.f_top; rig[1, :, 3] = w.rig.f_bot
    want = odriver.run_vo([trimmed(fr, i) for i in range(n_frames)],
                          (dict(w.rig.pano), w.rig.f_top, w.rig.f_bot, w.cfg.cap),
                          (w.hyp_host, "bearing", w.cfg.ransac_threshold, rig, 0.125 * 0.5 * w.cfg.pano_cols),
                          thresholds=th, refine=refine)
    assert want["status"] == "ok" and 3 <= len(want["keyframe_ids"]) <= n_frames
    assert (len(want["keyframe_ids"]) < n_frames) == (pos_min > 0.01)
    # the decisions the comparison relies on are not within rounding of a threshold
    for dist, ang, *_ in want["decisions"]:
        assert abs(dist - th["pos_min"]) > 1e-3 and abs(dist - th["pos_max"]) > 1e-3 and abs(ang - th["ang_max"]) > 1e-3

    vo = BatchedVO(ctx, w.cfg, w.lut, w.hyp, KeyframePolicy(**th))
    est, keys = io.StringIO(), io.StringIO()
    res = vo.run(per_frame(fr, n_frames), est_poses_file=est, keyframe_ids_file=keys)
    assert res.status == "ok"
    assert res.keyframe_ids == want["keyframe_ids"]
    assert res.frame_ids == list(range(n_frames))
    # fp32 device pipeline vs float64 oracle: 1e-4 relative on rotation entries, 1e-4 * scene scale (7 m) on translations
    got_rel, want_rel = np.array(res.poses_wrt_keyframe), np.array(want["poses_wrt_keyframe"])
    assert np.allclose(got_rel[:, :3, :3], want_rel[:, :3, :3], atol=1e-4)
    assert np.allclose(got_rel[:, :3, 3], want_rel[:, :3, 3], atol=7e-4)
    assert np.allclose(np.array(res.poses_wrt_S), np.array(want["poses_wrt_S"]), atol=2e-3)
    assert res.tracked == want["tracked"]
    # every frame was resolved with at most one device step per batch + one stage-B re-run per in-batch keyframe
    assert res.device_steps == -(-n_frames // B) and res.device_retracks <= len(res.keyframe_ids)
    if pos_min < 0.01:                               # chain prediction: only the first batch needs re-runs
        assert res.device_retracks <= B
    # files: one TUM line per frame, one id per keyframe (pose_est_tools.py:1557, 1609-1612)
    lines = est.getvalue().strip().split("\n")
    assert len(lines) == n_frames and lines[3] == tum_line(3, res.poses_wrt_S[3])
    assert [int(x) for x in keys.getvalue().split()] == res.keyframe_ids
    # sanity against the ground-truth trajectory (not a parity bar)
    T_gt = np.linalg.inv(w.trajectory[0]) @ w.trajectory[n_frames - 1]
    assert np.allclose(res.poses_wrt_S[-1][:3, 3], T_gt[:3, 3], atol=0.25)
    vo.close()


def test_keyframe_primitives_against_consecutive_chain(ctx):
    """ref_slot[i] = i reproduces the consecutive chain bit for bit; promote + retrack re-tracks later frames against a
    promoted slot exactly as a fresh step would."""
    from vo_single_camera_sos_b200 import ops, workload
    B = 3
    w = workload.build(ctx, "tiny", batch=B, n_frames=B, seed=9, score_mode=ops.SCORE_EUCLID)
    fr = workload.make_frames(w, 0, B, render=False)
    dev = workload.to_device(ctx, fr)
    fe = w.frontend(ctx)
    fe.step(*dev)
    torch.cuda.synchronize()
    a = host(fe.buffers())
    fe.close()
    w.cfg.keyframe_mode = True
    fk = w.frontend(ctx)
    fk.set_ref_slots([0, 1, 2])
    fk.step(*dev)
    torch.cuda.synchronize()
    b = host(fk.buffers())
    for k in ("pose", "ransac_pose", "best_hyp", "best_count", "inlier_mask", "n_corr", "tm_pair_count"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert b["n"][0] == 0                      # no automatic carry-over in keyframe mode
    # pair 2 against slot 1 (frame 0) directly, and through promote(1) + slot 0
    fk.set_ref_slots([-1, -1, 1])
    fk.retrack()
    torch.cuda.synchronize()
    c = host(fk.buffers())
    assert c["best_hyp"][0] == -1 and c["n_corr"][0] == 0 and c["n_corr"][2] > 0
    fk.promote(1)
    fk.set_ref_slots([-1, -1, 0])
    fk.retrack()
    torch.cuda.synchronize()
    d = host(fk.buffers())
    assert d["n"][0] == d["n"][1]
    for k in ("pose", "best_hyp", "best_count", "n_corr"):
        assert np.array_equal(c[k][2], d[k][2]), k
    assert np.array_equal(c["inlier_mask"][2], d["inlier_mask"][2])
    with pytest.raises(Exception):
        fk.set_ref_slots([1, 0, 0])            # a frame cannot track against itself
    fk.close()
