"""GPU parity of the batched front-end (sos_frontend_*): every stage of a step is compared with the oracle on the same
inputs — the stage's inputs are read back from the device, so each stage is checked in isolation at its own bar
(bit-exact for panoramas, match lists, compaction indices and RANSAC inlier sets; 1e-4 relative for float32 geometry)."""
import numpy as np
import pytest
import torch

from oracle import geometry, hamming, ransac, remap

pytestmark = pytest.mark.gpu


def host(b):
    return {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in b.items()}


def check_step(w, fr, buf, prev, mode):
    """fr: host inputs of this step; buf: host copies of the device buffers after the step; prev: store of slot 0 before
    the step (dict with n, xyz, b_top, b_bot, uv_top, uv_bot, desc) or None."""
    cfg = w.cfg
    B, cap, nb, F = cfg.batch, cfg.cap, cfg.n_buckets, cfg.max_feat_per_view
    pano_g = dict(zip(("cols", "rows", "pixel_size", "cyl_height_max", "cyl_circumference", "cyl_radius"), w.rig.pano_vector()))
    # ---- step 1: panoramas, bit-exact against cv2.remap on the masked image
    for b in range(B):
        for v, which in enumerate(("top", "bot")):
            mx, my = w.maps[which]
            ref = remap.remap_reference(remap.masked_image(fr["omni"][b], w.masks[which]), mx, my)
            assert np.array_equal(buf["pano"][b, v], ref), (b, which)
    n_store = buf["n"]
    for b in range(B):
        # ---- step 2a: stereo pairs per bucket (bit-exact)
        rq_all, rt_all = [], []
        for k in range(nb):
            q0, q1 = fr["boff_bot"][b, k], fr["boff_bot"][b, k + 1]
            t0, t1 = fr["boff_top"][b, k], fr["boff_top"][b, k + 1]
            qi, ti, dd = hamming.match_select(fr["desc_bot"][b, q0:q1], fr["desc_top"][b, t0:t1], "nn",
                                              px_q=fr["px_bot"][b, q0:q1], px_t=fr["px_top"][b, t0:t1], max_du=2.5, min_dv=1.0)
            s = b * nb + k
            cnt = buf["st_pair_count"][s]
            assert cnt == len(qi), (b, k)
            start = buf["st_q_start"][s]
            assert start == b * F + q0
            assert np.array_equal(buf["st_pair_q"][start:start + cnt], b * F + q0 + qi)
            assert np.array_equal(buf["st_pair_t"][start:start + cnt], b * F + t0 + ti)
            assert np.array_equal(buf["st_pair_d"][start:start + cnt], dd)
            rq_all.append(q0 + qi); rt_all.append(t0 + ti)
        rq, rt = np.concatenate(rq_all), np.concatenate(rt_all)
        # ---- steps 3+4: lifting, triangulation, range gate, ordered compaction
        az1, el1 = geometry.pano_pixel_to_angles(pano_g, fr["px_top"][b][rt])
        az2, el2 = geometry.pano_pixel_to_angles(pano_g, fr["px_bot"][b][rq])
        xyz = geometry.triangulate_midpoint(az1, el1, az2, el2, w.rig.f_top, w.rig.f_bot)
        homo = np.hstack([xyz, np.ones((len(xyz), 1))])
        keep = geometry.range_filter(homo, cfg.min_range, cfg.max_range)
        nrm = np.linalg.norm(homo, axis=1)
        decided = np.isnan(nrm) | ((np.abs(nrm - cfg.min_range) > 1e-9) & (np.abs(nrm - cfg.max_range) > 1e-9))
        assert decided.all()
        n = int(n_store[b + 1])
        assert n == min(int(keep.sum()), cap) and n > 0
        sl = slice(0, n)
        assert np.array_equal(buf["src_top"][b + 1, sl], b * F + rt[keep][:n])
        assert np.array_equal(buf["src_bot"][b + 1, sl], b * F + rq[keep][:n])
        assert np.array_equal(buf["uv_top"][b + 1, sl], fr["px_top"][b][rt[keep][:n]])
        assert np.array_equal(buf["uv_bot"][b + 1, sl], fr["px_bot"][b][rq[keep][:n]])
        ref_xyz = xyz[keep][:n]
        err = np.linalg.norm(buf["xyz"][b + 1, sl] - ref_xyz, axis=1) / np.linalg.norm(ref_xyz, axis=1)
        assert err.max() < 1e-4, err.max()
        assert np.allclose(buf["b_top"][b + 1, sl], geometry.angles_to_sphere(az1, el1)[keep][:n], rtol=1e-4, atol=1e-6)
        assert np.allclose(buf["b_bot"][b + 1, sl], geometry.angles_to_sphere(az2, el2)[keep][:n], rtol=1e-4, atol=1e-6)
        desc_top_c = buf["desc_c"][0, b + 1, :n].view(np.uint8).reshape(n, 32)
        desc_bot_c = buf["desc_c"][1, b + 1, :n].view(np.uint8).reshape(n, 32)
        assert np.array_equal(desc_top_c, fr["desc_top"][b][rt[keep][:n]])
        assert np.array_equal(desc_bot_c, fr["desc_bot"][b][rq[keep][:n]])
    # the store of the reference frames: slot 0 is what the previous step left behind; the step ends by copying slot B
    # to slot 0, so rebuild the pre-step slot 0 from `prev`
    store = {k: buf[k].copy() for k in ("xyz", "b_top", "b_bot", "uv_top", "uv_bot")}
    desc_store = buf["desc_c"].copy()
    n_pre = n_store.copy()
    if prev is None:
        n_pre[0] = 0
    else:
        n_pre[0] = prev["n"]
        for k in store:
            store[k][0] = prev[k]
        desc_store[:, 0] = prev["desc"]
    max_du = 0.125 * 0.5 * cfg.pano_cols
    thr = cfg.ransac_threshold
    rig = np.zeros((2, 3, 4)); rig[:, :, :3] = np.eye(3); rig[0, :, 3] = w.rig.f_top; rig[1, :, 3] = w.rig.f_bot
    for i in range(B):
        n_ref, n_cur = int(n_pre[i]), int(n_pre[i + 1])
        corr = []
        for view, uvk in ((0, "uv_top"), (1, "uv_bot")):
            qd = desc_store[view, i + 1, :n_cur].view(np.uint8).reshape(n_cur, 32)
            td = desc_store[view, i, :n_ref].view(np.uint8).reshape(n_ref, 32)
            qi, ti, dd = hamming.match_select(qd, td, "nn", px_q=store[uvk][i + 1, :n_cur], px_t=store[uvk][i, :n_ref],
                                              max_du=max_du, min_dv=-1.0)
            s = view * B + i
            assert buf["tm_pair_count"][s] == len(qi), (i, view)
            start = buf["tm_q_start"][s]
            base = view * (B + 1) * cap
            assert np.array_equal(buf["tm_pair_q"][start:start + len(qi)] - base, (i + 1) * cap + qi)
            assert np.array_equal(buf["tm_pair_t"][start:start + len(qi)] - base, i * cap + ti)
            corr.append((view, qi, ti))
        # ---- stacked correspondences (pose_est_tools.py:752-778)
        p_ref = np.concatenate([store["xyz"][i][ti] for _, _, ti in corr])
        p_cur = np.concatenate([store["xyz"][i + 1][qi] for _, qi, _ in corr])
        f_cur = np.concatenate([store["b_top" if v == 0 else "b_bot"][i + 1][qi] for v, qi, _ in corr])
        cam = np.concatenate([np.full(len(qi), v, np.uint8) for v, qi, _ in corr])
        m = len(p_ref)
        assert buf["n_corr"][i] == m and buf["n_corr_top"][i] == len(corr[0][1])
        assert np.array_equal(buf["p_ref"][i, :m], p_ref) and np.array_equal(buf["p_cur"][i, :m], p_cur)
        assert np.array_equal(buf["f_cur"][i, :m], f_cur) and np.array_equal(buf["cam"][i, :m], cam)
        # ---- step 5: RANSAC on exactly these float32 correspondences
        o = ransac.ransac_p3d(p_ref, p_cur, w.hyp_host, mode, thr, f_cur=f_cur, cam=cam, rig=rig)
        assert buf["best_hyp"][i] == o["best_hyp"], (i, buf["best_hyp"][i], o["best_hyp"])
        assert buf["best_count"][i] == o["best_count"]
        assert o["margin"] > 1e-9
        assert np.array_equal(buf["inlier_mask"][i, :m].astype(bool), o["mask"])
        assert not buf["inlier_mask"][i, m:].any()
        assert tuple(buf["stats"][i]) == (n_pre[i + 1], m, o["best_count"], o["best_hyp"])
        if o["best_hyp"] >= 0:
            assert np.allclose(buf["ransac_pose"][i], o["pose"], rtol=1e-4, atol=1e-5)
            assert np.allclose(buf["pose"][i], ransac.refit(p_ref, p_cur, o["mask"]), rtol=1e-4, atol=1e-5)
    # what the next step will find in slot 0
    return dict(n=int(n_store[B]), desc=buf["desc_c"][:, B].copy(),
                **{k: buf[k][B].copy() for k in ("xyz", "b_top", "b_bot", "uv_top", "uv_bot")})


@pytest.mark.parametrize("mode,graph", [("bearing", True), ("euclid", False)])
def test_frontend_two_steps_stage_by_stage(ctx, mode, graph):
    from vo_single_camera_sos_b200 import ops, workload
    B = 3
    w = workload.build(ctx, "tiny", batch=B, n_frames=2 * B, seed=3,
                       score_mode=ops.SCORE_BEARING if mode == "bearing" else ops.SCORE_EUCLID)
    fe = w.frontend(ctx)
    fe.set_graph(graph)
    prev = None
    rel = []
    for step in range(2):
        fr = workload.make_frames(w, step * B, B)
        fe.step(*workload.to_device(ctx, fr))
        torch.cuda.synchronize()
        buf = host(fe.buffers())
        # slot 0 now holds the carried frame == slot B (the copy at the end of the step)
        nB = int(buf["n"][B])
        assert buf["n"][0] == nB and np.array_equal(buf["xyz"][0, :nB], buf["xyz"][B, :nB])
        assert np.array_equal(buf["desc_c"][:, 0, :nB], buf["desc_c"][:, B, :nB])
        prev = check_step(w, fr, buf, prev, mode)
        rel.append((buf["pose"].copy(), buf["stats"].copy()))
    # sanity against the ground-truth motion (not a parity bar): the refit pose of pair i is frame i wrt frame i-1
    poses, stats = rel[1]
    for i in range(B):
        k = B + i
        T_rel = np.linalg.inv(w.trajectory[k - 1]) @ w.trajectory[k]
        assert stats[i, 2] > 20
        assert np.allclose(poses[i][:, :3], T_rel[:3, :3], atol=0.05)
        assert np.allclose(poses[i][:, 3], T_rel[:3, 3], atol=0.15)
    fe.close()


@pytest.mark.parametrize("upload", ["bands", "full"])
def test_frontend_host_api_matches_device_api(ctx, upload, monkeypatch):
    """Host-buffer API against the device API.  `bands` uploads only the LUT-reachable row spans of each omni image (the
    default), `full` the whole images (SOS_FULL_UPLOAD=1, read when the staging buffers are first built)."""
    from vo_single_camera_sos_b200 import workload
    if upload == "full":
        monkeypatch.setenv("SOS_FULL_UPLOAD", "1")
    else:
        monkeypatch.delenv("SOS_FULL_UPLOAD", raising=False)
    B = 2
    w = workload.build(ctx, "tiny", batch=B, n_frames=3 * B, seed=5)
    fe_dev, fe_host = w.frontend(ctx), w.frontend(ctx)
    frames = [workload.make_frames(w, s * B, B, render=(s == 0)) for s in range(3)]
    ref = []
    for fr in frames:
        fe_dev.step(*workload.to_device(ctx, fr))
        torch.cuda.synchronize()
        b = fe_dev.buffers()
        ref.append((b["pose"].cpu().numpy().copy(), b["stats"].cpu().numpy().copy()))
    pinned = [workload.to_pinned(fr) for fr in frames]
    # pipelined: submit two steps before waiting for the first
    t0 = fe_host.submit_host(*pinned[0])
    t1 = fe_host.submit_host(*pinned[1])
    out = [fe_host.wait_host(t0), fe_host.wait_host(t1)]
    out.append(fe_host.step_host(*pinned[2]))
    for (p, s), (rp, rs) in zip(out, ref):
        assert np.array_equal(s, rs)
        assert np.array_equal(p, rp, equal_nan=True)
    with pytest.raises(Exception):
        fe_host.wait_host(t0)  # ticket already consumed
    h2d, d2h = fe_host.host_bytes()
    full = B * w.cfg.src_h * w.cfg.src_w * 3
    assert (h2d > full) if upload == "full" else (h2d < full + B * w.cfg.max_feat_per_view * 2 * 40 + 4096)
    fe_dev.close(); fe_host.close()


def test_frontend_oversized_bucket_is_clamped_and_reported(ctx):
    """A bucket with more features than max_feat_per_bucket: the device path clamps the query side of that segment, counts
    what it dropped in buffers()["overflow"], and every other stage stays consistent with the clamped segment (no stale
    scratch is read: pair lists of the other buckets equal the oracle's); the host path refuses the offsets."""
    from vo_single_camera_sos_b200 import workload
    B = 2
    w = workload.build(ctx, "tiny", batch=B, n_frames=B, seed=8)
    fr = workload.make_frames(w, 0, B, render=False)
    cfg = w.cfg
    nb, F, cap_b = cfg.n_buckets, cfg.max_feat_per_view, cfg.max_feat_per_bucket
    # merge all buckets of frame 0 / bottom view into bucket 0 (and of the top view too, so that matches exist)
    total_bot, total_top = int(fr["boff_bot"][0, -1]), int(fr["boff_top"][0, -1])
    assert total_bot > cap_b
    fr["boff_bot"][0, 1:] = total_bot
    fr["boff_top"][0, 1:] = total_top
    fe = w.frontend(ctx)
    fe.step(*workload.to_device(ctx, fr))
    torch.cuda.synchronize()
    buf = host(fe.buffers())
    assert total_top > cap_b
    assert buf["overflow"][0] == (total_bot - cap_b) + (total_top - cap_b) and buf["overflow"][1] == 0
    assert buf["st_q_len"][0] == cap_b and buf["st_t_len"][0] == cap_b
    qi, ti, dd = hamming.match_select(fr["desc_bot"][0, :cap_b], fr["desc_top"][0, :cap_b], "nn",
                                      px_q=fr["px_bot"][0, :cap_b], px_t=fr["px_top"][0, :cap_b], max_du=2.5, min_dv=1.0)
    cnt = int(buf["st_pair_count"][0])
    assert cnt == len(qi)
    assert np.array_equal(buf["st_pair_q"][:cnt], qi) and np.array_equal(buf["st_pair_t"][:cnt], ti)
    assert not buf["st_pair_count"][1:nb].any()
    # frame 1 is untouched by frame 0's overflow
    for k in range(nb):
        q0, q1 = fr["boff_bot"][1, k], fr["boff_bot"][1, k + 1]
        t0, t1 = fr["boff_top"][1, k], fr["boff_top"][1, k + 1]
        qi, ti, dd = hamming.match_select(fr["desc_bot"][1, q0:q1], fr["desc_top"][1, t0:t1], "nn",
                                          px_q=fr["px_bot"][1, q0:q1], px_t=fr["px_top"][1, t0:t1], max_du=2.5, min_dv=1.0)
        s_ = nb + k
        assert buf["st_pair_count"][s_] == len(qi)
    assert buf["n"][1] <= cnt and buf["n"][2] > 0
    with pytest.raises(ValueError):
        fe.step_host(*workload.to_pinned(fr))
    fe.close()


def test_frontend_rejects_bad_inputs(ctx):
    from vo_single_camera_sos_b200 import workload
    w = workload.build(ctx, "tiny", batch=2, n_frames=2, seed=1)
    fe = w.frontend(ctx)
    fr = workload.make_frames(w, 0, 2, render=False)
    dev = workload.to_device(ctx, fr)
    with pytest.raises(ValueError):
        fe.step(dev[0][:1], *dev[1:])
    with pytest.raises(TypeError):
        fe.step(dev[0].cpu(), *dev[1:])
    fe.close()


def test_frontend_lm_refinement_matches_minpack(ctx):
    """SOS_REFINE_LM inside the captured chain: the output pose is the Levenberg-Marquardt minimiser of the bearing
    residual over the RANSAC inliers, started at the RANSAC pose (pose_est_tools.py:824-834) — checked against the
    MINPACK oracle on exactly the device's float32 correspondences (1e-6 absolute; float32 output)."""
    from vo_single_camera_sos_b200 import ops, workload
    B = 3
    w = workload.build(ctx, "tiny", batch=B, n_frames=2 * B, seed=4, score_mode=ops.SCORE_BEARING)
    w.cfg.refit = ops.REFINE_LM
    w.cfg.refine_iters = 60
    fe = w.frontend(ctx)
    rig = np.zeros((2, 3, 4)); rig[:, :, :3] = np.eye(3); rig[0, :, 3] = w.rig.f_top; rig[1, :, 3] = w.rig.f_bot
    for step in range(2):
        fr = workload.make_frames(w, step * B, B)
        fe.step(*workload.to_device(ctx, fr))
        torch.cuda.synchronize()
        buf = host(fe.buffers())
        for i in range(B):
            m = int(buf["n_corr"][i])
            if buf["best_hyp"][i] < 0:
                assert np.array_equal(buf["pose"][i], buf["ransac_pose"][i]) or np.isnan(buf["ransac_pose"][i]).any()
                continue
            mask = buf["inlier_mask"][i, :m].astype(bool)
            want, c0, c1 = ransac.refine_pose_lm(buf["p_ref"][i, :m], buf["f_cur"][i, :m], buf["ransac_pose"][i],
                                                 buf["cam"][i, :m], rig, mask)
            rs = buf["refine_stats"][i]
            assert rs[3] == mask.sum() and rs[2] <= 60
            assert rs[1] <= c1 * (1 + 1e-9) and rs[1] <= rs[0]
            np.testing.assert_allclose(buf["pose"][i], want, rtol=0, atol=1e-6)
    # sanity against the ground-truth motion
    for i in range(B):
        T_rel = np.linalg.inv(w.trajectory[B + i - 1]) @ w.trajectory[B + i]
        assert np.allclose(buf["pose"][i][:, :3], T_rel[:3, :3], atol=0.05)
        assert np.allclose(buf["pose"][i][:, 3], T_rel[:3, 3], atol=0.15)
    fe.close()


def test_frontend_bearing_only_solver_matches_oracle(ctx):
    """solver = SOLVER_P3P inside the captured chain: RANSAC from the bearings of the current frame and the 3D points of the
    reference frame only (what the reference hands to OpenGV, pose_est_tools.py:785) — winner, count and inlier set against
    oracle/p3p.py on exactly the device's float32 correspondences."""
    from oracle import p3p as op3p
    from vo_single_camera_sos_b200 import ops, workload
    B = 3
    w = workload.build(ctx, "tiny", batch=B, n_frames=2 * B, seed=4, score_mode=ops.SCORE_BEARING, solver=ops.SOLVER_P3P)
    w.cfg.refit = ops.REFINE_LM
    fe = w.frontend(ctx)
    rig = np.zeros((2, 3, 4)); rig[:, :, :3] = np.eye(3); rig[0, :, 3] = w.rig.f_top; rig[1, :, 3] = w.rig.f_bot
    for step in range(2):
        fr = workload.make_frames(w, step * B, B)
        fe.step(*workload.to_device(ctx, fr))
        torch.cuda.synchronize()
        buf = host(fe.buffers())
        for i in range(B):
            m = int(buf["n_corr"][i])
            if step == 0 and i == 0:
                continue                                                  # no reference frame yet
            M, h, c, inl, _ = op3p.ransac_p3p(buf["p_ref"][i, :m], buf["f_cur"][i, :m], buf["cam"][i, :m], rig, w.hyp_host,
                                              w.cfg.ransac_threshold)
            assert buf["best_hyp"][i] == h and buf["best_count"][i] == c and c > 0.5 * m
            np.testing.assert_array_equal(buf["inlier_mask"][i, :m].astype(bool), inl)
            np.testing.assert_allclose(buf["ransac_pose"][i], M, atol=2e-6)
    for i in range(B):
        T_rel = np.linalg.inv(w.trajectory[B + i - 1]) @ w.trajectory[B + i]
        assert np.allclose(buf["pose"][i][:, :3], T_rel[:3, :3], atol=0.05)
        assert np.allclose(buf["pose"][i][:, 3], T_rel[:3, 3], atol=0.15)
    fe.close()


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_frontend_full_size_stage_by_stage(ctx, name):
    """The same stage-by-stage comparison at the BASELINE sizes — config 1 (1280 x 960 omni, 1200-wide panoramas, 2000
    features per view, 210 hypotheses) and config 2 (2048 x 2048, 2400-wide, 8000 features, 4096 hypotheses): every stage of
    two captured steps against the oracle."""
    from vo_single_camera_sos_b200 import ops, workload
    B = 2
    w = workload.build(ctx, name, batch=B, n_frames=2 * B, seed=5, score_mode=ops.SCORE_BEARING)
    fe = w.frontend(ctx)
    prev = None
    for step in range(2):
        fr = workload.make_frames(w, step * B, B)
        fe.step(*workload.to_device(ctx, fr))
        torch.cuda.synchronize()
        buf = host(fe.buffers())
        prev = check_step(w, fr, buf, prev, "bearing")
    for i in range(B):
        T_rel = np.linalg.inv(w.trajectory[B + i - 1]) @ w.trajectory[B + i]
        assert buf["stats"][i, 2] > 500
        assert np.allclose(buf["pose"][i][:, :3], T_rel[:3, :3], atol=0.02)
        assert np.allclose(buf["pose"][i][:, 3], T_rel[:3, 3], atol=0.05)
    fe.close()
