"""GPU: dense triangulation of panoramic disparity maps — kernel vs the reference's own output
(golden dense.npz: resolve_pano_correspondences_from_disparity_map + lifting + midpoint triangulation), fp32 tolerance
1e-4 relative; validity masks exact; full C2-size batch through size-independent properties."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import geometry

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def panos(g):
    pt = {k[9:]: float(g[k]) for k in g if k.startswith("pano_top_")}
    pb = {k[9:]: float(g[k]) for k in g if k.startswith("pano_bot_")}
    return pt, pb


def test_dense_triangulation_golden(ctx):
    g = load_golden("dense.npz")
    pt, pb = panos(g)
    for tag in ("all", "roi"):
        a = g[tag + "_args"]
        roi = None if a[2] < 0 else (int(a[2]), int(a[3]))
        xyz, valid = ctx.dense_triangulate(geometry.pano_vector(pt), geometry.pano_vector(pb), dev(g["disparity"]), g["f1"],
                                           g["f2"], a[0], a[1], float(g["lowest_reference_row"]), roi)
        xyz, valid = xyz.cpu().numpy(), valid.cpu().numpy().astype(bool)
        vt = valid.T
        uu, vv = np.nonzero(vt)
        assert np.array_equal(np.stack([uu, vv], 1), g[tag + "_top_px"])          # same pixels, same (u-major) order
        pts = xyz.transpose(1, 0, 2)[vt].astype(np.float64)
        want = g[tag + "_xyz"]
        err = np.linalg.norm(pts - want, axis=1) / np.linalg.norm(want, axis=1)
        assert err.max() < 1e-4, err.max()                                       # fp32 output vs float64 reference
        assert np.isnan(xyz[~valid]).all()


def test_dense_triangulation_mirror(ctx):
    """OmniStereoModel.triangulate_from_depth_map of the mirror returns the reference's point list."""
    from test_gpu_mirror import build_gums
    g = load_golden("dense.npz")
    gr = load_golden("remap.npz")
    gs = build_gums(gr)
    gs.set_current_omni_image(gr["img"], pano_width_in_pixels=200, generate_panoramas=True, view=False, apply_mask=True,
                              mask_RGB=(0, 0, 0))
    gs.disparity_map = g["disparity"]
    pts, coords = gs.triangulate_from_depth_map(min_disparity=2, max_disparity=9, roi_cols=(30, 150))
    assert np.array_equal(coords[0], g["roi_top_px"])
    assert np.allclose(pts[0, :, :3], g["roi_xyz"], rtol=1e-4, atol=1e-5) and np.all(pts[0, :, 3] == 1)


def test_dense_triangulation_c2_batch(ctx):
    """16 maps of 849 x 2400: validity equals the oracle's chain on a sample map; range consistent with the geometry."""
    from vo_single_camera_sos_b200 import synth
    rig = synth.make_rig(2048, 2048, 2400, seed=0)
    pg = dict(rig.pano)
    rows, cols = pg["rows"], pg["cols"]
    gen = torch.Generator(device="cuda").manual_seed(5)
    disp = torch.rand((16, rows, cols), device="cuda", generator=gen) * 40.0
    disp[disp < 6.0] = 0.0
    disp = torch.minimum(disp, torch.arange(rows, device="cuda", dtype=torch.float32)[None, :, None])
    xyz, valid = ctx.dense_triangulate(rig.pano_vector(), rig.pano_vector(), disp, rig.f_top, rig.f_bot, 1.0, 0.0, rows - 1.0)
    torch.cuda.synchronize()
    k = 7
    oxyz, ovalid = geometry.dense_triangulate(pg, pg, disp[k].cpu().numpy(), rig.f_top, rig.f_bot, 1.0, 0.0, rows - 1.0)
    assert np.array_equal(valid[k].cpu().numpy().astype(bool), ovalid)
    got = xyz[k].cpu().numpy().astype(np.float64)
    ok = ovalid & np.isfinite(oxyz).all(-1)
    err = np.linalg.norm(got[ok] - oxyz[ok], axis=1) / np.linalg.norm(oxyz[ok], axis=1)
    assert err.max() < 1e-4, err.max()
    assert torch.isnan(xyz[~valid.bool()]).all()
    # larger disparity = closer point (rectified vertical stereo): monotone along any valid column sample
    rng = np.linalg.norm(got, axis=-1)
    d = disp[k].cpu().numpy()
    sel = ok & (d > 0)
    r, c = np.nonzero(sel)
    pick = np.random.default_rng(0).choice(len(r), 2000, replace=False)
    same_row = {}
    for i in pick:
        same_row.setdefault(r[i], []).append((d[r[i], c[i]], rng[r[i], c[i]]))
    for lst in same_row.values():
        lst.sort()
        assert all(a[1] >= b[1] - 1e-3 * a[1] for a, b in zip(lst, lst[1:]))
