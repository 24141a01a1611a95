"""Builds a complete synthetic workload for the batched front-end: rig -> device LUTs, frames -> input arrays.

Shared by bench.py and the parity tests so that both drive the front-end with the very same inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from . import ops, synth
from .frontend import Frontend, FrontendConfig

CONFIGS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case (demo_vo_sos.py path)
    "c1": dict(width=1280, height=960, pano_cols=1200, feat=2000, n_hyp=210, cap=2048, max_bucket=512),
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "c2": dict(width=2048, height=2048, pano_cols=2400, feat=8000, n_hyp=4096, cap=8192, max_bucket=2048),
    # small case for tests / smoke
    "tiny": dict(width=320, height=240, pano_cols=300, feat=300, n_hyp=64, cap=512, max_bucket=128),
}


@dataclass
class Workload:
    name: str
    rig: synth.Rig
    scene: synth.Scene
    cfg: FrontendConfig
    lut: torch.Tensor          # [2, rows, cols] int64 on the device
    hyp: torch.Tensor          # [n_hyp, 3] int32 on the device
    hyp_host: np.ndarray       # uint32
    masks: dict
    maps: dict                 # float64 LUT maps on the host (for the CPU baseline / oracle)
    trajectory: np.ndarray
    moving_fraction: float = 0.0   # share of moving landmarks (synth.make_frame_features `dynamic`): 0.79 -> ~35 % RANSAC inliers

    def frontend(self, ctx: ops.Context) -> Frontend:
        return Frontend(ctx, self.cfg, self.lut, self.hyp)


def build(ctx: ops.Context, name: str, batch: int, n_frames: int, seed: int = 0, score_mode: int = ops.SCORE_BEARING,
          n_hyp: int | None = None, solver: int = ops.SOLVER_ARUN, moving_fraction: float = 0.0) -> Workload:
    c = CONFIGS[name]
    rig = synth.make_rig(c["width"], c["height"], c["pano_cols"], seed=seed)
    scene = synth.make_scene(int(c["feat"] * 2.0), seed=seed)
    p = rig.pano
    rows, cols = p["rows"], p["cols"]
    luts, masks, maps = [], {}, {}
    for which in ("top", "bot"):
        lo, hi = rig.elev_top if which == "top" else rig.elev_bot
        mx, my = ctx.lut_build(rig.gum_vector(which), rows, cols, p["cyl_height_max"], p["cyl_height_min"], lo, hi)
        mask = rig.mask(which)
        masks[which] = mask
        maps[which] = (mx.cpu().numpy(), my.cpu().numpy())
        luts.append(ctx.lut_pack(mx, my, (rig.height, rig.width), mask=torch.from_numpy(mask).to(ctx.device)))
    lut = torch.stack(luts).contiguous()
    H = c["n_hyp"] if n_hyp is None else n_hyp
    hyp_host = np.random.default_rng(seed + 7).integers(0, 2 ** 32, (H, 4 if solver == ops.SOLVER_P3P else 3), dtype=np.uint64).astype(np.uint32)
    hyp = torch.from_numpy(hyp_host.view(np.int32)).to(ctx.device)
    thr = 1.0 - math.cos(math.radians(5.0)) if score_mode == ops.SCORE_BEARING else 0.05
    cfg = FrontendConfig(batch=batch, src_h=rig.height, src_w=rig.width, pano_rows=rows, pano_cols=cols,
                         pano_top=rig.pano_vector(), pano_bot=rig.pano_vector(), f_top=rig.f_top, f_bot=rig.f_bot,
                         max_feat_per_view=c["cap"], max_feat_per_bucket=c["max_bucket"], cap=c["cap"], n_hyp=H,
                         score_mode=score_mode, ransac_threshold=thr, solver=solver)
    traj = synth.make_trajectory(n_frames, seed=seed)
    return Workload(name, rig, scene, cfg, lut, hyp, hyp_host, masks, maps, traj, float(moving_fraction))


def make_frames(w: Workload, first: int, count: int, render: bool = True, lift=None, renderer: "DeviceRenderer" = None):
    """Host input arrays for frames [first, first+count): dict of omni [n,H,W,3], px/desc/boff per view, landmark ids.
    With `renderer` the omni images are rendered on the device (same scene, same geometry, much faster)."""
    c = CONFIGS[w.name]
    cfg = w.cfg
    n = count
    out = dict(
        omni=np.zeros((n, cfg.src_h, cfg.src_w, 3), np.uint8),
        px_top=np.zeros((n, cfg.max_feat_per_view, 2), np.float32), px_bot=np.zeros((n, cfg.max_feat_per_view, 2), np.float32),
        desc_top=np.zeros((n, cfg.max_feat_per_view, 32), np.uint8), desc_bot=np.zeros((n, cfg.max_feat_per_view, 32), np.uint8),
        boff_top=np.zeros((n, cfg.n_buckets + 1), np.int32), boff_bot=np.zeros((n, cfg.n_buckets + 1), np.int32),
        lid_top=np.zeros((n, cfg.max_feat_per_view), np.int64), lid_bot=np.zeros((n, cfg.max_feat_per_view), np.int64))
    for i in range(n):
        T = w.trajectory[first + i]
        f = synth.make_frame_features(w.rig, w.scene, T, c["feat"], cfg.n_buckets, seed=1000 * (first + i) + 17,
                                      cap=cfg.max_feat_per_view, dynamic=w.moving_fraction)
        for which in ("top", "bot"):
            out[f"px_{which}"][i] = f[which]["px"]
            out[f"desc_{which}"][i] = f[which]["desc"]
            out[f"boff_{which}"][i] = f[which]["bucket_off"]
            out[f"lid_{which}"][i] = f[which]["landmark"]
        if renderer is not None:
            out["omni"][i] = renderer.render(T).cpu().numpy()
        elif render:
            out["omni"][i] = synth.render_omni(w.rig, w.scene, T, lift=lift)
    return out


def device_lift(ctx: ops.Context, w: Workload):
    """GUM lifting callback for synth.render_omni running on the device (sos_lift_gum)."""
    def lift(which, uv):
        sphere, _, _ = ctx.lift_gum(w.rig.gum_vector(which), torch.from_numpy(np.ascontiguousarray(uv)).to(ctx.device))
        return sphere.cpu().numpy()
    return lift


class DeviceRenderer:
    """synth.render_omni on the device: the per-pixel viewing rays come from sos_lift_gum once per view, the ray/box
    intersection and the texture lookup are plain torch ops (input generation only, not part of the hot path)."""

    def __init__(self, ctx: ops.Context, w: Workload):
        self.ctx, self.w = ctx, w
        rig = w.rig
        self.tex = torch.from_numpy(w.scene.texture).to(ctx.device)
        self.half = torch.from_numpy(w.scene.half_extent).to(ctx.device)
        self.views = []
        for which in ("top", "bot"):
            m = torch.from_numpy(rig.mask(which) != 0).to(ctx.device)
            idx = m.nonzero()  # (y, x)
            uv = torch.stack([idx[:, 1], idx[:, 0]], 1).to(torch.float64).contiguous()
            rays, _, _ = ctx.lift_gum(rig.gum_vector(which), uv)
            f = torch.from_numpy(rig.f_top if which == "top" else rig.f_bot).to(ctx.device)
            self.views.append((idx, rays, f))

    def render(self, T_c_wrt_w: np.ndarray) -> torch.Tensor:
        rig = self.w.rig
        img = torch.zeros((rig.height, rig.width, 3), dtype=torch.uint8, device=self.ctx.device)
        R = torch.from_numpy(T_c_wrt_w[:3, :3].copy()).to(self.ctx.device)
        t = torch.from_numpy(T_c_wrt_w[:3, 3].copy()).to(self.ctx.device)
        for idx, rays, f in self.views:
            o = R @ f + t
            d = rays @ R.T
            tt = torch.where(d > 0, (self.half - o) / d, (-self.half - o) / d)
            tmin, k = tt.min(dim=1)
            hit = o + d * tmin[:, None]
            ar = torch.arange(len(d), device=d.device)
            face = 2 * k + (d[ar, k] > 0).long()
            T = self.tex.shape[1]
            iu = torch.remainder(torch.floor(hit[ar, (k + 1) % 3] / 0.12).long(), T)
            iv = torch.remainder(torch.floor(hit[ar, (k + 2) % 3] / 0.12).long(), T)
            img[idx[:, 0], idx[:, 1]] = self.tex[face, iu, iv]
        return img


INPUT_KEYS = ("omni", "px_top", "desc_top", "boff_top", "px_bot", "desc_bot", "boff_bot")


def to_device(ctx: ops.Context, frames: dict) -> list:
    return [torch.from_numpy(frames[k]).to(ctx.device) for k in INPUT_KEYS]


def to_pinned(frames: dict) -> list:
    return [torch.from_numpy(frames[k]).pin_memory() for k in INPUT_KEYS]
