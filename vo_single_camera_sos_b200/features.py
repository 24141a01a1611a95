"""Feature front on the device chained in front of the batched front-end: omni images in, poses out.

Per step: panoramas (sos_remap_u8) -> 11 x 11 median -> BGR2GRAY -> Shi-Tomasi corners per azimuthal mask and view ->
ORB description -> the bucketed feature arrays `Frontend.step` expects.  This is the device form of what
StereoPanoramicFrame does per frame with OpenCV (pose_est_tools.py:326-327 -> camera_models.py:1610-1797, default
detector "GFT"): same corners, same descriptors, frames and views batched.  Only index plumbing (bucket offsets,
compaction of the kept keypoints) is done with torch ops.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import ops
from .frontend import Frontend


def azimuthal_masks(rows: int, cols: int, n_buckets: int = 12, valid_rows: Optional[Sequence[bool]] = None) -> np.ndarray:
    """Column-range masks like Panorama.generate_azimuthal_masks (panorama.py:520-589) with overlap 0: bucket k covers the
    azimuth range [k, k+1) * 360 / n_buckets degrees; panorama columns run clockwise from azimuth 0 at the right edge
    (panorama.py:691-701).  `valid_rows` restricts the masks to the rows inside the view's elevation range."""
    masks = np.zeros((n_buckets, rows, cols), np.uint8)
    rr = np.ones(rows, bool) if valid_rows is None else np.asarray(valid_rows, bool)
    for k in range(n_buckets):
        # azimuth a maps to column cols - 1 - floor(a * cols / 2pi)
        c_hi = cols - 1 - int(np.floor(k * cols / n_buckets))
        c_lo = cols - 1 - int(np.floor((k + 1) * cols / n_buckets)) + 1 if k + 1 < n_buckets else 0
        masks[k][np.ix_(rr, np.arange(c_lo, c_hi + 1))] = 255
    return masks


class FeatureFront:
    """masks_top / masks_bot: uint8 [n_buckets, rows, cols] (pixel-disjoint masks run as one selection launch)."""

    def __init__(self, ctx: ops.Context, masks_top, masks_bot, corners_per_bucket: int, max_feat_per_view: int,
                 median: bool = True, quality_level: float = 0.01, min_distance: float = 5.0):
        self.ctx = ctx
        dev = lambda m: (m if isinstance(m, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(m))).to(ctx.device)
        self.masks = (dev(masks_top), dev(masks_bot))
        self.n_buckets = int(self.masks[0].shape[0])
        self.N = int(corners_per_bucket)
        self.F = int(max_feat_per_view)
        self.median, self.q, self.d = bool(median), float(quality_level), float(min_distance)

    def detect(self, pano: torch.Tensor):
        """pano uint8 [B, 2, rows, cols, 3] -> (px_top [B,F,2] f32, desc_top [B,F,32] u8, boff_top [B,n_buckets+1] i32,
        px_bot, desc_bot, boff_bot) in the layout of Frontend.step (features in bucket order, strongest first)."""
        ctx = self.ctx
        B, V, rows, cols, ch = pano.shape
        assert V == 2 and ch == 3
        imgs = pano.reshape(B * 2, rows, cols, 3)
        if self.median:
            imgs = ctx.median_blur_11(imgs)   # (median_blur_11_gray fuses the next call; measured 0.05 ms slower than the pair)
        gray = ctx.bgr_to_gray(imgs).reshape(B, 2, rows, cols)
        out = []
        for view in range(2):
            g = gray[:, view].contiguous()
            xy, cnt = ctx.gft_detect(g, self.masks[view], self.N, self.q, self.d)            # [B, nb, N, 2], [B, nb]
            nb, N = self.n_buckets, self.N
            pts = xy.reshape(B * nb * N, 2)
            img_idx = torch.arange(B, device=pts.device, dtype=torch.int32).repeat_interleave(nb * N)
            desc, keep = ctx.orb_describe(g, pts, None, img_idx)
            slot = torch.arange(N, device=pts.device)[None, None, :]
            keep = keep.reshape(B, nb, N).bool() & (slot < cnt[:, :, None])                   # detected AND described
            # bucket offsets and compaction in (bucket, strength) order
            per_bucket = keep.sum(-1)                                                           # [B, nb]
            boff = torch.zeros((B, nb + 1), dtype=torch.int64, device=pts.device)
            boff[:, 1:] = per_bucket.cumsum(1)
            boff = boff.clamp(max=self.F)
            flat = keep.reshape(B, nb * N)
            pos = flat.cumsum(1) - 1                                                            # position inside the frame
            ok = flat & (pos < self.F)
            b_idx = torch.arange(B, device=pts.device)[:, None].expand_as(pos)[ok]
            p_idx = pos[ok]
            px = torch.zeros((B, self.F, 2), dtype=torch.float32, device=pts.device)
            ds = torch.zeros((B, self.F, 32), dtype=torch.uint8, device=pts.device)
            px[b_idx, p_idx] = pts.reshape(B, nb * N, 2)[ok]
            ds[b_idx, p_idx] = desc.reshape(B, nb * N, 32)[ok]
            out += [px, ds, boff.to(torch.int32).contiguous()]
        return tuple(out)


def step_images(fe: Frontend, front: FeatureFront, omni: torch.Tensor, lut: torch.Tensor):
    """One front-end step from omni images alone: omni uint8 [B, H, W, 3] on the device -> results in fe.buffers().
    The panoramas are computed once here for the feature front and once more inside fe.step (3 % of the step)."""
    pano = fe.ctx.remap(omni, lut)
    feats = front.detect(pano)
    fe.step(omni, *feats)
    return feats
