#!/usr/bin/env python
"""Image-in throughput of the whole front-end on one GPU (SURVEY §8f N3 + the hot path): rendered C2 omni images resident
in HBM -> panoramas, features, matching, triangulation, RANSAC poses.  python scripts/bench_images.py [--batch 16]"""
import argparse
import json
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vo_single_camera_sos_b200 import ops, workload  # noqa: E402
from vo_single_camera_sos_b200.features import FeatureFront, azimuthal_masks, step_images  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    ctx = ops.Context(0)
    B = a.batch
    w = workload.build(ctx, a.workload, batch=B, n_frames=2 * B + 1, seed=0, score_mode=ops.SCORE_BEARING)
    c = workload.CONFIGS[a.workload]
    rows, cols = w.cfg.pano_rows, w.cfg.pano_cols
    N = min(w.cfg.max_feat_per_bucket, c["feat"] // 12 + 1)
    valid = ((w.lut >> 48) & 0xF) != 0
    masks = []
    for view in range(2):
        v = cv2.erode(valid[view].cpu().numpy().astype(np.uint8), np.ones((7, 7), np.uint8)).astype(bool)
        masks.append(azimuthal_masks(rows, cols, 12) & (v[None] * 255).astype(np.uint8))
    front = FeatureFront(ctx, masks[0], masks[1], corners_per_bucket=N, max_feat_per_view=w.cfg.max_feat_per_view)
    renderer = workload.DeviceRenderer(ctx, w)
    sets = [torch.stack([renderer.render(w.trajectory[s * B + i]) for i in range(B)]).contiguous() for s in range(2)]
    fe = w.frontend(ctx)
    for s in range(3):
        step_images(fe, front, sets[s % 2], w.lut)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(a.steps):
        step_images(fe, front, sets[s % 2], w.lut)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    ctx.profile_begin()                                  # device time of this library's own launches inside one step
    step_images(fe, front, sets[0], w.lut)
    agg = {}
    for name, t in ctx.profile_end():
        agg[name.split("#")[0].split(":")[0]] = agg.get(name.split("#")[0].split(":")[0], 0.0) + t
    st = fe.buffers()["stats"].cpu().numpy()
    print(json.dumps({"metric": "image_in_frame_pairs_per_s", "value": B / (ms * 1e-3), "unit": "frame-pairs/s", "workload": a.workload,
                      "batch": B, "ms_per_step": ms, "library_ms_by_entry_point": {k: round(v, 3) for k, v in agg.items()},
                      "library_ms": round(sum(agg.values()), 3), "corners_per_bucket": N,
                      "stereo_correspondences": st[:, 0].tolist(), "temporal_correspondences": st[:, 1].tolist(),
                      "ransac_inliers": st[:, 2].tolist()}))


if __name__ == "__main__":
    main()
