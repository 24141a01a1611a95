#!/usr/bin/env python
"""Throughput of the batched keyframe-mode VO driver (BASELINE config 3 geometry) on one GPU:
python scripts/bench_vo.py [--workload c1] [--frames 256] [--batch 32] [--refine arun|lm]
Frames are synthetic (features from the random scene, blank omni images: the remap does the same work on any content).
Prints one JSON line: frames/s through BatchedVO.run (H2D of every batch, device chain, D2H of poses, host policy)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vo_single_camera_sos_b200 import ops, workload  # noqa: E402
from vo_single_camera_sos_b200.driver import BatchedVO, INPUT_KEYS, KeyframePolicy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c1")
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--refine", default="arun", choices=["arun", "lm"])
    ap.add_argument("--pos-min", type=float, default=0.01, help="keyframe translation threshold [m] (reference: 0.01)")
    a = ap.parse_args()
    ctx = ops.Context(0)
    w = workload.build(ctx, a.workload, batch=a.batch, n_frames=a.frames, seed=0, score_mode=ops.SCORE_BEARING)
    w.cfg.keyframe_mode = True
    w.cfg.refit = ops.REFINE_LM if a.refine == "lm" else ops.REFINE_ARUN
    t0 = time.perf_counter()
    fr = workload.make_frames(w, 0, a.frames, render=False)
    frames = [{k: fr[k][i] for k in INPUT_KEYS} for i in range(a.frames)]
    t_gen = time.perf_counter() - t0
    vo = BatchedVO(ctx, w.cfg, w.lut, w.hyp, KeyframePolicy(pos_min=a.pos_min))
    vo.run(frames[: 2 * a.batch])            # warm-up: graph capture, arena sizing
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = vo.run(frames)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    T_gt = np.linalg.inv(w.trajectory[0]) @ w.trajectory[len(res.frame_ids) - 1]
    err = float(np.linalg.norm(res.poses_wrt_S[-1][:3, 3] - T_gt[:3, 3]))
    print(json.dumps({
        "metric": "vo_frames_per_s", "value": len(res.frame_ids) / dt, "unit": "frames/s", "workload": a.workload,
        "frames": len(res.frame_ids), "batch": a.batch, "refine": a.refine, "status": res.status,
        "keyframes": len(res.keyframe_ids), "device_steps": res.device_steps, "device_retracks": res.device_retracks,
        "seconds": dt, "feature_synthesis_s": t_gen, "end_position_error_m": err,
        "path_length_m": float(sum(np.linalg.norm((np.linalg.inv(w.trajectory[i]) @ w.trajectory[i + 1])[:3, 3])
                                   for i in range(len(res.frame_ids) - 1)))}))


if __name__ == "__main__":
    main()
