#!/usr/bin/env python
"""BASELINE config 4 — stress RANSAC: 50 000 3D-3D correspondences x 65 536 hypotheses, the hypothesis list split across
the GPUs of one box and the winners combined by ONE 8-byte NCCL MAX all-reduce.

    python scripts/bench_c4.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_c4.py

Strong scaling (the total work is fixed).  Rank 0 also runs the undivided list on its own GPU and checks that the split
run returns the identical winner, inlier count, inlier mask and pose.  Prints one JSON line.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vo_single_camera_sos_b200 import ops, parallel  # noqa: E402


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ops.Context(local)
    n, H, steps, warm = 50000, 65536, 20, 3
    rng = np.random.default_rng(4)  # same data on every rank: the correspondences are replicated
    p_cur = rng.normal(size=(1, n, 3))
    p_cur = (p_cur / np.linalg.norm(p_cur, axis=2, keepdims=True) * rng.uniform(0.5, 7.0, (1, n, 1))).astype(np.float32)
    ang = np.deg2rad(2.0)
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
    p_ref = (p_cur @ R.T + np.float32([0.03, -0.01, 0.02]) + rng.normal(0, 0.005, p_cur.shape)).astype(np.float32)
    out = rng.random(n) > 0.35
    p_ref[0, out] = (rng.normal(size=(int(out.sum()), 3)) * 3).astype(np.float32)
    hyp = rng.integers(0, 2 ** 32, (H, 3), dtype=np.uint64).astype(np.uint32)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    P_ref, P_cur, N = d(p_ref), d(p_cur), torch.tensor([n], dtype=torch.int32, device="cuda")
    hyp_d = d(hyp.view(np.int32))
    run = lambda: parallel.ransac_split(ctx, P_ref, P_cur, N, hyp_d, ops.SCORE_EUCLID, 0.05)
    for _ in range(warm):
        res = run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        res = run()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    pose, count, mask, winner = res
    verified = None
    if rank == 0:
        fp, fh, fc, fm, _ = ctx.ransac_p3d(P_ref, P_cur, N, hyp_d, ops.SCORE_EUCLID, 0.05)
        verified = bool(int(fh[0]) == int(winner[0]) and int(fc[0]) == int(count[0]) and torch.equal(fm, mask)
                        and torch.equal(fp, pose))
        pairs = float(n) * H
        print(json.dumps({
            "metric": "ransac_hypothesis_point_pairs_per_s", "value": pairs / (float(ms[0]) * 1e-3), "unit": "pairs/s",
            "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": float(ms[0]), "higher_is_better": True,
            "scaling": "strong", "config": {"workload": f"c4: {n} correspondences x {H} hypotheses, euclid score",
                                            "collective": "one int64 MAX all-reduce per step (NCCL)"},
            "winner": int(winner[0]), "inliers": int(count[0]), "matches_single_gpu_full_list": verified,
            "tflops_at_30_flop_per_pair": pairs * 30 / (float(ms[0]) * 1e-3) / 1e12}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and verified is False:
        sys.exit(1)


if __name__ == "__main__":
    main()
