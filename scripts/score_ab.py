"""A/B of the score-kernel variants (SOS_SCORE_VARIANT, experiment) and the Hamming ring depth (SOS_HAMMING_STAGES) on the GPU
box.  Times sos_ransac_p3d (C2 shape, 35 % inliers) per variant and checks that every variant returns the same counts."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vo_single_camera_sos_b200 import ops  # noqa: E402
from scripts import kbench, hamming_ab  # noqa: E402


def main():
    ctx = ops.Context(0)
    rng = np.random.default_rng(0)
    B, n, H, cap = 32, 6650, 4096, 16384
    p_cur = rng.normal(size=(B, cap, 3)).astype(np.float32) * 2
    ang = 0.02
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
    p_ref = p_cur @ R.T + np.float32([0.03, 0.01, -0.02]) + rng.normal(0, 0.01, p_cur.shape).astype(np.float32)
    out = rng.random((B, cap)) > 0.35
    p_ref[out] = rng.normal(size=(int(out.sum()), 3)).astype(np.float32) * 2
    f = p_cur / np.linalg.norm(p_cur, axis=2, keepdims=True)
    cam = np.zeros((B, cap), np.uint8)
    cam[:, n // 2:] = 1
    rig = np.zeros((2, 3, 4)); rig[:, :, :3] = np.eye(3); rig[0, 2, 3] = 0.12
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    hyp = d(rng.integers(0, 2 ** 32, (H, 3), dtype=np.uint64).astype(np.uint32).view(np.int32))
    args = (d(p_ref), d(p_cur), torch.full((B,), n, dtype=torch.int32, device="cuda"), hyp, ops.SCORE_BEARING,
            1.0 - np.cos(np.deg2rad(5.0)))
    kw = dict(f_cur=d(f.astype(np.float32)), cam=d(cam), rig=rig, n_cams=2)
    ref = None
    res = {}
    for v in (0, 1, 2, 3, 4, 0):
        os.environ["SOS_SCORE_VARIANT"] = str(v)
        o = ctx.ransac_p3d(*args, **kw)
        torch.cuda.synchronize()
        got = (o[1].cpu().numpy(), o[2].cpu().numpy(), o[3].cpu().numpy())
        if ref is None:
            ref = got
        same = all(np.array_equal(a, b) for a, b in zip(ref, got))
        ms = kbench.timeit(lambda: ctx.ransac_p3d(*args, **kw), iters=10)
        res[f"score_variant_{v}"] = dict(ms_all_ransac_kernels=round(ms, 4), identical=bool(same), inliers=int(got[1][0]))
        print("score variant", v, res[f"score_variant_{v}"], flush=True)
    os.environ.pop("SOS_SCORE_VARIANT")
    prob = hamming_ab.problem(np.random.default_rng(1), 64, 4200, 4600, 8192)
    prob_s = hamming_ab.problem(np.random.default_rng(2), 384, 600, 740, 2048)
    for st in ("3", "4", "3", "4"):
        os.environ["SOS_HAMMING_STAGES"] = st
        _, t = hamming_ab.run(ctx, prob, 8192, "mma", False)
        _, ts = hamming_ab.run(ctx, prob_s, 2048, "mma", False)
        print("hamming stages", st, "temporal ms", round(t, 4), "stereo ms", round(ts, 4), flush=True)
        res[f"hamming_stages_{st}"] = dict(temporal_ms=round(t, 4), stereo_ms=round(ts, 4))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
