"""A/B of the two engines of sos_hamming_top2 on the C2 shapes of a front-end step (run on the GPU box):
temporal matching = 64 segments of ~4400 x ~4400 descriptors, stereo matching = 384 buckets of ~667 x ~667.
Checks that both engines return identical (index, distance) arrays and prints CUDA-event times per call."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vo_single_camera_sos_b200 import ops  # noqa: E402


def problem(rng, n_seg, n_lo, n_hi, cap):
    lens_q = rng.integers(n_lo, n_hi + 1, n_seg).astype(np.int32)
    lens_t = rng.integers(n_lo, n_hi + 1, n_seg).astype(np.int32)
    q = rng.integers(0, 256, (n_seg * cap, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (n_seg * cap, 32), dtype=np.uint8)
    # plant near matches so that distances spread like real data
    for s in range(n_seg):
        m = min(lens_q[s], lens_t[s]) // 2
        src = rng.integers(0, lens_t[s], m)
        noise = np.packbits(rng.random((m, 256)) < 0.08, axis=1)
        q[s * cap:s * cap + m] = t[s * cap + src] ^ noise
    start = (np.arange(n_seg) * cap).astype(np.int32)
    return q, t, start, lens_q, lens_t


def run(ctx, prob, cap, engine, want_second, reps=20):
    os.environ["SOS_HAMMING_ENGINE"] = engine
    q, t, start, lq, lt = prob
    d = lambda a: torch.from_numpy(a).cuda()
    qd, td, sd, lqd, ltd = d(q), d(t), d(start), d(lq), d(lt)
    out = ctx.hamming_top2(qd, td, sd, lqd, sd, ltd, cap, cap, want_second=want_second)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ctx.hamming_top2(qd, td, sd, lqd, sd, ltd, cap, cap, want_second=want_second, out=out)
    e0.record()
    for _ in range(reps):
        ctx.hamming_top2(qd, td, sd, lqd, sd, ltd, cap, cap, want_second=want_second, out=out)
    e1.record()
    torch.cuda.synchronize()
    return [None if o is None else o.cpu().numpy() for o in out], e0.elapsed_time(e1) / reps


def main():
    ctx = ops.Context(0)
    rng = np.random.default_rng(0)
    results = {}
    for name, (n_seg, lo, hi, cap) in {"temporal_c2": (64, 4200, 4600, 8192), "stereo_c2": (384, 600, 740, 2048),
                                       "temporal_c1": (64, 1000, 1200, 2048)}.items():
        prob = problem(rng, n_seg, lo, hi, cap)
        pairs = float((prob[3].astype(np.int64) * prob[4]).sum())
        for want_second in (False, True):
            ref, t_popc = run(ctx, prob, cap, "popc", want_second)
            got, t_mma = run(ctx, prob, cap, "mma", want_second)
            same = all((a is None and b is None) or np.array_equal(a, b) for a, b in zip(ref, got))
            if not same:
                for k, (a, b) in enumerate(zip(ref, got)):
                    if a is not None and not np.array_equal(a, b):
                        bad = np.nonzero(a != b)[0]
                        print(f"  MISMATCH {name} out[{k}]: {len(bad)} rows, first {bad[:8]}, popc {a[bad[:8]]}, mma {b[bad[:8]]}")
            results[f"{name}{'_top2' if want_second else ''}"] = dict(
                pairs=pairs, popc_ms=round(t_popc, 4), mma_ms=round(t_mma, 4), speedup=round(t_popc / t_mma, 2),
                mma_tera_pairs_per_s=round(pairs / t_mma / 1e9, 3), identical=bool(same))
            print(name, "top2" if want_second else "nn", results[f"{name}{'_top2' if want_second else ''}"], flush=True)
    print(json.dumps(results))
    return 0 if all(r["identical"] for r in results.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
