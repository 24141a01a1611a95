"""GPU: BASELINE config 3 — a sequence sharded into contiguous blocks of frames with a one-frame overlap gives the same
relative poses, bit for bit, as the undivided run (vo_single_camera_sos_b200/sequence.py), and the chained trajectory
follows the ground-truth motion.  Ranks are emulated one after the other on the single test GPU; the multi-process gather is
covered on CPU (tests/test_parallel_cpu.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sharded_sequence_equals_the_undivided_run(ctx):
    from vo_single_camera_sos_b200 import sequence, workload
    n, batch = 23, 4
    w = workload.build(ctx, "tiny", batch=batch, n_frames=n, seed=6)
    fe = w.frontend(ctx)
    plan = sequence.plan_shard(n, 1, 0, batch)
    rel1, st1 = sequence.run_shard(fe, sequence.make_shard_batches(w, plan, batch), plan, batch)
    assert rel1.shape == (n, 3, 4) and st1[0, 3] == -1 and (st1[1:, 3] >= 0).all()
    for world in (2, 3, 5):
        rel, st = [], []
        for r in range(world):
            p = sequence.plan_shard(n, world, r, batch)
            a, b = sequence.run_shard(fe, sequence.make_shard_batches(w, p, batch), p, batch)
            assert len(a) == p.last - p.first
            rel.append(a); st.append(b)
        rel, st = np.concatenate(rel), np.concatenate(st)
        assert np.array_equal(rel[1:], rel1[1:]) and np.array_equal(st[1:], st1[1:])
        assert sequence.trajectory_digest(rel, st) == sequence.trajectory_digest(rel1, st1)
    traj, failed = sequence.chain_trajectory(rel1, st1)
    assert failed == []
    # sanity against the ground-truth motion, pair by pair (tiny geometry: the bars of tests/test_gpu_frontend.py)
    for f in range(1, n):
        gt = np.linalg.inv(w.trajectory[f - 1]) @ w.trajectory[f]
        assert np.allclose(rel1[f][:, :3], gt[:3, :3], atol=0.05) and np.allclose(rel1[f][:, 3], gt[:3, 3], atol=0.15)
    assert np.isfinite(traj).all()
    fe.close()
