"""CPU (gloo, world_size 2): the host-side logic of the multi-GPU path — frame sharding and the packed-key MAX reduce
that combines the per-rank RANSAC winners."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vo_single_camera_sos_b200 import parallel


def test_shard_frames_partitions_every_pair_once():
    for n in (0, 1, 7, 64, 1000, 1001):
        for world in (1, 2, 3, 8):
            owned = []
            for r in range(world):
                a, b = parallel.shard_frames(n, world, r)
                assert 0 <= a <= b <= n
                owned += list(range(a, b))
            assert owned == list(range(n))
            sizes = [np.diff(parallel.shard_frames(n, world, r))[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_frames(10, 2, 2)


def test_key_packing_prefers_count_then_lowest_index():
    assert parallel.unpack_key(parallel.pack_key(17, 123456)) == (17, 123456)
    assert parallel.unpack_key(parallel.pack_key(0, 0)) == (0, 0)
    assert parallel.pack_key(-1, 5) == 0 and parallel.unpack_key(0) == (-1, -1)
    assert parallel.pack_key(10, 99) > parallel.pack_key(9, 0)          # higher count wins
    assert parallel.pack_key(10, 3) > parallel.pack_key(10, 4)          # tie: lower index wins
    assert parallel.pack_key(0, 65535) > parallel.pack_key(-1, 0)       # any valid model beats a rejected sample
    assert parallel.pack_key(2 ** 31 - 2, 0) < 2 ** 63                  # fits a signed int64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, counts, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, H = counts.shape
    lo, hi = parallel.shard_hypotheses(H, world, rank)
    keys = []
    for b in range(B):
        local = [parallel.pack_key(int(c), lo + i) for i, c in enumerate(counts[b, lo:hi])]
        keys.append(max(local) if local else 0)
    key = parallel.reduce_best_key(torch.tensor(keys, dtype=torch.int64))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), key.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_reduce_matches_single_list_argmax(tmp_path, world):
    rng = np.random.default_rng(world)
    B, H = 5, 1000
    counts = rng.integers(-1, 40, (B, H))      # many ties, some rejected samples
    counts[1, :] = -1                          # a problem with no valid hypothesis at all
    counts[2, :] = 7                           # all tied: index 0 must win
    counts[3, 600:] = -1
    counts[3, 777] = 39                        # the winner sits in the last rank's slice
    mp.spawn(_worker, args=(world, _free_port(), counts, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    for r in range(1, world):
        assert np.array_equal(res[0], res[r])   # every rank ends with the same winner
    for b in range(B):
        cnt, idx = parallel.unpack_key(res[0][b])
        if counts[b].max() < 0:
            assert (cnt, idx) == (-1, -1)
        else:
            assert idx == int(np.argmax(counts[b])) and cnt == counts[b].max()   # first maximum


# ---- config 3: sharded sequence (host logic of vo_single_camera_sos_b200/sequence.py) ---------------------------------
def _random_relative_poses(n, seed):
    rng = np.random.default_rng(seed)
    rel = np.zeros((n, 3, 4), np.float32)
    for i in range(n):
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        ang = rng.uniform(0.0, 0.05)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        rel[i, :, :3] = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        rel[i, :, 3] = rng.normal(0, 0.03, 3)
    stats = np.zeros((n, 4), np.int32)
    stats[:, 3] = rng.integers(0, 100, n)
    stats[0, 3] = -1                                  # frame 0 has no pair
    return rel, stats


def test_shard_plan_reads_the_overlap_frame():
    from vo_single_camera_sos_b200 import sequence
    for n, world, batch in ((1000, 8, 32), (1000, 1, 32), (10, 4, 3), (5, 8, 2)):
        covered = []
        for r in range(world):
            p = sequence.plan_shard(n, world, r, batch)
            assert p.read_first == max(p.first - 1, 0)
            assert p.n_batches * batch >= p.n_read > (p.n_batches - 1) * batch or p.n_read == 0
            covered += list(range(p.first, p.last))
        assert covered == list(range(n))


def _seq_worker(rank, world, port, n, out_dir):
    from vo_single_camera_sos_b200 import sequence
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rel, stats = _random_relative_poses(n, 3)        # every rank can compute any pair: pairs are independent
    plans = [sequence.plan_shard(n, world, r, 4) for r in range(world)]
    mine = plans[rank]
    counts = [p.last - p.first for p in plans]
    pose_all = sequence.gather_to_rank0(torch.from_numpy(rel[mine.first:mine.last].copy()), counts)
    stat_all = sequence.gather_to_rank0(torch.from_numpy(stats[mine.first:mine.last].copy()), counts)
    if rank == 0:
        np.save(os.path.join(out_dir, "rel.npy"), pose_all.numpy())
        np.save(os.path.join(out_dir, "stats.npy"), stat_all.numpy())
    else:
        assert pose_all is None and stat_all is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sequence_gather_and_chain_equal_the_single_rank_result(tmp_path, world):
    from vo_single_camera_sos_b200 import sequence
    n = 23                                            # not divisible: ranks own different numbers of frames
    mp.spawn(_seq_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    rel, stats = _random_relative_poses(n, 3)
    got_rel, got_stats = np.load(tmp_path / "rel.npy"), np.load(tmp_path / "stats.npy")
    assert np.array_equal(got_rel, rel) and np.array_equal(got_stats, stats)
    assert sequence.trajectory_digest(got_rel, got_stats) == sequence.trajectory_digest(rel, stats)
    traj, failed = sequence.chain_trajectory(got_rel, got_stats)
    assert failed == []
    T = np.eye(4)
    for f in range(1, n):                             # pose_est_tools.py:837 with the previous frame as the reference
        S = np.eye(4)
        S[:3] = rel[f]
        T = T @ S
        assert np.allclose(traj[f], T, atol=1e-12)
    # a rejected pair contributes the identity and is reported
    stats[7, 3] = -1
    traj2, failed2 = sequence.chain_trajectory(rel, stats)
    assert failed2 == [7] and np.array_equal(traj2[7], traj2[6])
    lines = []
    sequence.write_tum(str(tmp_path / "tum.txt"), traj)
    lines = open(tmp_path / "tum.txt").read().strip().split("\n")
    assert len(lines) == n and len(lines[3].split()) == 8 and lines[3].split()[0] == "3"
