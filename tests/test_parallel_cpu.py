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
