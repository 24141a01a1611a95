"""BASELINE config 3: one long SOS sequence whose frame pairs are sharded across the GPUs of a box.

The unit of work is the consecutive frame pair (f-1, f): remap + stereo matching + triangulation of frame f, temporal
matching against frame f-1, RANSAC, refinement — what the reference's loop does per frame (pose_est_tools.py:1416-1628)
with the previous frame as the tracking reference.  Pairs are independent, so rank r takes the contiguous block
parallel.shard_frames gives it and additionally READS frame first-1 (the one-frame overlap) as the reference of its first
pair.  No data-path collective: the only exchange is the gather of the relative poses (48 bytes per pair) to rank 0, which
chains them into the trajectory (pose_est_tools.py:837) and writes the TUM file (pose_est_tools.py:1609-1612).  The result
does not depend on the number of ranks: a pair's pose is a function of its two frames only.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import parallel, workload
from .driver import tum_line

INPUT_KEYS = workload.INPUT_KEYS


@dataclass
class ShardPlan:
    first: int          # frames [first, last) are OWNED by this rank (pair (f-1, f) belongs to the owner of f)
    last: int
    read_first: int     # first frame the rank reads: first - 1 for every rank but the one that owns frame 0
    n_batches: int

    @property
    def n_read(self) -> int:
        return self.last - self.read_first


def plan_shard(n_frames: int, world: int, rank: int, batch: int) -> ShardPlan:
    first, last = parallel.shard_frames(n_frames, world, rank)
    read_first = max(first - 1, 0)
    n_read = last - read_first
    return ShardPlan(first, last, read_first, (n_read + batch - 1) // batch if n_read > 0 else 0)


def make_shard_batches(w: workload.Workload, plan: ShardPlan, batch: int, omni: Optional[np.ndarray] = None) -> List[list]:
    """Pinned host input batches for the frames the shard reads; the tail of the last batch holds empty frames (no
    features: nothing to match).  `omni` [H,W,3]: one image used for every frame (the remap's work does not depend on the
    content); None renders nothing (black images)."""
    out = []
    for b in range(plan.n_batches):
        f0 = plan.read_first + b * batch
        cnt = min(batch, plan.last - f0)
        fr = workload.make_frames(w, f0, cnt, render=False)
        full = {}
        for k in INPUT_KEYS:
            a = np.zeros((batch,) + fr[k].shape[1:], fr[k].dtype)
            a[:cnt] = fr[k]
            full[k] = a
        if omni is not None:
            full["omni"][:cnt] = omni
        out.append(workload.to_pinned(full))
    return out


def run_shard(fe, batches: Sequence[list], plan: ShardPlan, batch: int) -> Tuple[np.ndarray, np.ndarray]:
    """All batches of the shard through the host API (copies of batch k+1 overlap the kernels of batch k).
    -> (relative poses [n_owned, 3, 4] float32 of the pairs (f-1, f) for f in [first, last), stats [n_owned, 4])."""
    fe.reset()
    poses, stats = [], []
    prev = None
    for b in batches:
        tk = fe.submit_host(*b)
        if prev is not None:
            p, s = fe.wait_host(prev)
            poses.append(p); stats.append(s)
        prev = tk
    if prev is not None:
        p, s = fe.wait_host(prev)
        poses.append(p); stats.append(s)
    n_owned = plan.last - plan.first
    if not poses:
        return np.zeros((0, 3, 4), np.float32), np.zeros((0, 4), np.int32)
    poses, stats = np.concatenate(poses), np.concatenate(stats)
    skip = plan.first - plan.read_first            # the overlap frame has no pair of its own on this rank
    return poses[skip:skip + n_owned].copy(), stats[skip:skip + n_owned].copy()


def gather_to_rank0(local: torch.Tensor, counts: Sequence[int], group=None) -> Optional[torch.Tensor]:
    """Concatenate per-rank row blocks (rank order) on rank 0.  Works on NCCL (cuda) and gloo (cpu) tensors; ranks may
    own different numbers of rows (padded to the maximum for the collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    m = max(counts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    if rank != 0:
        return None
    return torch.cat([out[r][: counts[r]] for r in range(world)])


def chain_trajectory(rel: np.ndarray, stats: np.ndarray, units_to_m: float = 1.0):
    """Relative poses of the pairs (f-1, f), f = 0..n-1 (entry 0 = the first frame, which has no pair) -> absolute poses
    T_C_f_wrt_C_0 [n,4,4] float64: T_f = T_{f-1} @ T_rel(f), the reference's composition with the previous frame as the
    tracking reference (pose_est_tools.py:837).  A pair RANSAC rejected (best hypothesis < 0) contributes the identity and
    is reported."""
    n = len(rel)
    T = np.eye(4)
    out = np.empty((n, 4, 4))
    out[0] = T
    failed = []
    for f in range(1, n):
        if stats[f, 3] < 0 or not np.isfinite(rel[f]).all():
            failed.append(f)
        else:
            S = np.eye(4)
            S[:3] = rel[f].astype(np.float64)
            S[:3, 3] *= units_to_m
            T = T @ S
        out[f] = T
    return out, failed


def trajectory_digest(rel: np.ndarray, stats: np.ndarray) -> str:
    """SHA-256 over the raw relative poses (pairs 1..n-1) and their statistics: equal digests <=> bit-identical results,
    whatever the number of ranks that produced them."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(rel[1:]).tobytes())
    h.update(np.ascontiguousarray(stats[1:]).tobytes())
    return h.hexdigest()


def write_tum(path: str, poses: np.ndarray, frame_ids: Optional[Sequence[int]] = None):
    with open(path, "w") as f:
        for i, T in enumerate(poses):
            print(tum_line(i if frame_ids is None else frame_ids[i], T), file=f)
