"""Multi-GPU plumbing of the SOS front-end: one process per GPU, torch.distributed for the rendezvous.

Two things shard:
  * frame pairs — independent units, contiguous blocks of the sequence per rank with one overlapping frame at each
    boundary, NO data-path collective (weak scaling);
  * the hypotheses of ONE huge RANSAC problem (BASELINE config 4: 50 k correspondences x 65 536 hypotheses) — every rank
    scores its slice of the shared seeded hypothesis list and the winners are combined with a single 8-byte MAX
    all-reduce on the packed key (inlier_count + 1) << 32 | (0xFFFFFFFF - global hypothesis index), which prefers the
    higher count and, on ties, the LOWER index — the same first-maximum rule as the single-GPU argmax.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

KEY_IDX_MASK = 0xFFFFFFFF


def shard_frames(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Frames [first, last) processed by `rank`; pair (f-1, f) is owned by the rank that owns frame f, and every rank
    but the first also READS frame first-1 as its initial reference (the one-frame overlap)."""
    if n_frames < 0 or world < 1 or not 0 <= rank < world:
        raise ValueError("bad shard request")
    base, extra = divmod(n_frames, world)
    first = rank * base + min(rank, extra)
    last = first + base + (1 if rank < extra else 0)
    return first, last


def shard_hypotheses(n_hyp: int, world: int, rank: int) -> Tuple[int, int]:
    return shard_frames(n_hyp, world, rank)


def pack_key(count: int, global_index: int) -> int:
    """Same packing as csrc/ransac.cu::argmax_kernel; count < 0 (rejected sample) packs to 0."""
    if count < 0:
        return 0
    return ((count + 1) << 32) | (KEY_IDX_MASK - (global_index & KEY_IDX_MASK))


def unpack_key(key: int) -> Tuple[int, int]:
    """-> (count, global hypothesis index); (-1, -1) when no hypothesis was valid."""
    key = int(key)
    if key == 0:
        return -1, -1
    return (key >> 32) - 1, KEY_IDX_MASK - (key & KEY_IDX_MASK)


def reduce_best_key(local_key: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce(MAX) of the packed int64 keys ([n_problems]); works on NCCL (cuda) and gloo (cpu) tensors.
    Keys are < 2^63 because counts are < 2^31, so signed max == unsigned max."""
    if local_key.dtype != torch.int64:
        raise TypeError("keys must be int64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_key, op=dist.ReduceOp.MAX, group=group)
    return local_key


def ransac_split(ctx, p_ref, p_cur, n, hyp_all, score_mode: int, threshold: float, f_cur=None, cam=None, rig=None,
                 n_cams: int = 0, group=None, rank: Optional[int] = None, world: Optional[int] = None):
    """RANSAC over the full hypothesis list `hyp_all` [H,3] with the list split across the ranks of `group`.

    Every rank holds the same correspondences (1.2 MB at 50 k points: replicated).  Returns
    (pose [B,3,4], count [B], mask [B,cap], winner [B] global hypothesis index) — identical on every rank and
    identical to the single-GPU result on the undivided list."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_hypotheses(hyp_all.shape[0], world, rank)
    local = hyp_all[lo:hi].contiguous()
    _, _, _, _, key = ctx.ransac_p3d(p_ref, p_cur, n, local, score_mode, threshold, f_cur=f_cur, cam=cam, rig=rig,
                                     n_cams=n_cams, hyp_offset=lo, want_mask=False)
    key = reduce_best_key(key, group)
    # winner's sample triple, gathered on the device (no host round trip): index = 0xFFFFFFFF - low 32 bits
    win = (KEY_IDX_MASK - (key & KEY_IDX_MASK)).clamp_(0, hyp_all.shape[0] - 1)
    valid = key != 0
    rows = hyp_all.index_select(0, win.to(torch.int64)).contiguous()
    pose, count, mask = ctx.ransac_p3d_eval(p_ref, p_cur, n, rows, score_mode, threshold, f_cur=f_cur, cam=cam, rig=rig,
                                            n_cams=n_cams)
    winner = torch.where(valid, win, torch.full_like(win, -1)).to(torch.int32)
    count = torch.where(valid, count, torch.full_like(count, -1))
    mask = mask * valid[:, None].to(mask.dtype)
    return pose, count, mask, winner
