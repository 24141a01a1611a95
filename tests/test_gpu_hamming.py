"""GPU parity: Hamming top-2 / match-select kernels vs the oracle (bit-exact) and the committed golden vectors."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import hamming
from oracle.gen_golden import make_descriptors

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


@pytest.fixture(autouse=True, params=["popc", "mma"])
def hamming_engine(request, monkeypatch):
    """Every test of this file runs on both engines of sos_hamming_top2: XOR + POPC on the integer pipe and the int8
    tensor-core engine (csrc/hamming_mma.cu).  The library reads SOS_HAMMING_ENGINE per call; without it the engine is
    chosen by problem size."""
    monkeypatch.setenv("SOS_HAMMING_ENGINE", request.param)
    return request.param


def run_top2(ctx, qs, ts):
    """qs, ts: lists of [n,32] uint8 arrays, one per segment."""
    seg_q = np.concatenate([[0], np.cumsum([len(x) for x in qs])]).astype(np.int32)
    seg_t = np.concatenate([[0], np.cumsum([len(x) for x in ts])]).astype(np.int32)
    q = np.concatenate(qs) if seg_q[-1] else np.zeros((0, 32), np.uint8)
    t = np.concatenate(ts) if seg_t[-1] else np.zeros((0, 32), np.uint8)
    qd = dev(q) if len(q) else torch.zeros((0, 32), dtype=torch.uint8, device="cuda")
    td = dev(t) if len(t) else torch.zeros((1, 32), dtype=torch.uint8, device="cuda")
    out = ctx.hamming_top2(qd, td, dev(seg_q[:-1]), dev(np.diff(seg_q)), dev(seg_t[:-1]), dev(np.diff(seg_t)),
                           max(len(x) for x in qs), max(len(x) for x in ts))
    return [o.cpu().numpy() for o in out], seg_q, seg_t, qd, td


def check_segments(res, seg_q, qs, ts):
    i0, d0, i1, d1 = res
    for s, (q, t) in enumerate(zip(qs, ts)):
        a, b = seg_q[s], seg_q[s + 1]
        oi0, od0, oi1, od1 = hamming.knn2(q, t)
        assert np.array_equal(i0[a:b], oi0) and np.array_equal(d0[a:b], od0)
        assert np.array_equal(i1[a:b], oi1) and np.array_equal(d1[a:b], od1)


def test_golden_vectors(ctx):
    g = load_golden("hamming.npz")
    res, seg_q, seg_t, _, _ = run_top2(ctx, [g["q"]], [g["t"]])
    i0, d0, i1, d1 = res
    assert np.array_equal(np.stack([i0, i1], 1), g["knn_t"])
    assert np.array_equal(np.stack([d0, d1], 1), g["knn_d"].astype(np.int32))


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 2), (3, 1), (255, 257), (256, 256), (1000, 1300), (2000, 2000)])
def test_single_segment_sizes(ctx, nq, nt):
    rng = np.random.default_rng(nq * 7 + nt)
    q, t = make_descriptors(rng, max(nq, nt) + 50, nq, nt, n_ties=min(8, nq // 4, nt // 4))
    res, seg_q, _, _, _ = run_top2(ctx, [q], [t])
    check_segments(res, seg_q, [q], [t])


def test_ragged_segments_with_empty_ones(ctx):
    rng = np.random.default_rng(11)
    sizes = [(37, 32), (0, 10), (52, 0), (18, 400), (300, 7), (1, 1), (0, 0), (513, 511)]
    qs, ts = [], []
    for nq, nt in sizes:
        q, t = make_descriptors(rng, max(nq, nt, 1) + 20, nq, nt, n_ties=2 if min(nq, nt) > 10 else 0)
        qs.append(q); ts.append(t)
    res, seg_q, _, _, _ = run_top2(ctx, qs, ts)
    check_segments(res, seg_q, qs, ts)


def test_nearest_only_path_many_tiles_and_ties(ctx):
    """want_second=False (the front-end's default 1-NN path) over several 128-row train tiles per segment, with exact
    duplicates of the best match planted in LATER tiles: the lowest train index must win (cv2.BFMatcher)."""
    rng = np.random.default_rng(5)
    qs, ts = [], []
    for nq, nt in [(700, 900), (130, 513), (257, 128), (5, 1300)]:
        q, t = make_descriptors(rng, max(nq, nt) + 40, nq, nt, n_ties=6)
        i0 = hamming.knn2(q, t)[0]
        for k in range(0, min(nq, 40), 3):                    # copy the winner of query k to a later train row
            later = int(rng.integers(i0[k] + 1, nt)) if i0[k] + 1 < nt else None
            if later is not None:
                t[later] = t[i0[k]]
        qs.append(q); ts.append(t)
    seg_q = np.concatenate([[0], np.cumsum([len(x) for x in qs])]).astype(np.int32)
    seg_t = np.concatenate([[0], np.cumsum([len(x) for x in ts])]).astype(np.int32)
    i0, d0, i1, d1 = ctx.hamming_top2(dev(np.concatenate(qs)), dev(np.concatenate(ts)), dev(seg_q[:-1]), dev(np.diff(seg_q)),
                                      dev(seg_t[:-1]), dev(np.diff(seg_t)), 700, 1300, want_second=False)
    assert i1 is None and d1 is None
    i0, d0 = i0.cpu().numpy(), d0.cpu().numpy()
    for s, (q, t) in enumerate(zip(qs, ts)):
        oi0, od0, _, _ = hamming.knn2(q, t)
        assert np.array_equal(i0[seg_q[s]:seg_q[s + 1]], oi0) and np.array_equal(d0[seg_q[s]:seg_q[s + 1]], od0)


def test_many_small_segments_walk_the_persistent_work_list(ctx):
    """More work items than SMs with short train ranges: the tensor-core engine runs persistent (one CTA per SM takes several
    items in turn, pipelines and tensor memory carried across items); ragged lengths, some empty segments."""
    rng = np.random.default_rng(8)
    n_seg = 420
    qs, ts = [], []
    for s_ in range(n_seg):
        nq = int(rng.integers(0, 300)) if s_ % 37 else 0
        nt = int(rng.integers(1, 400)) if s_ % 41 else 0
        q, t = make_descriptors(rng, max(nq, nt, 1) + 10, nq, nt, n_ties=1 if min(nq, nt) > 20 else 0)
        qs.append(q); ts.append(t)
    res, seg_q, _, _, _ = run_top2(ctx, qs, ts)
    check_segments(res, seg_q, qs, ts)


def test_train_rows_beyond_max_nt_are_ignored(ctx):
    """t_len[s] > max_nt: both engines clamp the train range to max_nt rows (the bound the scratch is sized for)."""
    rng = np.random.default_rng(6)
    q, t = make_descriptors(rng, 700, 300, 640, n_ties=3)
    z = dev(np.zeros(1, np.int32))
    i0, d0, i1, d1 = ctx.hamming_top2(dev(q), dev(t), z, dev(np.array([300], np.int32)), z, dev(np.array([640], np.int32)), 300, 512)
    oi0, od0, oi1, od1 = hamming.knn2(q, t[:512])
    assert np.array_equal(i0.cpu().numpy(), oi0) and np.array_equal(d1.cpu().numpy(), od1)


def test_all_equal_descriptors_tie_break(ctx):
    q = np.zeros((70, 32), np.uint8)
    t = np.zeros((300, 32), np.uint8)
    res, _, _, _, _ = run_top2(ctx, [q], [t])
    assert np.all(res[0] == 0) and np.all(res[2] == 1) and np.all(res[1] == 0) and np.all(res[3] == 0)
    t[:] = 0xFF
    res, _, _, _, _ = run_top2(ctx, [q], [t])
    assert np.all(res[1] == 256) and np.all(res[0] == 0) and np.all(res[2] == 1)


def test_full_size_8k_property(ctx):
    """C2-size (8000 x 8000): distances are symmetric, so the (q -> t) and (t -> q) scans must agree on mutual pairs,
    and a permutation of the train rows must permute the answer; spot-check 64 rows against the oracle."""
    rng = np.random.default_rng(5)
    q, t = make_descriptors(rng, 10000, 8000, 8000, n_ties=16)
    res, seg_q, _, _, _ = run_top2(ctx, [q], [t])
    rev, _, _, _, _ = run_top2(ctx, [t], [q])
    i0, d0 = res[0], res[1]
    r0, rd0 = rev[0], rev[1]
    mutual = r0[i0] == np.arange(len(q))
    assert mutual.sum() > 4000
    assert np.array_equal(d0[mutual], rd0[i0[mutual]])
    rows = rng.integers(0, 8000, 64)
    oi0, od0, oi1, od1 = hamming.knn2(q[rows], t)
    assert np.array_equal(i0[rows], oi0) and np.array_equal(d0[rows], od0)
    assert np.array_equal(res[2][rows], oi1) and np.array_equal(res[3][rows], od1)
    perm = rng.permutation(8000)
    resp, _, _, _, _ = run_top2(ctx, [q], [t[perm]])
    same_d = np.array_equal(resp[1], d0)
    assert same_d
    # where the best distance is unique the permuted index must map back to the same train row
    uniq = d0 < res[3]
    assert np.array_equal(perm[resp[0][uniq]], i0[uniq])


@pytest.mark.parametrize("mode", ["nn", "ratio", "cross"])
def test_match_select_modes_and_gate(ctx, mode):
    rng = np.random.default_rng(21)
    sizes = [(300, 280), (0, 5), (41, 1), (1500, 1400), (64, 64)]
    qs, ts, pqs, pts = [], [], [], []
    for nq, nt in sizes:
        q, t = make_descriptors(rng, max(nq, nt, 1) + 30, nq, nt, n_ties=4 if min(nq, nt) > 20 else 0)
        qs.append(q); ts.append(t)
        pq = rng.uniform(0, 1200, (nq, 2)).astype(np.float32)
        pt = rng.uniform(0, 1200, (nt, 2)).astype(np.float32)
        if nq and nt:  # make many pairs geometrically close, some exactly on the gate boundary
            i0 = hamming.knn2(q, t)[0]
            pt[i0[: nq // 2]] = pq[: nq // 2] + rng.uniform(-3, 3, (nq // 2, 2)).astype(np.float32)
            pt[i0[0]] = pq[0] + np.array([2.5, 1.0], np.float32)
        pqs.append(pq); pts.append(pt)
    res, seg_q, seg_t, qd, td = run_top2(ctx, qs, ts)
    i0, d0, i1, d1 = [dev(x) for x in res]
    rev = None
    if mode == "cross":
        rres, _, _, _, _ = run_top2(ctx, ts, qs)
        rev = dev(rres[0])
    px_q = dev(np.concatenate(pqs)); px_t = dev(np.concatenate(pts))
    code = {"nn": 0, "ratio": 1, "cross": 2}[mode]
    for gate in (False, True):
        oq, ot, od, oc = ctx.match_select(code, i0, d0, d1, dev(seg_q[:-1]), dev(np.diff(seg_q)), dev(seg_t[:-1]), rev_idx0=rev,
                                          px_q=px_q if gate else None, px_t=px_t if gate else None,
                                          max_du=2.5 if gate else -1, min_dv=1.0 if gate else -1, ratio=0.75)
        oq, ot, od, oc = (x.cpu().numpy() for x in (oq, ot, od, oc))
        for s, (q, t) in enumerate(zip(qs, ts)):
            eq, et, ed = hamming.match_select(q, t, mode, 0.75, pqs[s] if gate else None, pts[s] if gate else None,
                                              2.5 if gate else -1, 1.0 if gate else -1)
            a = seg_q[s]
            assert oc[s] == len(eq), (mode, gate, s)
            assert np.array_equal(oq[a:a + oc[s]] - seg_q[s], eq)
            assert np.array_equal(ot[a:a + oc[s]] - seg_t[s], et)
            assert np.array_equal(od[a:a + oc[s]], ed)


def test_reference_frame_matching_goldens(ctx):
    """match_features_panoramic_top_bottom / match_features_frame_to_frame outputs of the reference."""
    g = load_golden("matching_frames.npz")
    nb = int(g["n_buckets"])
    qs = [g[f"b{b}_d_bot"] for b in range(nb)]
    ts = [g[f"b{b}_d_top"] for b in range(nb)]
    pq = np.concatenate([g[f"b{b}_pt_bot"] for b in range(nb)])
    pt = np.concatenate([g[f"b{b}_pt_top"] for b in range(nb)])
    res, seg_q, seg_t, _, _ = run_top2(ctx, qs, ts)
    i0, d0, i1, d1 = [dev(x) for x in res]
    oq, ot, od, oc = ctx.match_select(0, i0, d0, d1, dev(seg_q[:-1]), dev(np.diff(seg_q)), dev(seg_t[:-1]), px_q=dev(pq), px_t=dev(pt),
                                      max_du=2.5, min_dv=1.0)
    oq, ot, oc = oq.cpu().numpy(), ot.cpu().numpy(), oc.cpu().numpy()
    m_top = np.concatenate([pt[ot[seg_q[s]:seg_q[s] + oc[s]]] for s in range(nb)])
    m_bot = np.concatenate([pq[oq[seg_q[s]:seg_q[s] + oc[s]]] for s in range(nb)])
    assert np.array_equal(m_top.astype(np.float64), g["stereo_m_top"][:, :2])
    assert np.array_equal(m_bot.astype(np.float64), g["stereo_m_bot"][:, :2])
    # temporal
    q, t = g["f2f_q"], g["f2f_t"]
    res, seg_q, seg_t, _, _ = run_top2(ctx, [q], [t])
    i0, d0, i1, d1 = [dev(x) for x in res]
    oq, ot, od, oc = ctx.match_select(0, i0, d0, d1, dev(seg_q[:-1]), dev(np.diff(seg_q)), dev(seg_t[:-1]), px_q=dev(g["f2f_pq"][:, :2], torch.float32),
                                      px_t=dev(g["f2f_pt"][:, :2], torch.float32), max_du=float(g["f2f_max_du"]), min_dv=-1.0)
    n = int(oc[0])
    assert np.array_equal(oq[:n].cpu().numpy(), g["f2f_query_idx"]) and np.array_equal(ot[:n].cpu().numpy(), g["f2f_train_idx"])
