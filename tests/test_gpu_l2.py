"""GPU parity: the float-descriptor L2 matcher (sos_l2_top2: tcgen05.mma kind::f16 on bfloat16 operands, exact for the
integer-valued descriptors cv2's SIFT returns) against goldens from the reference's FeatureMatcher("SIFT" / "SURF", "BF", k)
and against the exact-integer oracle — indices, float32 distances and output order bit-equal."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import hamming
from oracle.gen_golden import make_sift_like

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_feature_matcher_sift_and_surf_goldens(ctx):
    import cv2
    from omnistereo.camera_models import FeatureMatcher
    g = load_golden("l2.npz")
    for name, method, k, qk, tk in (("sift1", "SIFT", 1, "q", "t"), ("sift2", "SIFT", 2, "q", "t"), ("surf2", "SURF", 2, "q", "t"),
                                    ("surf1_64", "SURF", 1, "q64", "t64")):
        m = FeatureMatcher(method, "BF", k).match(query_descriptors=g[qk], train_descriptors=g[tk])
        assert isinstance(m[0], cv2.DMatch)
        assert [x.queryIdx for x in m] == g[f"{name}_q"].tolist() and [x.trainIdx for x in m] == g[f"{name}_t"].tolist(), name
        assert np.array_equal(np.array([x.distance for x in m], np.float32), g[f"{name}_d"]), name
    assert FeatureMatcher("SIFT", "BF", 1).match(np.zeros((0, 128), np.float32), g["t"]) == []


def test_ragged_segments_against_the_oracle(ctx):
    rng = np.random.default_rng(3)
    for dim in (128, 64, 33):
        sizes = [(700, 900), (130, 513), (257, 128), (5, 1300), (0, 10), (40, 0), (1, 1)]
        qs, ts = [], []
        for nq, nt in sizes:
            q, t = make_sift_like(rng, max(nq, nt, 1) + 40, max(nq, 1), max(nt, 1), dim=dim, n_ties=5 if nt > 20 else 0)
            q, t = q[:nq], t[:nt]
            if nq and nt > 300:                               # exact duplicates of a winner in a LATER 128-row tile
                i0 = hamming.l2_knn2(q, t)[0]
                for k in range(0, min(nq, 30), 3):
                    if i0[k] + 130 < nt:
                        t[i0[k] + 130] = t[i0[k]]
            qs.append(q); ts.append(t)
        seg_q = np.concatenate([[0], np.cumsum([len(x) for x in qs])]).astype(np.int32)
        seg_t = np.concatenate([[0], np.cumsum([len(x) for x in ts])]).astype(np.int32)
        for want_second in (True, False):
            out = ctx.l2_top2(dev(np.concatenate(qs)), dev(np.concatenate(ts)), dev(seg_q[:-1]), dev(np.diff(seg_q)), dev(seg_t[:-1]),
                              dev(np.diff(seg_t)), 700, 1300, want_second=want_second)
            i0, d0 = out[0].cpu().numpy(), out[1].cpu().numpy()
            for s, (q, t) in enumerate(zip(qs, ts)):
                a, b = seg_q[s], seg_q[s + 1]
                oi0, od0, oi1, od1 = hamming.l2_knn2(q, t)
                assert np.array_equal(i0[a:b], oi0) and np.array_equal(d0[a:b], od0), (dim, s)
                if want_second:
                    assert np.array_equal(out[2].cpu().numpy()[a:b], oi1) and np.array_equal(out[3].cpu().numpy()[a:b], od1), (dim, s)


def test_extreme_values_and_rejection_of_non_integer_descriptors(ctx):
    z = dev(np.zeros(1, np.int32))
    # all-255 against all-0: the largest possible distance, sqrt(128 * 255^2), and exact ties -> lowest train row
    q = np.full((300, 128), 255, np.float32)
    t = np.zeros((400, 128), np.float32)
    t[137] = 255                                   # one exact match in the second tile
    i0, d0, i1, d1 = ctx.l2_top2(dev(q), dev(t), z, dev(np.array([300], np.int32)), z, dev(np.array([400], np.int32)), 300, 400)
    assert (i0.cpu().numpy() == 137).all() and (d0.cpu().numpy() == 0).all()
    assert (i1.cpu().numpy() == 0).all() and np.array_equal(d1.cpu().numpy(), np.full(300, np.sqrt(np.float32(128 * 255 * 255)), np.float32))
    bad = q.copy()
    bad[5, 7] = 12.5
    with pytest.raises(ValueError):
        ctx.l2_top2(dev(bad), dev(t), z, dev(np.array([300], np.int32)), z, dev(np.array([400], np.int32)), 300, 400)
    bad[5, 7] = 256.0
    with pytest.raises(ValueError):
        ctx.l2_top2(dev(bad), dev(t), z, dev(np.array([300], np.int32)), z, dev(np.array([400], np.int32)), 300, 400)
