"""GPU: ORB description of given keypoints (description half) against cv2.ORB.compute itself — the call
the reference makes (camera_models.py:1683, 1766).  Bit-exact: descriptors, dropped keypoints, gray conversion; and the
blur model identified by scripts/derive_orb_pattern.py (restated in oracle/orb.py)."""
import cv2
import numpy as np
import pytest
import torch

from oracle import orb as orbm

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def textured(rng, h, w, sigma=1.5):
    return cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (0, 0), sigma)


def test_bgr_to_gray_bit_exact(ctx):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (2, 123, 457, 3), dtype=np.uint8)
    img[0, :4, :4] = [[255, 255, 255]]
    img[0, 4:8, :4] = [[0, 0, 0]]
    got = ctx.bgr_to_gray(dev(img)).cpu().numpy()
    for b in range(2):
        assert np.array_equal(got[b], cv2.cvtColor(img[b], cv2.COLOR_BGR2GRAY))


def test_blur_model(ctx):
    rng = np.random.default_rng(1)
    for h, w in ((97, 130), (64, 64), (33, 7), (200, 1)):
        g = rng.integers(0, 256, (3, h, w), dtype=np.uint8)
        got = ctx.orb_blur(dev(g)).cpu().numpy()
        for b in range(3):
            assert np.array_equal(got[b], orbm.blur(g[b])), (h, w)


@pytest.mark.parametrize("angles", ["gft", "random"])
def test_describe_matches_cv2_orb_compute(ctx, angles):
    rng = np.random.default_rng(2 if angles == "gft" else 3)
    h, w, n = 240, 420, 1500
    orb = cv2.ORB_create(nfeatures=50)
    grays = np.stack([textured(rng, h, w, 1.0 + 0.7 * i) for i in range(3)])
    xy = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], 1).astype(np.float32)
    xy[:8] = [[31, 31], [30.99, 100], [w - 31, 50], [w - 31.01, 50], [100, h - 31], [100, h - 31.5], [31.5, 31.49], [200.5, 100.5]]
    ang = np.full(n, -1.0, np.float32) if angles == "gft" else rng.uniform(0, 360, n).astype(np.float32)
    img_idx = rng.integers(0, 3, n).astype(np.int32)
    desc, keep = ctx.orb_describe(dev(grays), dev(xy), None if angles == "gft" else dev(ang), dev(img_idx))
    desc, keep = desc.cpu().numpy(), keep.cpu().numpy().astype(bool)
    for b in range(3):
        sel = np.flatnonzero(img_idx == b)
        kps = [cv2.KeyPoint(float(xy[i, 0]), float(xy[i, 1]), 1.0, float(ang[i]), 1.0, 0, int(i)) for i in sel]
        kps2, want = orb.compute(grays[b], kps)                    # class_id carries the original index
        kept = np.array([k.class_id for k in kps2])
        assert np.array_equal(np.sort(kept), sel[keep[sel]])       # the same keypoints survive the border filter
        assert np.array_equal(desc[kept], want)                    # bit-exact descriptors
        assert not desc[sel[~keep[sel]]].any()
    assert 0 < (~keep).sum() < n // 2


def test_describe_reference_call_sequence(ctx):
    """As the reference uses it: goodFeaturesToTrack -> KeyPoint_convert -> ORB.compute on the (gray) panorama."""
    rng = np.random.default_rng(5)
    gray = textured(rng, 300, 800, 2.0)
    pts = cv2.goodFeaturesToTrack(image=gray, maxCorners=400, qualityLevel=0.01, minDistance=5, mask=None, useHarrisDetector=False)
    kps = list(cv2.KeyPoint_convert(pts.reshape(-1, 2)))   # (N, 1, 2) is rejected by OpenCV 4.13
    assert all(k.angle == -1 and k.octave == 0 for k in kps)
    kps2, want = cv2.ORB_create(nfeatures=400).compute(gray, kps)
    xy = pts.reshape(-1, 2).astype(np.float32)
    desc, keep = ctx.orb_describe(dev(gray), dev(xy))
    keep = keep.cpu().numpy().astype(bool)
    assert keep.sum() == len(kps2)
    assert np.array_equal(desc.cpu().numpy()[keep], want)
