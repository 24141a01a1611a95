"""GPU: Shi-Tomasi detection (detection half) against cv2.cornerMinEigenVal / cv2.goodFeaturesToTrack,
the calls the reference makes (camera_models.py:1737).  The corner measure is bit-equal except at cv2's SIMD tail
columns; the corner lists are identical (same points, same order) on textured images."""
import cv2
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def textured(rng, h, w, sigma):
    return cv2.GaussianBlur(rng.integers(0, 256, (h, w), dtype=np.uint8), (0, 0), sigma)


def test_corner_measure_vs_cv2(ctx):
    rng = np.random.default_rng(0)
    for h, w in ((200, 300), (97, 131), (64, 512)):
        g = np.stack([textured(rng, h, w, 1.0 + i) for i in range(2)])
        got = ctx.corner_min_eigenval(dev(g)).cpu().numpy()
        for b in range(2):
            ref = cv2.cornerMinEigenVal(g[b], 3, ksize=3)
            same = got[b] == ref
            # all but the 3x3 neighbourhoods of cv2's SIMD tail columns (their last bit depends on width and CPU)
            assert same.mean() > 0.97, same.mean()
            assert np.abs(got[b] - ref).max() <= 1e-6 * np.abs(ref).max()


def cv_gft(gray, mask, n, q=0.01, d=5):
    pts = cv2.goodFeaturesToTrack(image=gray, maxCorners=n, qualityLevel=q, minDistance=d, mask=mask, useHarrisDetector=False)
    return np.zeros((0, 2), np.float32) if pts is None else pts.reshape(-1, 2)


@pytest.mark.parametrize("min_distance", [5.0, 0.0, 9.0])
def test_gft_matches_cv2(ctx, min_distance):
    rng = np.random.default_rng(1)
    h, w, n_img, n_masks, N = 180, 640, 3, 4, 300
    g = np.stack([textured(rng, h, w, 1.2 + 0.6 * i) for i in range(n_img)])
    masks = np.zeros((n_masks, h, w), np.uint8)
    for m in range(n_masks):                                   # azimuthal column ranges ANDed with an elevation band
        masks[m, 12:h - 9, m * w // n_masks:(m + 1) * w // n_masks] = 255
    masks[3, :, :40] = 255                                     # overlapping masks are allowed
    xy, cnt = ctx.gft_detect(dev(g), dev(masks), N, 0.01, min_distance)
    xy, cnt = xy.cpu().numpy(), cnt.cpu().numpy()
    total = same = 0
    for b in range(n_img):
        for m in range(n_masks):
            want = cv_gft(g[b], masks[m], N, 0.01, min_distance)
            got = xy[b, m, :cnt[b, m]]
            assert abs(len(got) - len(want)) <= 2
            k = min(len(got), len(want))
            total += len(want)
            same += int((got[:k] == want[:k]).all(axis=1).sum())
    assert total > 1500 and same >= 0.99 * total, (same, total)


def test_gft_no_mask_and_caps(ctx):
    rng = np.random.default_rng(2)
    g = textured(rng, 240, 320, 1.5)
    for N in (10, 1000):
        xy, cnt = ctx.gft_detect(dev(g), None, N)
        want = cv_gft(g, None, N)
        got = xy.cpu().numpy()[0, 0, :int(cnt[0, 0])]
        assert len(got) == len(want) and (got == want).all(axis=1).mean() >= 0.99
    flat = np.full((64, 64), 77, np.uint8)
    xy, cnt = ctx.gft_detect(dev(flat), None, 50)
    assert int(cnt[0, 0]) == 0 and cv_gft(flat, None, 50).shape[0] == 0


@pytest.mark.parametrize("N,min_distance", [(300, 5.0), (500, 25.0)])
def test_gft_long_candidate_lists(ctx, N, min_distance):
    """~10 k candidates in one list: the selection sorts only a histogram-picked prefix of the strongest candidates
    (N = 300) and falls back to the whole list when the prefix cannot supply N corners (minDistance = 25)."""
    rng = np.random.default_rng(3)
    g = textured(rng, 300, 800, 1.5)
    xy, cnt = ctx.gft_detect(dev(g), None, N, 0.01, min_distance)
    want = cv_gft(g, None, N, 0.01, min_distance)
    got = xy.cpu().numpy()[0, 0, :int(cnt[0, 0])]
    assert len(want) > 50 and abs(len(got) - len(want)) <= 1
    k = min(len(got), len(want))
    assert (got[:k] == want[:k]).all(axis=1).mean() >= 0.99
