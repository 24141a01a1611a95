"""GPU: the 11 x 11 median filter (SURVEY §8f N3 pre-processing) against cv2.medianBlur — bit-exact, 1 and 3 channels,
ragged sizes, constant and extreme images."""
import cv2
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape", [(97, 131, 3), (40, 33, 3), (260, 70, 3), (12, 300, 3), (129, 16, 1), (64, 65, 1)])
def test_median_blur_bit_exact(ctx, shape):
    rng = np.random.default_rng(sum(shape))
    h, w, ch = shape
    imgs = [rng.integers(0, 256, shape, dtype=np.uint8),
            cv2.GaussianBlur(rng.integers(0, 256, shape, dtype=np.uint8), (0, 0), 2.0).reshape(shape),
            np.full(shape, 255, np.uint8), (rng.random(shape) < 0.5).astype(np.uint8) * 255]
    batch = np.stack(imgs)
    got = ctx.median_blur_11(dev(batch if ch == 3 else batch[..., 0][..., None])).cpu().numpy()
    for i, im in enumerate(imgs):
        want = cv2.medianBlur(im if ch == 3 else im[..., 0], 11)
        assert np.array_equal(got[i].reshape(want.shape), want), (shape, i)


def test_median_blur_panorama_size(ctx):
    rng = np.random.default_rng(3)
    pano = cv2.GaussianBlur(rng.integers(0, 256, (849, 2400, 3), dtype=np.uint8), (0, 0), 1.0)
    got = ctx.median_blur_11(dev(pano)).cpu().numpy()
    assert np.array_equal(got, cv2.medianBlur(pano, 11))


@pytest.mark.parametrize("shape", [(97, 131, 3), (40, 33, 3), (849, 2400, 3)])
def test_median_blur_gray_fused(ctx, shape):
    """medianBlur + cvtColor(BGR2GRAY) in one kernel (camera_models.py:1708-1711): both outputs bit-exact."""
    rng = np.random.default_rng(sum(shape) + 1)
    imgs = np.stack([rng.integers(0, 256, shape, dtype=np.uint8),
                     cv2.GaussianBlur(rng.integers(0, 256, shape, dtype=np.uint8), (0, 0), 1.5)])
    gray, bgr = ctx.median_blur_11_gray(dev(imgs), want_bgr=True)
    gray_only = ctx.median_blur_11_gray(dev(imgs))
    for i in range(2):
        want = cv2.medianBlur(imgs[i], 11)
        assert np.array_equal(bgr[i].cpu().numpy(), want)
        assert np.array_equal(gray[i].cpu().numpy(), cv2.cvtColor(want, cv2.COLOR_BGR2GRAY))
    assert torch.equal(gray, gray_only)
