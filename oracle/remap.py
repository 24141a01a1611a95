"""Oracle for step 1 (panoramic remap). TEST INFRASTRUCTURE ONLY — see oracle/__init__.py."""
from __future__ import annotations

import cv2
import numpy as np

INT_MIN = -(1 << 31)


def masked_image(img: np.ndarray, mask: np.ndarray | None, background=(0, 0, 0)) -> np.ndarray:
    """dst = mask ? src : background — what OmniStereoModel.get_fully_masked_images produces for one view
    (reference camera_models.py:2990-3010: bitwise_and(img, img, mask=mask), then the background painted through
    the inverted mask)."""
    if mask is None:
        return img
    out = np.empty_like(img)
    bg = np.asarray(background, dtype=img.dtype).reshape(-1)
    if img.ndim == 3:
        out[...] = bg[: img.shape[2]]
        out[mask != 0] = img[mask != 0]
    else:
        out[...] = bg[0]
        out[mask != 0] = img[mask != 0]
    return out


def remap_reference(img: np.ndarray, map_x: np.ndarray, map_y: np.ndarray, border=(0, 0, 0)) -> np.ndarray:
    """The reference's own call: Panorama.get_panoramic_image (panorama.py:291-298) — float64 LUT cast to float32, then
    cv2.remap(INTER_LINEAR, BORDER_CONSTANT)."""
    mx = np.ascontiguousarray(map_x.astype("float32"))
    my = np.ascontiguousarray(map_y.astype("float32"))
    border = tuple(int(b) for b in np.asarray(border).reshape(-1))
    return cv2.remap(img, mx, my, cv2.INTER_LINEAR, None, cv2.BORDER_CONSTANT, border)


def cv_round_q5(x: np.ndarray) -> np.ndarray:
    """cvRound(x * 32) as cv::remap computes it (cvtps2dq): float32 product, round-half-even, and the x86
    'integer indefinite' INT_MIN for NaN / out-of-range."""
    with np.errstate(invalid="ignore", over="ignore"):
        xs = x.astype(np.float32) * np.float32(32.0)
        r = np.rint(xs).astype(np.float64)
        bad = ~(np.abs(xs.astype(np.float64)) < 2147483648.0)
    return np.where(bad, float(INT_MIN), r).astype(np.int64)


def lut_pack_spec(map_x: np.ndarray, map_y: np.ndarray, src_hw, mask: np.ndarray | None = None) -> np.ndarray:
    """Packed LUT entries exactly as include/sosfront.h documents them (sos_lut_entry)."""
    h, w = src_hw
    sx, sy = cv_round_q5(map_x), cv_round_q5(map_y)
    x0 = np.clip(sx >> 5, -32768, 32767)
    y0 = np.clip(sy >> 5, -32768, 32767)
    ax, ay = sx & 31, sy & 31
    inside = np.zeros(sx.shape, np.uint64)
    live = np.zeros(sx.shape, np.uint64)
    for tap in range(4):
        x, y = x0 + (tap & 1), y0 + (tap >> 1)
        ins = (x >= 0) & (x < w) & (y >= 0) & (y < h)
        inside |= ins.astype(np.uint64) << np.uint64(tap)
        if mask is None:
            lv = ins
        else:
            lv = ins & (mask[np.clip(y, 0, h - 1), np.clip(x, 0, w - 1)] != 0)
        live |= lv.astype(np.uint64) << np.uint64(tap)
    p1 = ((y0 + 1) * w + x0) * 3
    wide3 = ((live == 15) & (p1 + 16 <= h * w * 3)).astype(np.uint64)
    lo = (x0 & 0xFFFF).astype(np.uint64) | ((y0 & 0xFFFF).astype(np.uint64) << np.uint64(16))
    hi = (ax.astype(np.uint64) | (ay.astype(np.uint64) << np.uint64(5)) | (inside << np.uint64(16)) | (live << np.uint64(20))
          | (wide3 << np.uint64(24)))
    return lo | (hi << np.uint64(32))


def remap_spec(img: np.ndarray, map_x: np.ndarray, map_y: np.ndarray, border=(0, 0, 0), mask: np.ndarray | None = None,
               background=(0, 0, 0)) -> np.ndarray:
    """NumPy restatement of OpenCV's fixed-point bilinear remap (what cv2.remap does for 8-bit images with float maps):
    Q5 coordinates, Q15 weights (32-ay)(32-ax)*32 ..., (sum + 16384) >> 15, constant border; plus the folded mirror
    mask (a masked-out source pixel reads as `background`).  Bit-exact with remap_reference(masked_image(...))."""
    squeeze = img.ndim == 2
    src = img[..., None] if squeeze else img
    h, w, ch = src.shape
    sx, sy = cv_round_q5(map_x), cv_round_q5(map_y)
    x0 = np.clip(sx >> 5, -32768, 32767)
    y0 = np.clip(sy >> 5, -32768, 32767)
    ax, ay = sx & 31, sy & 31
    wts = [(32 - ay) * (32 - ax) * 32, (32 - ay) * ax * 32, ay * (32 - ax) * 32, ay * ax * 32]
    bd = np.asarray(border, np.int64).reshape(-1)[:ch]
    bg = np.asarray(background, np.int64).reshape(-1)[:ch]
    acc = np.full(sx.shape + (ch,), 16384, np.int64)
    for tap in range(4):
        x, y = x0 + (tap & 1), y0 + (tap >> 1)
        ins = (x >= 0) & (x < w) & (y >= 0) & (y < h)
        xc, yc = np.clip(x, 0, w - 1), np.clip(y, 0, h - 1)
        v = src[yc, xc].astype(np.int64)
        if mask is not None:
            v = np.where((mask[yc, xc] != 0)[..., None], v, bg)
        v = np.where(ins[..., None], v, bd)
        acc += wts[tap][..., None] * v
    out = np.clip(acc >> 15, 0, 255).astype(np.uint8)
    return out[..., 0] if squeeze else out
