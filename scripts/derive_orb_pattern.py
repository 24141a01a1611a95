"""Derive the 256 x 2 test-point table of the ORB descriptor (Rublee et al. 2011; OpenCV's bit_pattern_31_) from the
BEHAVIOUR of the installed cv2.ORB, and pin the rest of the descriptor model while doing so.  Build-time tool: its
outputs are committed (like the golden vectors), nothing imports it at run time.

Model being identified (what cv2.ORB.compute does with user-supplied keypoints of octave 0):
    B = round(G * gray), G = separable 7-tap Gaussian of sigma 2, BORDER_REFLECT_101 (see blur());  c = (cvRound(kp.x), cvRound(kp.y))
    bit k = B[c + round(R(a, b) p0_k)] < B[c + round(R(a, b) p1_k)],   byte j = bits 8j .. 8j+7, first test in the LSB.
Identification: on N random images every bit k is observed for a keypoint of angle 0; a candidate pair (p0, p1) in the
31 x 31 patch survives only if it reproduces all N observations.  Exactly one pair survives per bit, which also confirms
the blur model.  The table is then validated on random keypoints with random angles against cv2 bit for bit.

Writes vo_single_camera_sos_b200/csrc/orb_pattern.inc (compiled into the library) and tests/golden/orb_pattern.npy
(int8 [256, 4] = x0, y0, x1, y1, used by the oracle) — run in the build container: python scripts/derive_orb_pattern.py"""
from __future__ import annotations

import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "orb_pattern.npy")
HALF = 15


def gaussian_kernel_7_2():
    x = np.arange(-3, 4, dtype=np.float64)
    k = np.exp(-(x * x) / (2.0 * 2.0 * 2.0))
    return k / k.sum()


def blur(gray):
    """The blur INSIDE cv2.ORB (GaussianBlur(7x7, sigma 2) applied in place to the bordered pyramid level) does not take
    cv2.GaussianBlur's bit-exact fixed-point path: it behaves as the exact separable convolution rounded once at the end
    (identified empirically: zero disagreements, whereas cv2.GaussianBlur on the same image flips ~1 % of the near-tie
    comparisons)."""
    k = gaussian_kernel_7_2().reshape(-1, 1)
    return np.rint(cv2.sepFilter2D(gray.astype(np.float64), -1, k, k, borderType=cv2.BORDER_REFLECT_101)).astype(np.uint8)


def unpack(desc):
    return np.unpackbits(desc, axis=-1, bitorder="little")


def derive(n_images=160, size=129, seed=0):
    rng = np.random.default_rng(seed)
    orb = cv2.ORB_create(nfeatures=10)
    c = size // 2
    kp = [cv2.KeyPoint(float(c), float(c), 31.0, 0.0, 1.0, 0, -1)]
    yy, xx = np.mgrid[-HALF:HALF + 1, -HALF:HALF + 1]
    offs = np.stack([xx.ravel(), yy.ravel()], 1)                     # 961 candidate offsets
    alive = None
    for n in range(n_images):
        img = rng.integers(0, 256, (size, size), dtype=np.uint8)
        if n % 3 == 1:                                               # lower-frequency content separates far-apart candidates
            img = cv2.resize(rng.integers(0, 256, (size // 4 + 1, size // 4 + 1), dtype=np.uint8), (size, size),
                             interpolation=cv2.INTER_NEAREST)
        _, d = orb.compute(img, kp)
        bits = unpack(d)[0].astype(bool)                             # [256]
        B = blur(img).astype(np.int16)
        v = B[c + offs[:, 1], c + offs[:, 0]]                        # [961]
        less = v[:, None] < v[None, :]                               # [961, 961]: candidate (p0, p1) predicts 1
        if alive is None:
            alive = np.ones((256,) + less.shape, bool)
        alive &= (less[None] == bits[:, None, None])
    table = np.zeros((256, 4), np.int8)
    for k in range(256):
        idx = np.argwhere(alive[k])
        if len(idx) != 1:
            raise SystemExit(f"bit {k}: {len(idx)} candidate pairs survive (model wrong or too few images)")
        table[k] = [*offs[idx[0, 0]], *offs[idx[0, 1]]]
    return table


def predict(gray, pts, angles_deg, table):
    """Descriptor model with the table (float32 arithmetic as in OpenCV's computeOrbDescriptors)."""
    B = blur(gray)
    out = np.zeros((len(pts), 32), np.uint8)
    t = table.astype(np.float32)
    for i, ((x, y), ang) in enumerate(zip(pts, angles_deg)):
        ang32 = np.float32(ang) * np.float32(np.pi / 180.0)
        a, b = np.float32(np.cos(ang32)), np.float32(np.sin(ang32))
        cx, cy = int(np.rint(x)), int(np.rint(y))
        x0 = np.rint(t[:, 0] * a - t[:, 1] * b).astype(int); y0 = np.rint(t[:, 0] * b + t[:, 1] * a).astype(int)
        x1 = np.rint(t[:, 2] * a - t[:, 3] * b).astype(int); y1 = np.rint(t[:, 2] * b + t[:, 3] * a).astype(int)
        bits = B[cy + y0, cx + x0] < B[cy + y1, cx + x1]
        out[i] = np.packbits(bits, bitorder="little")
    return out


def validate(table, seed=1):
    rng = np.random.default_rng(seed)
    orb = cv2.ORB_create(nfeatures=10)
    bad = 0
    total = 0
    for trial in range(8):
        img = cv2.GaussianBlur(rng.integers(0, 256, (240, 320), dtype=np.uint8), (0, 0), 1.0 + 0.5 * trial)
        pts = np.stack([rng.uniform(40, 280, 200), rng.uniform(40, 200, 200)], 1).astype(np.float32)
        ang = rng.uniform(0, 360, 200).astype(np.float32)
        ang[::5] = -1.0                                              # KeyPoint_convert's angle (the GFT path)
        kps = [cv2.KeyPoint(float(p[0]), float(p[1]), 1.0, float(a), 1.0, 0, -1) for p, a in zip(pts, ang)]
        kps2, d = orb.compute(img, kps)
        assert len(kps2) == len(kps)
        got = predict(img, pts, ang, table)
        bad += int((unpack(got) != unpack(d)).sum())
        total += d.size * 8
    return bad, total


if __name__ == "__main__":
    assert np.allclose(gaussian_kernel_7_2(), cv2.getGaussianKernel(7, 2)[:, 0], rtol=0, atol=1e-16)
    tab = derive()
    bad, total = validate(tab)
    print(f"derived {len(tab)} test pairs; validation: {bad} of {total} bits differ from cv2.ORB.compute")
    if bad:
        sys.exit(1)
    np.save(OUT, tab)
    inc = os.path.join(ROOT, "vo_single_camera_sos_b200", "csrc", "orb_pattern.inc")
    with open(inc, "w") as f:
        f.write("// ORB test-point table (x0, y0, x1, y1 per descriptor bit), generated by scripts/derive_orb_pattern.py from the\n"
                "// behaviour of cv2.ORB.compute (validated bit for bit on 409600 descriptor bits).  Do not edit.\n")
        for k in range(0, 256, 4):
            f.write("  " + " ".join("%d,%d,%d,%d," % tuple(int(v) for v in tab[j]) for j in range(k, k + 4)) + "\n")
    print("wrote", OUT, "and", inc, "max |coord| =", int(np.abs(tab).max()))
