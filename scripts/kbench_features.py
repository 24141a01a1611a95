#!/usr/bin/env python
"""Feature front at C2 scale on one GPU: 16 frames x 2 panoramas of 849 x 2400 x 3, 12 azimuthal masks,
667 corners per mask -> median 11x11, BGR2GRAY, Shi-Tomasi per mask, ORB description.  Per-launch CUDA-event times and
the same work done by cv2 on the host for ONE panorama (the reference's path, camera_models.py:1706-1768)."""
import json
import os
import sys
import time

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vo_single_camera_sos_b200 import ops  # noqa: E402


def main():
    ctx = ops.Context(0)
    rng = np.random.default_rng(0)
    n_img, H, W, n_masks, N = 32, 849, 2400, 12, 667
    base = [cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), 1.5) for _ in range(4)]
    pano = torch.from_numpy(np.stack([base[i % 4] for i in range(n_img)])).cuda()
    masks = np.zeros((n_masks, H, W), np.uint8)
    for m in range(n_masks):
        masks[m, 20:-20, m * W // n_masks:(m + 1) * W // n_masks] = 255
    md = torch.from_numpy(masks).cuda()

    def run():
        med = ctx.median_blur_11(pano)
        gray = ctx.bgr_to_gray(med)
        xy, cnt = ctx.gft_detect(gray, md, N)
        pts = xy.reshape(n_img, n_masks * N, 2)
        idx = torch.arange(n_img, device="cuda", dtype=torch.int32).repeat_interleave(n_masks * N)
        desc, keep = ctx.orb_describe(gray, pts.reshape(-1, 2), None, idx)
        return cnt, desc, keep

    run()
    torch.cuda.synchronize()
    ctx.profile_begin()
    cnt, desc, keep = run()
    marks = ctx.profile_end()
    agg = {}
    for name, t in marks:
        agg[name] = agg.get(name, 0.0) + t
    total = sum(agg.values())
    t0 = time.perf_counter()
    b = cv2.cvtColor(cv2.medianBlur(base[0], 11), cv2.COLOR_BGR2GRAY)
    t1 = time.perf_counter()
    orb = cv2.ORB_create(nfeatures=N)
    nk = 0
    for m in range(n_masks):
        p = cv2.goodFeaturesToTrack(image=b, maxCorners=N, qualityLevel=0.01, minDistance=5, mask=masks[m], useHarrisDetector=False)
        k, d = orb.compute(b, list(cv2.KeyPoint_convert(p.reshape(-1, 2))))
        nk += len(k)
    t2 = time.perf_counter()
    print(json.dumps({"kernel": "feature_front", "panoramas": n_img, "ms_total": total, "ms_by_entry_point": {k: round(v, 3) for k, v in agg.items()},
                      "gft_launches_ms": [round(t, 3) for name, t in marks if name.startswith("sos_gft")],
                      "panoramas_per_s": n_img / (total * 1e-3), "corners": int(cnt.sum()), "described": int(keep.sum()),
                      "cv2_one_panorama_ms": {"median+gray": (t1 - t0) * 1e3, "gft+orb x12 masks": (t2 - t1) * 1e3, "keypoints": nk},
                      "cv2_threads": cv2.getNumThreads()}))


if __name__ == "__main__":
    main()
