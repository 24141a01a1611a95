#!/usr/bin/env python
"""Probe what bounds the remap kernel: vary source residency (batch size) and the gather pattern (annulus vs identity)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from vo_single_camera_sos_b200 import ops
from scripts.kbench import timeit

def run(ctx, B, pattern, nsrc=2, H=2048, W=2048, rows=849, cols=2400):
    g = torch.Generator(device="cuda").manual_seed(2)
    srcs = [torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(nsrc)]
    r = torch.arange(rows, device="cuda", dtype=torch.float64)[:, None]
    c = torch.arange(cols, device="cuda", dtype=torch.float64)[None, :]
    luts = []
    for v, (lo, hi) in enumerate(((0.13 * H, 0.47 * H), (0.03 * H, 0.12 * H))):
        if pattern == "annulus":
            rad = lo + (hi - lo) * r / rows; ang = 2 * np.pi * (1 - c / cols)
            mx = (W / 2 + rad * torch.cos(ang)).contiguous(); my = (H / 2 + rad * torch.sin(ang)).contiguous()
        else:  # identity-like: panorama pixel (r, c) reads source (r + 100 v, c * 0.8 + 3.3): streaming rows
            mx = (c * 0.8 + 3.3 + 0 * r).contiguous(); my = (r + 100.25 * (v + 1) + 0 * c).contiguous()
        luts.append(ctx.lut_pack(mx, my, (H, W)))
    lut = torch.stack(luts).contiguous()
    out = ctx.remap(srcs[0], lut)
    it = [0]
    def f():
        ctx.remap(srcs[it[0] % nsrc], lut, out=out); it[0] += 1
    ms = timeit(f)
    return dict(B=B, pattern=pattern, nsrc=nsrc, ms=ms, us_per_frame=ms * 1e3 / B, src_MB=B * H * W * 3 / 1e6,
                gbps_src_dst=B * (2 * rows * cols * 3 + H * W * 3) / (ms * 1e-3) / 1e9)

ctx = ops.Context(0)
for B, pat, nsrc in ((16, "annulus", 2), (2, "annulus", 1), (16, "identity", 2), (2, "identity", 1), (64, "annulus", 1)):
    print(json.dumps(run(ctx, B, pat, nsrc)))
