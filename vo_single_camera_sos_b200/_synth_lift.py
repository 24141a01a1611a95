"""NumPy GUM pixel -> viewing-ray lifting used ONLY to render synthetic omni images (synth.render_omni) when no device
lifting callback is supplied.  Same geometry as sos_lift_gum (gum.py:2653-2762 of the reference)."""
import numpy as np


def gum_lift(g: dict, uv: np.ndarray) -> np.ndarray:
    u, v = uv[:, 0], uv[:, 1]
    xd = (u - g["u_center"]) / g["gamma1"] - g["alpha_c"] * (v - g["v_center"]) / g["gamma2"]
    yd = (v - g["v_center"]) / g["gamma2"]
    r2 = xd * xd + yd * yd
    if g["use_distortion"]:
        if g["l1"] != 0:
            f = 1 + g["l1"] * r2 + g["l2"] * r2 ** 2 + g["l3"] * r2 ** 3
            xu, yu = xd * f, yd * f
        else:
            r4 = r2 * r2
            rad = g["k1"] * r2 + g["k2"] * r4
            dx = xd * rad + g["p2"] * (r2 + 2 * xd * xd) + 2 * g["p1"] * xd * yd
            dy = yd * rad + g["p1"] * (r2 + 2 * yd * yd) + 2 * g["p2"] * xd * yd
            inv = 1 / (1 + 4 * g["k1"] * r2 + 6 * g["k2"] * r4 + 8 * g["p1"] * yd + 8 * g["p2"] * xd)
            xu, yu = xd - inv * dx, yd - inv * dy
    else:
        xu, yu = xd, yd
    cp = np.array([g["xi1"], g["xi2"], g["xi3"]])
    p = np.stack([cp[0] + xu, cp[1] + yu, np.full_like(xu, g["plane_k"])], 1)
    d = p - cp
    a = np.sum(d * d, 1)
    b = 2 * np.sum(d * p, 1)
    c = np.sum(p * p, 1) - 1
    t = (-b + np.sqrt(np.maximum(b * b - 4 * a * c, 0))) / (2 * a)
    return p + t[:, None] * d
