"""CPU oracle for the SOS front-end hot path — TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy (plus the same OpenCV calls the reference makes), the algorithms of the five
hot-path steps of ubuntuslave/vo_single_camera_sos.  Every function cites the reference file:line it follows.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import it;
the product package `vo_single_camera_sos_b200` never does (tests/test_no_oracle_in_product.py enforces that).

Pinning status:
  * remap, Hamming matching, pixel gates, panorama lifting, GUM projection / lifting, midpoint triangulation,
    range gates, RGB-D back-projection, Arun/Kabsch: PINNED — `oracle/gen_golden.py` ran the reference's own
    functions (imported from /root/reference under four harness shims) on seeded inputs and the outputs are
    committed under tests/golden/; tests/test_oracle_golden.py checks this package against them.
  * RANSAC loop (hypothesis sampling, scoring, argmax): PARITY UNPINNED against OpenGV — pyopengv is neither
    vendored nor installed and the reference pins no RANSAC output.  The loop is a float64 restatement of the
    semantics written down in include/sosfront.h, built from the pinned Arun solver and the reference's own score
    function (pose_est_tools.py:150-203).
"""
from . import geometry, hamming, ransac, remap  # noqa: F401
