"""CPU: the product package never imports, links or executes anything under oracle/ (or the reference)."""
import os
import re

from conftest import ROOT


def test_product_never_touches_oracle():
    pkg = os.path.join(ROOT, "vo_single_camera_sos_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                continue
            text = open(os.path.join(d, f), errors="replace").read()
            if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "/root/reference" in text or "oracle/" in text:
                bad.append(os.path.join(d, f))
    assert not bad, bad
