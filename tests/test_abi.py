"""CPU: the C-ABI library builds, loads, and exports exactly what include/sosfront.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "sosfront.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sos_[a-z0-9_]+)\s*\(", text)))


def test_header_is_plain_c():
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "sosfront.h"\nint main(void){return SOS_OK;}\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", f"-I{ROOT}/include", "-c", src, "-o", os.path.join(d, "t.o")], check=True)


def test_library_exports_every_declared_symbol():
    from vo_single_camera_sos_b200 import _build, _lib
    path = _build.build()
    assert path.exists()
    names = declared_functions()
    assert len(names) >= 30
    lib = ctypes.CDLL(str(path))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in sosfront.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototype table out of sync with sosfront.h"
    loaded = _lib.load()
    assert loaded.sos_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vo_single_camera_sos_b200 import ops
    with pytest.raises(RuntimeError):
        ops.Context(0)
