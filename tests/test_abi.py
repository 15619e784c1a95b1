"""The C-ABI surface (include/cpuvox_b200.h): the library loads, exports every declared symbol, struct layouts match the
reference's blittable types, and — without a GPU — the rendering entry points fail loudly instead of falling back."""
from __future__ import annotations

import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu


def declared_functions():
    text = open(os.path.join(ROOT, "include", "cpuvox_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cvx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(cv):
    from cpuvox_b200 import native
    names = declared_functions()
    assert len(names) >= 40
    lib = C.CDLL(native.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cpuvox_b200.h but not exported"
    assert sorted(native.SYMBOLS) == names, "cpuvox_b200/native.py must bind exactly the header's functions"


def test_struct_layouts(cv):
    from cpuvox_b200 import native as N
    assert C.sizeof(N.Segment) == 36           # RenderManager.SegmentData: 4 x float2 + int (RenderManager.cs:503-510)
    assert C.sizeof(N.Camera) == 16 * 4 + 8 + 4 + 4 + 4 + 6 * 4
    assert C.sizeof(N.FrameSetup) == 4 * 36 + C.sizeof(N.Camera) + 8
    assert C.sizeof(N.Counters) == 48
    assert C.sizeof(N.RayState) == 18 * 4
    assert C.sizeof(N.Pose) == 12 * 4


def test_no_oracle_in_the_product():
    """The product path never links, imports or calls the oracle."""
    import subprocess
    from cpuvox_b200 import native
    deps = subprocess.run(["ldd", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps
    for root, _, files in os.walk(os.path.join(ROOT, "cpuvox_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"


@pytest.mark.skipif(has_gpu(), reason="checks the no-device failure mode")
def test_create_fails_loudly_without_a_device(cv):
    from cpuvox_b200 import native as N
    ctx = C.c_void_p()
    cfg = N.Config(0, 0)
    rc = N.lib.cvx_create(C.byref(cfg), C.byref(ctx))
    assert rc == -2 and not ctx.value  # CVX_ERR_NO_DEVICE
    assert b"no CPU fallback" in N.lib.cvx_last_error(None)
    with pytest.raises(cv.CvxError):
        cv.RenderManager(0)


def test_argument_errors_do_not_crash(cv):
    from cpuvox_b200 import native as N
    assert N.lib.cvx_create(None, None) == -1
    assert N.lib.cvx_set_resolution(None, 10, 10) == -1
    assert N.lib.cvx_draw(None, None) == -1
    assert N.lib.cvx_world_upload(None, 0, 1, 1, 1, None, 0, 0) == -1
    assert N.lib.cvx_destroy(None) == 0
    assert N.lib.cvx_launch_count(None) == 0
