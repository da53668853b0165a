"""C-ABI boundary checks that need no GPU: the shared library loads, exports every symbol
include/skmp_b200.h declares, validates arguments, compiles plans on the host, and FAILS LOUDLY
(no CPU fallback) when a compute entry point is called without a CUDA device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "skmp_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"SKB_API[^;(]*?\b(skb_\w+)\s*\(", text)))


def has_cuda():
    import torch

    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    from skmp_b200 import _lib

    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/skmp_b200.h but not exported"
    # and the ctypes binding covers the whole header
    assert set(names) <= set(_lib.EXPORTED_SYMBOLS), set(names) - set(_lib.EXPORTED_SYMBOLS)
    assert L.skb_version() >= 100


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C (no C++ / torch types)."""
    import subprocess

    src = tmp_path / "t.c"
    src.write_text('#include "skmp_b200.h"\nint main(void){ return SKB_MODE_CLOSEST + SKB_F64 > 0 ? 0 : 1; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", f"-I{HEADER.parent}", str(src)], check=True)


def _chain_model(L, n=4):
    parent = np.array([-1] + list(range(n - 1)), dtype=np.int32)
    jtype = np.array([0] + [1] * (n - 1), dtype=np.int32)
    origin = np.zeros((n, 6)); origin[1:, 0] = 0.3
    axis = np.tile([0.0, 0.0, 1.0], (n, 1))
    h = C.c_void_p()
    from skmp_b200 import _lib

    _lib.check(L.skb_model_create(n, _lib.dptr(parent, C.c_int32), _lib.dptr(jtype, C.c_int32), _lib.dptr(origin), _lib.dptr(axis), C.byref(h)))
    return h


def test_argument_validation_and_error_messages():
    from skmp_b200 import _lib

    L = _lib.lib()
    h = C.c_void_p()
    bad_parent = np.array([0, 0], dtype=np.int32)
    z6, z3, jt = np.zeros((2, 6)), np.zeros((2, 3)), np.zeros(2, dtype=np.int32)
    assert L.skb_model_create(2, _lib.dptr(bad_parent, C.c_int32), _lib.dptr(jt, C.c_int32), _lib.dptr(z6), _lib.dptr(z3), C.byref(h)) != 0
    assert b"root" in L.skb_last_error()
    m = _chain_model(L)
    n = C.c_int32()
    _lib.check(L.skb_model_num_links(m, C.byref(n)))
    assert n.value == 4
    lid = C.c_int32()
    xyz = np.array([0.1, 0.0, 0.0])
    _lib.check(L.skb_model_add_link(m, 3, _lib.dptr(xyz), None, C.byref(lid)))
    assert lid.value == 4
    assert L.skb_model_add_link(m, 99, _lib.dptr(xyz), None, C.byref(lid)) != 0
    # plans compile on the host: dims for every base type
    feats = np.array([4], dtype=np.int32)
    joints = np.array([1, 2, 3], dtype=np.int32)
    for base, extra in ((0, 0), (1, 3), (2, 6)):
        for rot, t in ((0, 3), (1, 6), (2, 7)):
            p = C.c_void_p()
            _lib.check(L.skb_plan_create(m, _lib.dptr(feats, C.c_int32), 1, _lib.dptr(joints, C.c_int32), 3, base, rot, C.byref(p)))
            d, f, tt = C.c_int32(), C.c_int32(), C.c_int32()
            _lib.check(L.skb_plan_dims(p, C.byref(d), C.byref(f), C.byref(tt)))
            assert (d.value, f.value, tt.value) == (3 + extra, 1, t)
            L.skb_plan_destroy(p)
    dup = np.array([1, 1], dtype=np.int32)
    p = C.c_void_p()
    assert L.skb_plan_create(m, _lib.dptr(feats, C.c_int32), 1, _lib.dptr(dup, C.c_int32), 2, 0, 0, C.byref(p)) != 0
    assert b"duplicated" in L.skb_last_error()
    assert L.skb_plan_create(m, _lib.dptr(np.array([77], dtype=np.int32), C.c_int32), 1, _lib.dptr(joints, C.c_int32), 3, 0, 0, C.byref(p)) != 0
    L.skb_model_destroy(m)
    # SDF descriptors: argument checks that do not touch the device
    s = C.c_void_p()
    _lib.check(L.skb_sdf_create(C.byref(s)))
    half = np.array([0.1, 0.2, 0.3])
    _lib.check(L.skb_sdf_add_box(s, None, None, _lib.dptr(half)))
    _lib.check(L.skb_sdf_add_sphere(s, _lib.dptr(xyz), 0.2))
    _lib.check(L.skb_sdf_num_prims(s, C.byref(n)))
    assert n.value == 2
    dims = np.array([1, 4, 4], dtype=np.int32)
    vals = np.zeros(16, dtype=np.float32)
    assert L.skb_sdf_add_grid(s, None, None, _lib.dptr(xyz), _lib.dptr(xyz), _lib.dptr(dims, C.c_int32), 1.0, C.c_void_p(vals.ctypes.data), 0) != 0
    L.skb_sdf_destroy(s)


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_compute_without_gpu_fails_loudly():
    """north_star: no CPU fallback.  Without a device every compute path raises."""
    from skmp_b200 import SkbError
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.sdf import Box

    pr2 = PR2(); pr2.reset_manip_pose()
    colkin = PR2Config().get_collision_kin()
    const = CollFreeConst(colkin, Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9]), pr2)
    with pytest.raises(SkbError):
        const.evaluate(np.zeros((4, 7)), True)
    with pytest.raises(SkbError):
        colkin.map(np.zeros((4, 7)))
    with pytest.raises(SkbError):
        const.is_valid(np.zeros(7))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under scikit-motionplan_b200/ may reference it."""
    pkg = ROOT / "scikit-motionplan_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cuh")):
        text = path.read_text(errors="ignore")
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
        assert "libskmp_oracle" not in text and not re.search(r"#include\s*[<\"].*oracle", text), path
