"""CPU-side checks of the C ABI: the in-tree shared library loads and exports every function that
include/sph_b200.h declares (no compute calls -- there is no GPU here), and the binding table in
sph_taichi_b200/_lib.py covers exactly that set."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sph_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sph_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sph_taichi_b200 import _lib
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    from sph_taichi_b200 import _lib
    assert sorted(_lib.ABI_SYMBOLS) == _declared()
    lib = _lib.load()
    for n in _lib.ABI_SYMBOLS:
        assert getattr(lib, n).argtypes is not None, n


def test_no_cuda_device_fails_loudly():
    """Without a GPU the product path must raise, never fall back to a CPU implementation."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sph_taichi_b200 import ParticleSystem, SimConfig, scene
    with pytest.raises(RuntimeError, match="CUDA"):
        ParticleSystem(SimConfig(scene.cube_8k()))
    from sph_taichi_b200 import _lib
    lib = _lib.load()
    p = _lib.SphParams()
    p.dim = 3
    p.grid_num = (ctypes.c_int32 * 3)(8, 8, 8)
    p.h = 0.04
    p.density0 = 1000.0
    ws = ctypes.create_string_buffer(1 << 20)
    addr = (ctypes.addressof(ws) + 255) // 256 * 256
    h = ctypes.c_void_p()
    rc = lib.sph_create(ctypes.byref(p), 16, 0, 0, 0, ctypes.c_void_p(addr), 1 << 19, ctypes.byref(h))
    assert rc == -2 and b"no CUDA device" in lib.sph_last_error(None)


def test_product_package_never_imports_the_oracle():
    import re
    pkg = os.path.join(ROOT, "sph_taichi_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|sph_oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f"{f} reaches into oracle/"


def test_capacity_limit_is_reported_before_any_device_work():
    """Neighbour-list slots are 32-bit: n_max beyond 2^32 / 96 is refused with SPH_E_CAPACITY (-3)."""
    from sph_taichi_b200 import _lib
    lib = _lib.load()
    p = _lib.SphParams()
    p.dim = 3
    p.grid_num = (ctypes.c_int32 * 3)(8, 8, 8)
    p.h = 0.04
    p.density0 = 1000.0
    h = ctypes.c_void_p()
    ok_n, bad_n = 44_000_000, 45_000_000
    assert lib.sph_workspace_bytes(ctypes.byref(p), ok_n, 0, 0) > 96 * 4 * ok_n
    rc = lib.sph_create(ctypes.byref(p), bad_n, 0, 0, 0, None, 0, ctypes.byref(h))
    assert rc < 0 and b"32-bit" in lib.sph_last_error(None), lib.sph_last_error(None)
    assert not h.value


def test_struct_layouts_match_the_header(tmp_path):
    """Every struct that crosses the ABI: size and field offsets of the ctypes mirror in _lib.py equal what a C
    compiler makes of include/sph_b200.h (a silent mismatch would scramble parameters, not fail)."""
    import subprocess
    from sph_taichi_b200 import _lib
    structs = ["SphParams", "SphFields", "SphRigidBody", "SphTransport", "SphDfsphStep"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sph_b200.h"', "int main(void) {"]
    for sn in structs:
        st = getattr(_lib, sn)
        lines.append(f'  printf("{sn} %zu\\n", sizeof({sn}));')
        for fn, *_ in st._fields_:
            lines.append(f'  printf("{sn}.{fn} %zu\\n", offsetof({sn}, {fn}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for sn in structs:
        st = getattr(_lib, sn)
        assert ctypes.sizeof(st) == int(out[sn]), sn
        for fn, *_ in st._fields_:
            assert getattr(st, fn).offset == int(out[f"{sn}.{fn}"]), f"{sn}.{fn}"
