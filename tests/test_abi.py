"""The C-ABI library loads on a CPU-only box and exports every symbol the header declares."""

import ctypes
import os
import re

import pytest

from conftest import ROOT
from grf_b200 import _lib


@pytest.fixture(scope="module")
def so_path():
    return _lib.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "grf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(grf_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(so_path):
    lib = ctypes.CDLL(so_path)
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_version_and_pure_host_entry_points(so_path):
    lib = _lib.lib()
    assert lib.grf_abi_version() == _lib.ABI_VERSION == 8
    assert lib.grf_walk_stage_stride(100, 5) == 401
    assert lib.grf_walk_stage_stride(7, 1) == 1
    assert lib.grf_scan_workspace_bytes(0) >= 8
    assert lib.grf_scan_workspace_bytes(10_000_000) >= 8 * (10_000_000 // 2048)


def test_argument_errors_surface_as_value_error(so_path):
    """Invalid arguments are rejected on the host before any CUDA call."""
    lib = _lib.lib()
    g = _lib.GrfGraph(4, 8, None, None, None)
    c = _lib.GrfWalkCfg(0, 4, 0, 3, 0.1, 0, 0, 42, None, None)   # W = 0
    rc = lib.grf_walk(ctypes.byref(g), ctypes.byref(c), 401, None, None, None, None, None)
    assert rc == _lib.GRF_ERR_INVALID
    with pytest.raises(ValueError, match="walks_per_node"):
        _lib.check(rc)
    c = _lib.GrfWalkCfg(0, 9, 5, 3, 0.1, 0, 0, 42, None, None)   # start range outside the graph
    with pytest.raises(ValueError, match="start range"):
        _lib.check(lib.grf_walk(ctypes.byref(g), ctypes.byref(c), 401, None, None, None, None, None))
    c = _lib.GrfWalkCfg(0, 4, 5, 3, 0.1, 1, 0, 42, None, None)   # replay without a trace
    with pytest.raises(ValueError, match="replay"):
        _lib.check(lib.grf_walk(ctypes.byref(g), ctypes.byref(c), 401, None, None, None, None, None))


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from grf_b200.engine import DeviceGraph
    import numpy as np

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeviceGraph(np.zeros(2, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0), 1)


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct in include/grf_b200.h, as a C compiler sees them, equal the
    ctypes mirrors in grf_b200/_lib.py (the header is plain C: it must compile with gcc)."""
    import ctypes
    import shutil
    import subprocess

    from grf_b200 import _lib

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = {"GrfEntry": None, "GrfGraph": _lib.GrfGraph, "GrfWalkCfg": _lib.GrfWalkCfg,
               "GrfLongRows": _lib.GrfLongRows, "GrfPhi": _lib.GrfPhi}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "grf_b200.h"', 'int main(void) {']
    for name, mirror in structs.items():
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        if mirror is not None:
            for field, _ in mirror._fields_:
                lines.append(f'  printf("{name} {field} %zu\\n", offsetof({name}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", _lib.INCLUDE, "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True)
    seen = 0
    for line in out.splitlines():
        name, what, value = line.split()
        mirror = structs[name]
        if mirror is None:
            assert what == "size" and int(value) == 8
            continue
        if what == "size":
            assert ctypes.sizeof(mirror) == int(value), name
        else:
            assert getattr(mirror, what).offset == int(value), (name, what)
        seen += 1
    assert seen == sum(1 + len(m._fields_) for m in structs.values() if m is not None)
