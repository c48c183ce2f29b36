"""CPU-side checks of the C-ABI library: it builds for sm_100a, loads, exports every symbol
include/evosops.h declares, and its pure-host helpers agree with the oracle.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import evosops_b200 as E
from evosops_b200 import capi
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "evosops.h")).read()
    declared = set(re.findall(r"\b(sops_[a-z_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS)
    L = capi.lib()
    for name in declared:
        assert getattr(L, name) is not None


def test_struct_layouts_match_header():
    assert C.sizeof(capi.SopsSize) == 6
    assert capi.RESULT_DTYPE.itemsize == 24 and capi.RESULT_DTYPE.fields["norm"][1] == 16
    assert O.RESULT_DTYPE == capi.RESULT_DTYPE


@pytest.mark.parametrize("beh,size", [(E.AGG, (6, 3)), (E.AGG, (13, 9)), (E.AGG, (26, 18)), (E.SEP, (13, 9)),
                                      (E.COAT, (20, 5, 2)), (E.COAT, (26, 8, 4)), (E.LOCO, (13, 9))])
def test_instance_shape_matches_oracle(beh, size):
    G, n, dur = E.instance_shape(beh, size)
    env = O.Env(beh, np.zeros(E.GENOME_LEN[beh], dtype=np.uint8), *size, seed=1)
    assert (G, n, dur) == (env.G, env.n, env.sim_duration)


@pytest.mark.parametrize("beh,size", [(E.AGG, (3, 6)), (E.AGG, (0, 0)), (E.AGG, (128, 3)), (E.COAT, (6, 3, 2))])
def test_invalid_sizes_are_errors_not_hangs(beh, size):
    with pytest.raises(E.SopsError) as ei:
        E.instance_shape(beh, size)
    assert ei.value.code == capi.E_SIZE


@pytest.mark.parametrize("size", [(20, 5, 2), (26, 8, 4), (12, 2, 1)])
def test_coat_distance_grid_matches_reference_per_cell_bfs(size):
    # the library uses one multi-source BFS; the oracle restates coating.rs:101-129 literally (BFS per cell)
    env = O.Env(O.COAT, np.zeros(600, dtype=np.uint8), *size, seed=1)
    assert np.array_equal(E.coat_distance_grid(size), env.dist_grid())


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(E.SopsError) as ei:
        E.Context()
    assert ei.value.code == capi.E_CUDA


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "evosops_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_rust_binding_source_matches_header():
    # rust/evosops-ffi cannot be compiled here (no cargo); keep its extern block in sync with the header by text
    hdr = open(os.path.join(ROOT, "include", "evosops.h")).read()
    rs = open(os.path.join(ROOT, "rust", "evosops-ffi", "src", "lib.rs")).read()
    ext = rs[rs.index('extern "C" {'):]
    ext = ext[:ext.index("\n}\n")]
    for name, args in re.findall(r"fn (sops_[a-z_]+)\(([^;]*)\)", ext, flags=re.S):
        m = re.search(r"\b%s\s*\(([^;]*)\);" % name, hdr, flags=re.S)
        assert m, name
        assert len([a for a in args.split(",") if a.strip()]) == len([a for a in m.group(1).split(",") if a.strip()]), name
    for flag, val in re.findall(r"pub const (F_[A-Z_]+): u32 = (\d+);", rs):
        assert re.search(r"SOPS_%s = %s\b" % (flag, val), hdr), flag
    assert "pub arena: u16, pub layers: u16, pub coat: u16" in rs
    assert "pub i0: u32, pub i1: u32, pub i2: u32, pub f: f32, pub norm: f64" in rs
