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


C_TO_RUST = {          # C parameter type (whitespace-normalised) -> the Rust FFI type that has the same ABI
    "sops_ctx**": "*mut *mut SopsCtx", "sops_ctx*": "*mut SopsCtx", "const sops_ctx*": "*const SopsCtx",
    "const int*": "*const c_int", "int": "c_int", "const uint8_t*": "*const u8", "uint8_t*": "*mut u8", "uint8_t": "u8",
    "uint32_t": "u32", "uint32_t*": "*mut u32", "uint64_t": "u64", "uint64_t*": "*mut u64", "const uint64_t*": "*const u64",
    "float": "f32", "double*": "*mut f64", "const sops_size*": "*const SopsSize", "sops_size": "SopsSize",
    "sops_result*": "*mut SopsResult", "uint16_t*": "*mut u16", "uint64_t[6]": "*mut u64",
}
C_RET = {"int": "c_int", "void": None, "const char*": "*const c_char"}


def _c_params(text):
    out = []
    for a in text.split(","):
        a = re.sub(r"/\*.*?\*/", "", a, flags=re.S).strip()
        m = re.match(r"^(.*?)([A-Za-z_][A-Za-z_0-9]*)(\[\d+\])?$", a)      # type, name, optional array suffix
        typ = re.sub(r"\s+", " ", m.group(1)).strip().replace(" *", "*") + (m.group(3) or "")
        out.append(typ)
    return out


def test_rust_binding_source_matches_header():
    # rust/evosops-ffi cannot be compiled here (no cargo); its extern block is checked against the header by text:
    # every bound function exists, with the same number of parameters AND the ABI-equivalent type in every position,
    # and the same return type; flag constants and struct layouts are compared too
    hdr = open(os.path.join(ROOT, "include", "evosops.h")).read()
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    rs = open(os.path.join(ROOT, "rust", "evosops-ffi", "src", "lib.rs")).read()
    ext = rs[rs.index('extern "C" {'):]
    ext = ext[:ext.index("\n}\n")]
    bound = re.findall(r"fn (sops_[a-z_]+)\(([^;]*?)\)\s*(?:->\s*([^;]+?))?;", ext, flags=re.S)
    assert len(bound) >= 7
    for name, args, ret in bound:
        m = re.search(r"([A-Za-z_][A-Za-z_0-9 ]*\*?)\s*\b%s\s*\(([^;]*)\);" % name, hdr_nc, flags=re.S)
        assert m, name
        c_ret = re.sub(r"\s+", " ", m.group(1)).strip().replace(" *", "*")
        assert C_RET[c_ret] == (ret.strip() if ret else None), (name, c_ret, ret)
        c_types = _c_params(m.group(2))
        r_types = [re.sub(r"\s+", " ", a.split(":", 1)[1]).strip() for a in args.split(",") if a.strip()]
        assert len(c_types) == len(r_types), name
        for k, (ct, rt) in enumerate(zip(c_types, r_types)):
            assert C_TO_RUST[ct] == rt, "%s parameter %d: C `%s` is bound as Rust `%s`" % (name, k, ct, rt)
    for flag, val in re.findall(r"pub const (F_[A-Z_]+): u32 = (\d+);", rs):
        assert re.search(r"SOPS_%s = %s\b" % (flag, val), hdr), flag
    for flag, val in re.findall(r"SOPS_(F_[A-Z_]+) = (\d+)", hdr):
        assert re.search(r"pub const %s: u32 = %s;" % (flag, val), rs), "flag %s missing from the Rust binding" % flag
    assert "pub arena: u16, pub layers: u16, pub coat: u16" in rs
    assert "pub i0: u32, pub i1: u32, pub i2: u32, pub f: f32, pub norm: f64" in rs
