"""CPU-side checks of the drop-in boundary: libls1b200.so builds for sm_100a, loads without a GPU, exports every symbol
include/ls1b200.h declares, and refuses to work without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from ls1_mardyn_b200 import build
    return build.build()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ls1b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ls1b200_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for s in ("ls1b200_create", "ls1b200_update", "ls1b200_traverse_cells", "ls1b200_leapfrog_pre", "ls1b200_leapfrog_post",
              "ls1b200_simulate", "ls1b200_halo_export", "ls1b200_halo_import", "ls1b200_decomp_attach", "ls1b200_exchange"):
        assert s in syms


def test_library_exports_every_declared_symbol(libpath):
    lib = C.CDLL(libpath)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    from ls1_mardyn_b200 import capi
    assert sorted(capi.EXPORTS) == declared_symbols()
    lib.ls1b200_abi_version.restype = C.c_int
    assert lib.ls1b200_abi_version() == 2


def declared_arity():
    """{symbol: number of parameters} parsed from the prototypes of include/ls1b200.h"""
    hdr = open(os.path.join(ROOT, "include", "ls1b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    out = {}
    for name, params in re.findall(r"\b(ls1b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        params = params.strip()
        out[name] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def test_ctypes_bindings_match_the_prototypes(libpath):
    """Every argtypes list of the ctypes mirror has as many entries as the C prototype has parameters (binding drift shows
    up here, not as a corrupted stack on the GPU box)."""
    from ls1_mardyn_b200 import capi
    lib = capi.lib()
    arity = declared_arity()
    assert set(arity) == set(declared_symbols())
    bad = []
    for name, n in arity.items():
        at = getattr(lib, name).argtypes
        if at is not None and len(at) != n:
            bad.append((name, n, len(at)))
    assert not bad, bad
    # the entry points added in round 1 are bound
    for name in ("ls1b200_peer_handle", "ls1b200_peer_attach", "ls1b200_peer_detach", "ls1b200_download_begin", "ls1b200_download_end"):
        assert getattr(lib, name).argtypes is not None, name


def test_struct_layouts_match_header(libpath):
    from ls1_mardyn_b200 import capi
    from ls1_mardyn_b200.components import CComponent, CSite
    assert C.sizeof(capi.Config) == 8 * 9 + 4 * (2 + 3 + 1 + 4)
    assert C.sizeof(CComponent) == 32 + 32 and C.sizeof(CSite) == 56
    assert C.sizeof(capi.Macro) == 48 and C.sizeof(capi.Kinetic) == 32 and C.sizeof(capi.StepRecord) == 64
    assert C.sizeof(capi.FlopCount) == 128
    assert C.sizeof(capi.Decomp) == 4 * 8 + 8 * 6 + 128 + 16 and capi.PARTICLE.itemsize == 120


def test_no_cpu_fallback(libpath):
    """Without a CUDA device ls1b200_create must fail with a message -- never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ls1_mardyn_b200 import capi
    with pytest.raises(capi.Ls1B200Error, match="no CUDA device"):
        capi.Context([0, 0, 0], [10, 10, 10], 2.5)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under ls1_mardyn_b200/ may reference it."""
    pkg = os.path.join(ROOT, "ls1_mardyn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, os.path.join(dirpath, f)
