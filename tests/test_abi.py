"""The C-ABI library loads on a CPU-only box and exports every symbol include/cellscopes_cuda.h declares
(no compute call is made here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cellscopes_cuda.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("csc_create", "csc_sparse_upload", "csc_normalize", "csc_sparse_select_rows", "csc_hvg_finish", "csc_scale_fill",
              "csc_scale_gram_partial", "csc_pca_partial", "csc_pca_solve", "csc_pca_transform", "csc_knn", "csc_graph_build"):
        assert s in syms


def test_library_exports_every_declared_symbol(cs):
    lib = ctypes.CDLL(cs._capi.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_binding_table_matches_header(cs):
    assert sorted(cs._capi.SIGNATURES) == declared_symbols()
    lib = cs.load_library()
    assert lib.csc_abi_version() == 2


def test_header_cites_reference_lines():
    text = open(HEADER).read()
    for cite in ("processing.jl:15-31", "processing.jl:133-155", "processing.jl:65-93", "processing.jl:164-193",
                 "processing.jl:218-232"):
        assert cite in text


def test_no_gpu_fails_loudly(cs):
    """Without a B200 the product path raises; it never falls back to the CPU."""
    import pytest
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(cs.CellScopesCudaError):
        cs.Context(0)
    with pytest.raises(cs.CellScopesCudaError):
        import numpy as np
        cs.normalize_object(np.eye(3))
