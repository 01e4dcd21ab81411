"""CPU-only: libsnarkv_cuda.so loads and exports every symbol include/snarkv_cuda.h declares; no compute without a GPU."""
import ctypes
import os
import re

import pytest

import snark_verifier_b200 as sv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "snarkv_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(snarkv_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_table_agree():
    assert declared_symbols() == sorted(sv.C_ABI), "snark_verifier_b200.C_ABI must list exactly the header's functions"


def test_library_exports_every_declared_symbol():
    assert os.path.exists(sv.LIB_PATH), "libsnarkv_cuda.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(sv.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    sv.load_library()
    assert b"sm_100a" in sv.load_library().snarkv_version()


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sv.CudaError):
        sv.CudaLoader(0)


def test_product_sources_never_reference_the_oracle():
    """The oracle is test infrastructure: nothing under snark_verifier_b200/ or include/ may include/import/link it."""
    bad = []
    for base in ("snark_verifier_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if not fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp", "Makefile")):
                    continue
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"(import\s+oracle|from\s+oracle|oracle/|liboracle|bn254_model)", txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
