"""The C-ABI library loads and exports every symbol include/knpemi_b200.h declares.
No compute calls: this box has no GPU, and the library must say so loudly."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "knpemi_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kem_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("kem_create", "kem_destroy", "kem_set_column", "kem_get_column", "kem_set_uniform",
                 "kem_step", "kem_step_timed", "kem_step_io", "kem_sync", "kem_last_error",
                 "kem_model_load", "kem_set_stimulus_mask", "kem_fp64_peak"):
        assert must in syms


def test_library_exports_every_declared_symbol(built):
    from knpemi_b200 import _cabi
    lib = ctypes.CDLL(_cabi.library_path())
    for sym in declared_symbols():
        assert hasattr(lib, sym), f"{sym} declared in include/knpemi_b200.h but not exported"


def test_ctypes_signature_table_covers_the_header(built):
    from knpemi_b200 import _cabi
    assert sorted(_cabi.SIGNATURES) == declared_symbols()
    assert _cabi.lib().kem_version() >= 100


def test_header_cites_the_reference_for_each_entry_group():
    text = open(HEADER).read()
    assert text.count("odeSolver.py:") >= 12


def test_structs_match_the_header_layout(built):
    from knpemi_b200 import _cabi
    assert ctypes.sizeof(_cabi.kem_step_times) == 4 * 8
    assert ctypes.sizeof(_cabi.kem_io_column) == 16
    assert ctypes.sizeof(_cabi.kem_model_info) == 5 * 4 + 64 * 4 + 64 + 32


def test_generated_model_libraries_export_one_descriptor(built):
    from knpemi_b200 import codegen
    from knpemi_b200.models import BUILTIN
    for name, mod in BUILTIN.items():
        path, em = codegen.model_library(mod)
        assert os.path.exists(path)
        lib = ctypes.CDLL(path)
        assert hasattr(lib, "kem_model_descriptor")
        # loading into the runtime only reads the descriptor: no device needed
        from knpemi_b200 import _cabi
        mid = _cabi.load_model(path)
        info = _cabi.model_info(mid)
        assert (info.ns, info.np) == (em.ns, em.np)
        assert info.name.decode() == name
        assert [info.out_cols[k] for k in range(info.n_out)] == em.out_cols


def _no_gpu():
    from knpemi_b200 import _cabi
    return _cabi.device_count() == 0


def test_no_cpu_fallback_without_a_device(built):
    """On a box without a GPU constructing a MembraneModel must fail loudly."""
    if not _no_gpu():
        pytest.skip("a CUDA device is present")
    import numpy as np
    from knpemi_b200._cabi import KemError
    from knpemi_b200.ducks import PointSpace
    from knpemi_b200.models import hh_test
    from knpemi_b200.odeSolver import MembraneModel
    with pytest.raises(KemError, match="no CUDA device|no CPU fallback"):
        MembraneModel(hh_test, None, 1, PointSpace(np.zeros((4, 3))), verbose=False)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may touch it."""
    pkg = os.path.join(ROOT, "knp-emi-fenics-x_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) and "_generated" not in dirpath:
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libknpemi_oracle" not in text, f
                assert "kemo_" not in text, f


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: the header must compile as C (no C++-isms, no CUDA types)."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "knpemi_b200.h"\nint main(void) { kem_handle h = 0; (void)h; return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I",
                        os.path.join(ROOT, "include"), "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(HEADER).read()
    assert "cudaStream_t" not in text and "#include <cuda" not in text
    assert "torch" not in text.lower()
