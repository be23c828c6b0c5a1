"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/cosdata_b200.h declares, and fails loudly (no CPU fallback) when
there is no CUDA device."""
import os
import re

import numpy as np
import pytest

import cosdata_b200 as cdb
from cosdata_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "cosdata_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cdb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cosdata_b200.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert lib.cdb_abi_version() == 1


def test_code_bytes_and_host_synth_match_oracle():
    import oracle as orc
    for st in range(6):
        for dim in (1, 7, 8, 9, 128, 768, 1024):
            assert cdb.code_bytes(st, dim) == orc.code_bytes(st, dim)
    assert np.array_equal(cdb.synth_matrix(5, 10, 33, first_row=3), orc.synth_matrix(5, 10, 33, first_row=3))


def test_no_cpu_fallback_without_device():
    if cdb.device_count() > 0:
        pytest.skip("CUDA device present")
    q = cdb.ScalarQuantization()
    with pytest.raises(cdb.CosdataError) as e:
        q.quantize(np.zeros(8, np.float32), cdb.StorageType.UnsignedByte)
    assert e.value.status == cdb.Status.CUDA_ERROR
    with pytest.raises(cdb.CosdataError):
        cdb.DenseIndex(dim=8, capacity=4)


def test_invalid_params_are_rejected_before_touching_the_device():
    lib = _lib.load()
    assert lib.cdb_quantize_batch(0, 9, -1.0, 1.0, None, 0, 8, None, None) == cdb.Status.INVALID_PARAMS
    assert lib.cdb_index_create(None, None) == cdb.Status.INVALID_PARAMS
    assert b"" != lib.cdb_last_error_string()


def test_header_is_plain_c11(tmp_path):
    """the boundary is a C ABI: the header must compile as C (no C++-isms, no CUDA/torch types)"""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "cosdata_b200.h"\nint main(void) { cdb_index_desc d; cdb_graph_metadata m; (void)d; (void)m; return 0; }\n')
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_integration_extern_block_covers_the_header():
    """INTEGRATION.md's `extern "C"` block (what the Rust shim binds) + the listed helper functions == every function of the header"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "cosdata_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    declared = set(re.findall(r"\b(cdb_[a-z0-9_]+)\s*\(", hdr))
    bound = set(re.findall(r"pub fn (cdb_[a-z0-9_]+)", doc))
    helpers = {"cdb_synth_fill_host", "cdb_index_append_synthetic", "cdb_index_read_codes", "cdb_index_stats",
               "cdb_index_last_candidate_counts", "cdb_index_last_kernel_ms", "cdb_index_scan_ms_history", "cdb_kernel_launch_count",
               "cdb_index_hnsw_profile", "cdb_debug_set_hnsw_flags", "cdb_debug_tensor_peak"}
    assert bound <= declared, sorted(bound - declared)
    assert declared - bound == helpers, (sorted(declared - bound - helpers), sorted(helpers - (declared - bound)))
    for h in helpers:
        assert h in doc, h
