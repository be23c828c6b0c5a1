"""The C++ host mirror (include/cosdata_b200.hpp) compiles against the C ABI, links libcosdata_b200.so and behaves:
without a device it fails loudly with CDB_CUDA_ERROR; with one (run under -m gpu too) it computes through the kernels."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(tmp_path, name="abi_smoke"):
    from cosdata_b200 import _lib
    _lib.load()
    exe = str(tmp_path / name)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.join(ROOT, "cosdata_b200")
    subprocess.run([cxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe, "-L", libdir, "-lcosdata_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return subprocess.run([exe], capture_output=True, text=True)


def test_cpp_mirror_builds_links_and_fails_loudly_or_computes(tmp_path):
    r = _build_and_run(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "devices=" in r.stdout


import pytest


@pytest.mark.gpu
def test_cpp_mirror_computes_on_gpu(tmp_path):
    r = _build_and_run(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cosine=" in r.stdout and "top=1 count=2" in r.stdout


def test_cpp_mirror_round2_surface_builds_and_fails_loudly_without_a_device(tmp_path):
    r = _build_and_run(tmp_path, "mirror_round2")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "devices=" in r.stdout

