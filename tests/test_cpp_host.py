"""The C++ host layer (include/sgx_b200.hpp) over the C ABI: compiled with g++ against the in-tree library and run.
Without a GPU the program checks the host-only half of the reference's behavioural checklist and that compute fails
loudly; on the B200 box (-m gpu) it also runs the known-answer / property half."""
import os
import subprocess

import pytest

from conftest import ROOT

LIB_DIR = os.path.join(ROOT, "spectrograms_b200", "lib")
SRC = os.path.join(ROOT, "tests", "cpp", "test_cpp_host.cpp")


def _build(tmp_path):
    exe = str(tmp_path / "test_cpp_host")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIB_DIR, "-lsgx_b200", f"-Wl,-rpath,{LIB_DIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_layer_builds_and_passes_host_checks(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "CPP_HOST_OK" in r.stdout


@pytest.mark.gpu
def test_cpp_host_layer_compute_checks(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "CPP_HOST_OK host+gpu" in r.stdout


def test_fast_divisors_match_integer_division(tmp_path):
    """fastdiv.hpp make_fastdiv / fd_div (the generic family's index math) against exact division, on the host."""
    exe = str(tmp_path / "test_fastdiv")
    src = os.path.join(ROOT, "tests", "cpp", "test_fastdiv.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-O2", src, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "FASTDIV_OK" in r.stdout, r.stdout + r.stderr
