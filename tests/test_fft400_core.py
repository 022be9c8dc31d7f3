"""CPU check of the n_fft=400 kernel's register-level core (fft400_core.cuh) by host emulation: the very task
functions the CUDA kernel runs are compiled for the host and driven sequentially over one tile, then compared with
the oracle's power spectrum. Catches index-map / twiddle / butterfly mistakes without a GPU."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from conftest import ROOT, make_signal, rel_l2

EMU_SRC = os.path.join(ROOT, "tests", "emu", "fft400_emu.cu")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libfft400_emu.so")


@pytest.fixture(scope="module")
def emu():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    dep = os.path.join(ROOT, "spectrograms_b200", "csrc", "fft400_core.cuh")
    if (not os.path.exists(EMU_LIB)) or os.path.getmtime(EMU_LIB) < max(os.path.getmtime(EMU_SRC), os.path.getmtime(dep)):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
                        "-cudart", "static", "-o", EMU_LIB, EMU_SRC], check=True, capture_output=True)
    lib = C.CDLL(EMU_LIB)
    lib.emu_fft400_tile.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("sig", ["noise", "sine", "chirp", "impulse"])
@pytest.mark.parametrize("f0", [0, 32, 64])
def test_tile_power_matches_oracle(emu, sig, f0):
    n = 16000 if f0 < 64 else 11000          # f0 = 64 with n = 11000: the tile runs past the end of the clip
    x = make_signal(sig, n, 16000.0, np.float32)
    plan = oracle.Plan(oracle.Desc(dtype="f64", n_fft=400, hop=160, sample_rate=16000.0))
    win = oracle.Plan(oracle.Desc(dtype="f32", n_fft=400, hop=160)).window()
    ref = plan.compute(x.astype(np.float64))                        # (201, n_frames)
    out = np.zeros((201, 32), dtype=np.float32)
    emu.emu_fft400_tile(x.ctypes.data, x.size, f0, win.ctypes.data, out.ctypes.data)
    nf = min(32, ref.shape[1] - f0)
    assert nf > 0
    got = out[:, :nf].astype(np.float64)
    want = ref[:, f0:f0 + nf]
    assert rel_l2(got, want) < 1e-6
    # per-frame accuracy, not just aggregate: every frame's spectrum within f32 FFT accuracy of its own energy
    for j in range(nf):
        assert rel_l2(got[:, j], want[:, j]) < 2e-6
