import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker (oracle) and make sure the product library exists. Building the checker is not using it."""
    import oracle
    oracle.build()
    from spectrograms_b200 import build as b
    if not os.path.exists(b.LIB_PATH):
        b.build()
    yield


def make_signal(kind: str, n: int, sr: float, dtype=np.float64, seed: int = 0, freq: float = 440.0) -> np.ndarray:
    """Synthetic inputs of SURVEY.md section 8(d): sine (tests/spectrogram_tests.rs:10-16), chirp (notebook cell 1),
    noise (python/tests/test_dtype_planner.py:18-19)."""
    t = np.arange(n) / sr
    if kind == "sine":
        x = np.sin(2.0 * np.pi * freq * np.arange(n) / sr)
    elif kind == "chirp":
        x = np.sin(2.0 * np.pi * (100.0 + 3000.0 * t * t) * t)
    elif kind == "noise":
        x = np.random.default_rng(seed).standard_normal(n)
    elif kind == "silence":
        x = np.zeros(n)
    elif kind == "impulse":
        x = np.zeros(n)
        x[n // 2] = 1.0
    else:
        raise ValueError(kind)
    return x.astype(dtype)


def rel_l2(a, b) -> float:
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.astype(np.complex128 if np.iscomplexobj(b) else np.float64))
    num = np.linalg.norm(a.astype(np.complex128 if np.iscomplexobj(a) else np.float64) - b)
    return float(num / den) if den > 0 else float(num)
