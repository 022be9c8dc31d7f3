"""Dense (ERB) filterbanks outside the n400_tc kernel: the plan's FFT family writes the linear power spectrogram into plan scratch
and the filterbank runs as row blocks of the tcgen05 GEMM kernel (dense_rows_tc = the dct2_lifter_tc kernel, 3xTF32, amplitude
scaling fused; src/erb.rs:374-402). Checked against the f64 oracle on every family that feeds it, in all three amplitude modes,
against the CUDA-core epilogue of the same plan (sgx_plan_set_tensor_cores(plan, 0)), on batches, partial tiles and silence."""
import numpy as np
import pytest

import oracle
import spectrograms_b200 as sg
from conftest import make_signal, rel_l2

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


CASES = [  # n_fft, hop, sample rate, bands, expected family
    (512, 160, 16000.0, 128, "r2c_fused_pow2"),       # two row blocks of 64
    (1024, 256, 22050.0, 64, "r2c_fused_pow2"),       # 513 columns: 48 + 16 rows
    (2048, 512, 22050.0, 40, "r2c_fused_pow2"),       # 1025 columns: 16-row blocks
    (800, 200, 16000.0, 40, "r2c_fused_mixed"),
    (1009, 250, 16000.0, 24, "r2c_fused_generic"),
    (400, 160, 16000.0, 128, "r2c_fused_n400"),       # more bands than the fused n400_tc kernel holds
]


@pytest.mark.parametrize("n_fft,hop,sr,bands,family", CASES)
@pytest.mark.parametrize("amp", ["power", "magnitude", "db"])
def test_dense_rows_tc_against_oracle(n_fft, hop, sr, bands, family, amp):
    torch = _torch()
    n = 7 * n_fft + 123
    x = np.stack([make_signal(k, n, sr, np.float32, seed=n_fft + i) for i, k in enumerate(("noise", "chirp", "sine"))])
    params = sg.SpectrogramParams(sg.StftParams(n_fft, hop, sg.WindowType.hanning(), True), sr)
    db = sg.LogParams(-80.0) if amp == "db" else None
    plan = sg.SpectrogramPlanner().erb_plan(params, sg.ErbParams(bands, 50.0, sr / 2), db, amp, "float32")
    assert plan.kernel_name() == family + "+dense_rows_tc"
    got = plan.compute_batch(torch.from_numpy(x).cuda()).cpu().numpy()
    assert plan.last_launch_count() >= 2
    cc = sg.SpectrogramPlanner().erb_plan(params, sg.ErbParams(bands, 50.0, sr / 2), db, amp, "float32")
    cc.set_tensor_cores(False)
    assert cc.kernel_name() == family
    ref_cc = cc.compute_batch(torch.from_numpy(x).cuda()).cpu().numpy()
    o = oracle.Plan(oracle.Desc(dtype="f64", n_fft=n_fft, hop=hop, sample_rate=sr, mapping="erb", n_bands=bands, f_min=50.0, f_max=sr / 2,
                                amp=amp, floor_db=-80.0 if amp == "db" else None))
    for i in range(x.shape[0]):
        ref = o.compute(x[i].astype(np.float64))
        assert got[i].shape == ref.shape
        if amp == "db":
            # tones leave most bands at the frame's f32 rounding floor (any f32 implementation, the reference's own included): the
            # 1e-3 dB bar applies within 60.2 dB of each frame's peak, as in tests/test_gpu_parity.py::db_check; noise everywhere
            d = np.abs(got[i] - ref)
            mask = ref >= (ref.max(axis=0, keepdims=True) - 60.2) if i else np.ones_like(ref, dtype=bool)
            assert d[mask].max() <= 1e-3, (i, d[mask].max(), d.max())
            assert np.abs(got[i] - ref_cc[i])[mask].max() <= 1e-3
        else:
            assert rel_l2(got[i], ref) <= 1e-5, (i, rel_l2(got[i], ref))
            assert rel_l2(got[i], ref_cc[i].astype(np.float64)) <= 1e-5
    assert (got >= (-80.0 if amp == "db" else 0.0)).all()


def test_dense_rows_tc_silence_single_frame_and_f64_stays_on_cuda_cores():
    torch = _torch()
    params = sg.SpectrogramParams(sg.StftParams(512, 160, sg.WindowType.hanning(), True), 16000.0)
    plan = sg.SpectrogramPlanner().erb_plan(params, sg.ErbParams(40, 50.0, 8000.0), sg.LogParams(-80.0), "db", "float32")
    z = plan.compute(torch.zeros(4000, dtype=torch.float32, device="cuda")).data.cpu().numpy()
    assert (z == -80.0).all()
    x = make_signal("noise", 3000, 16000.0, np.float32, seed=5)
    full = plan.compute(torch.from_numpy(x).cuda()).data.cpu().numpy()
    one = plan.compute_frame(torch.from_numpy(x).cuda(), 7)
    one = one.cpu().numpy() if hasattr(one, "cpu") else np.asarray(one)
    assert np.abs(one - full[:, 7]).max() <= 1e-4
    p64 = sg.SpectrogramPlanner().erb_plan(params, sg.ErbParams(40, 50.0, 8000.0), None, "power", "float64")
    assert p64.kernel_name() == "r2c_fused_pow2"
