"""Inverse-path widening (reference src/spectrogram.rs:4789-4911, src/fft_backend.rs:509-567):
irfft and istft. CPU: the oracle against SciPy's pocketfft and against round trips; GPU: the C ABI against the oracle and
the size-independent property istft(stft(x)) == x away from the edges."""
import numpy as np
import pytest
import scipy.fft

import oracle
import spectrograms_b200 as sg
from conftest import make_signal, rel_l2

SR = 16000.0
SIZES = [8, 15, 400, 512, 1000, 1009]


# ------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("n", SIZES)
def test_oracle_irfft_is_the_true_inverse(n):
    x = make_signal("noise", n, SR, seed=n)
    X = oracle.rfft(x, n)
    assert np.abs(oracle.irfft(X, n) - x).max() < 1e-13                           # tests/fft_padding_tests.rs round trip
    assert np.abs(oracle.irfft(X, n) - scipy.fft.irfft(X, n)).max() < 1e-13       # independent implementation
    X32 = X.astype(np.complex64)
    assert np.abs(oracle.irfft(X32, n) - x).max() < 5e-6 and oracle.irfft(X32, n).dtype == np.float32
    with pytest.raises(oracle.OracleError):
        oracle.irfft(X[:-1], n)                                                   # DimensionMismatch (:4797-4802)


@pytest.mark.parametrize("n_fft,hop,centre", [(512, 128, True), (400, 160, True), (512, 256, False), (1000, 250, True), (256, 256, True)])
def test_oracle_istft_round_trip_and_length(n_fft, hop, centre):
    x = make_signal("chirp", 9000, SR)
    p = oracle.Plan(oracle.Desc(dtype="f64", n_fft=n_fft, hop=hop, centre=centre))
    S = p.stft(x)
    y = oracle.istft(S, n_fft, hop, "hanning", centre)
    full = (S.shape[1] - 1) * hop + n_fft
    assert len(y) == (full - 2 * (n_fft // 2) if centre else full)                # :4836-4841, :4893-4902
    m = min(len(y), len(x))
    lo, hi = (n_fft, m - n_fft)
    if hop < n_fft:        # without overlap the Hann zeros at the frame edges cannot be divided back (norm <= 1e-10 is skipped)
        assert np.abs(y[lo:hi] - x[lo:hi]).max() < 1e-12                          # window-energy normalisation makes OLA exact


def test_oracle_istft_ignores_dc_and_nyquist_imaginary_parts():
    x = make_signal("noise", 4000, SR)
    S = oracle.Plan(oracle.Desc(dtype="f64", n_fft=256, hop=64)).stft(x)
    S2 = S.copy()
    S2[0] += 0.5j
    S2[-1] -= 0.25j
    assert np.array_equal(oracle.istft(S, 256, 64), oracle.istft(S2, 256, 64))    # realfft zeroes them (and reports an error)


# ------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_irfft_matches_oracle(n, dtype):
    x = make_signal("noise", n, SR, dtype=dtype, seed=3 * n)
    X = oracle.rfft(x, n)
    got = sg.irfft(X, n)
    assert got.dtype == dtype and got.shape == (n,)
    assert rel_l2(got, oracle.irfft(X, n)) < (1e-5 if dtype == np.float32 else 1e-12)
    assert rel_l2(got, x) < (1e-5 if dtype == np.float32 else 1e-12)
    with pytest.raises(sg.DimensionMismatchError) as e:
        sg.irfft(X[:-1], n)
    assert (e.value.expected, e.value.got) == (n // 2 + 1, n // 2)


@pytest.mark.gpu
@pytest.mark.parametrize("n_fft,hop,centre,window", [(512, 128, True, "hanning"), (400, 160, True, "hanning"), (2048, 512, True, "hamming"),
                                                     (512, 256, False, "hanning"), (1000, 250, True, "blackman"), (1009, 300, True, "hanning"),
                                                     (256, 256, True, "hanning")])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_istft_matches_oracle_and_round_trips(n_fft, hop, centre, window, dtype):
    dt = np.float32 if dtype == "float32" else np.float64
    tol = 1e-5 if dtype == "float32" else 1e-12
    x = np.stack([make_signal(k, 12000 + 5, SR, dtype=dt) for k in ("sine", "chirp", "noise")])
    sp = sg.SpectrogramParams(sg.StftParams(n_fft, hop, window, centre), SR)
    plan = sg.StftPlan(sp, dtype)
    S = plan.compute_batch(x)                                                       # (3, bins, frames) complex
    y = plan.istft(S)
    full = (S.shape[2] - 1) * hop + n_fft
    assert y.dtype == dt and y.shape == (3, full - 2 * (n_fft // 2) if centre else full)
    for i in range(3):
        want = oracle.istft(S[i], n_fft, hop, window, centre)
        # where the accumulated squared window is tiny but above 1e-10 (first / last samples without centring, frame
        # seams without overlap) the division amplifies the last-ulp differences of two f32 inverse FFTs by up to 1 / w:
        # the strict budget applies to the interior, the edges get the amplified one
        assert rel_l2(y[i, n_fft:-n_fft], want[n_fft:-n_fft]) < (tol if hop < n_fft else 50 * tol), i
        assert rel_l2(y[i], want) < 50 * tol, i
        m = min(y.shape[1], x.shape[1])
        if hop < n_fft:
            assert rel_l2(y[i, n_fft:m - n_fft], x[i, n_fft:m - n_fft]) < 10 * tol   # property: istft(stft(x)) == x
    one = sg.istft(S[1], n_fft, hop, window, centre)                                 # the reference's free function
    assert np.array_equal(one, y[1])


@pytest.mark.gpu
def test_istft_device_tensors_and_errors():
    import torch
    sp = sg.SpectrogramParams(sg.StftParams(512, 128, "hanning", True), SR)
    plan = sg.StftPlan(sp, "float32")
    x = torch.randn(4, 20000, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    S = plan.compute_batch(x)
    y = plan.istft(S)
    assert y.is_cuda and y.shape == (4, (S.shape[2] - 1) * 128 + 512 - 512)
    assert float((y[:, 512:19000] - x[:, 512:19000]).abs().max()) < 2e-5
    assert np.array_equal(plan.istft(S.cpu().numpy()), y.cpu().numpy())              # host path = device path
    with pytest.raises(sg.DimensionMismatchError):
        plan.istft(S[:, :200])
    with pytest.raises(sg.InvalidInputError, match="hop_size must be <= n_fft"):
        sg.istft(S[0].cpu().numpy(), 512, 600)
