"""spectrograms_b200 -- a B200-native (sm_100a) engine for the hot path of the ``spectrograms`` crate
(jmg049/Spectrograms): ``stft()``, ``StftPlan``, ``SpectrogramPlanner`` plans (linear / mel / ERB / LogHz x power /
magnitude / dB), ``mfcc_from_log_mel`` and the fused ``mfcc()``, in f32 and f64, plus the first adjacent caller of that
path: ``chromagram()`` / ``chromagram_from_spectrogram`` (src/chroma.rs) and the interaural cue spectrograms of
src/binaural.rs.

The package is a thin host-side mirror of the reference's plan API over a C-ABI CUDA library
(``include/sgx_b200.h`` -> ``spectrograms_b200/lib/libsgx_b200.so``). There is no CPU compute path: if the library
or a GPU is missing, compute calls raise ``FFTBackendError``.
"""
from .errors import (DimensionMismatchError, FFTBackendError, InternalError, InvalidInputError, SpectrogramError)
from .params import (ChromaNorm, ChromaParams, ErbParams, GammatoneParams, LogHzParams, LogParams, MelNorm, MelParams, MfccParams,
                     SpectrogramParams, StftParams, WindowType)
from .plan import (ChromaPlan, Chromagram, FftPlanner, Mfcc, MfccPlan, Spectrogram, SpectrogramPlan, SpectrogramPlanner, StftPlan, StftResult,
                   compute_erb_db_spectrogram, compute_erb_magnitude_spectrogram, compute_erb_power_spectrogram,
                   compute_linear_db_spectrogram, compute_linear_magnitude_spectrogram,
                   compute_linear_power_spectrogram, compute_loghz_db_spectrogram,
                   compute_loghz_magnitude_spectrogram, compute_loghz_power_spectrogram, compute_mel_db_spectrogram,
                   compute_mel_magnitude_spectrogram, compute_mel_power_spectrogram, compute_mfcc, compute_stft, fft, irfft, istft, build_chroma_filterbank, chromagram,
                   chromagram_from_spectrogram, compute_chromagram,
                   magnitude_spectrum, mfcc, mfcc_from_log_mel, power_spectrum, rfft, stft)
from .binaural import (BinauralSpectrogram, ILDSpectrogramParams, ILRSpectrogramParams, IPDSpectrogramParams,
                       ITDSpectrogramParams, binaural_from_stft, compute_ild_spectrogram, compute_ilr_spectrogram,
                       compute_ipd_spectrogram, compute_itd_spectrogram)
from .sharding import shard_range
from . import serde

__version__ = "0.1.0"
