"""Error taxonomy: the four ``SpectrogramError`` variants (reference src/error.rs:13-28) under the exception names
the reference's Python bindings use (src/python/error.rs:10-66)."""


class SpectrogramError(Exception):
    """Base class (reference: ``SpectrogramError``)."""


class InvalidInputError(SpectrogramError):
    """``SpectrogramError::InvalidInput(String)``."""


class DimensionMismatchError(SpectrogramError):
    """``SpectrogramError::DimensionMismatch { expected, got }``."""

    def __init__(self, message: str, expected: int = 0, got: int = 0):
        super().__init__(message)
        self.expected = expected
        self.got = got


class FFTBackendError(SpectrogramError):
    """``SpectrogramError::FftBackendError { backend: "cuda", msg }`` -- CUDA errors and a missing GPU land here."""


class InternalError(SpectrogramError):
    """``SpectrogramError::InternalError(String)``."""
