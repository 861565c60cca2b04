"""Mirror of ``odin.ml`` for the accelerated path (odin/ml/__init__.py:16-21 exports GMM,
Tmatrix, Ivector, PLDA, Scorer; only the GMM-UBM half is on this path)."""
from .gmm import GMM  # noqa: F401
