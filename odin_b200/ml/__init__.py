"""Mirror of ``odin.ml`` for the accelerated path (odin/ml/__init__.py:16-21 exports GMM,
Tmatrix, Ivector, PLDA, Scorer; the GMM-UBM half and the T-matrix / i-vector extractor are on this
path)."""
from .gmm import GMM  # noqa: F401
from .tmat import Tmatrix  # noqa: F401
from .ivector import Ivector  # noqa: F401
