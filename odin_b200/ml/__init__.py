"""Mirror of ``odin.ml`` for the accelerated path (odin/ml/__init__.py:16-21 exports GMM,
Tmatrix, Ivector, PLDA, Scorer; the GMM-UBM half and the T-matrix / i-vector extractor are the CUDA
path, the PLDA / cosine scoring back-end is small host linear algebra, see plda.py / scoring.py)."""
from .gmm import GMM  # noqa: F401
from .tmat import Tmatrix  # noqa: F401
from .ivector import Ivector  # noqa: F401
from .plda import PLDA  # noqa: F401
from .scoring import Scorer, VectorNormalizer  # noqa: F401
from . import scoring, plda  # noqa: F401
