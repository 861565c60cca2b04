"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's total-variability
model (``odin/ml/gmm_tmat.py`` class ``Tmatrix``, lines 1343-2090): T-matrix training on
Baum-Welch statistics and i-vector extraction.  Pinned to the reference by
``tests/test_oracle_vs_reference.py`` (reference executed under ``oracle/ref_shim.py``) and by
``tests/golden/tmat.npz`` (``oracle/make_golden.py tmat``).

Layout (as the reference): ``Z [n_files, M]`` zeroth-order statistics, ``F [n_files, M*D]``
centred first-order statistics with column ``m*D + d`` (gmm_tmat.py:754-757), ``Tm [tv, M*D]``,
``T_invS_Tt [M, tv(tv+1)/2]`` in ``np.tril_indices`` order.
"""
import numpy as np
from scipy import linalg

EPS = 1e-6  # gmm_tmat.py:27


def sigma_row(gmm_sigma, dtype=np.float64):
  """gmm_tmat.py:1466-1468: GMM variances [D, M] -> [1, M*D], mixture-major."""
  D, M = gmm_sigma.shape
  return np.array(gmm_sigma.reshape((1, D * M), order="F"), dtype=dtype)


def init_T(tv_dim, Sigma, seed=1234, dtype=np.float64):
  """gmm_tmat.py:1469-1471."""
  np.random.seed(seed)
  return (np.random.randn(tv_dim, Sigma.shape[1]) * Sigma.sum() * 0.001).astype(dtype)


def refresh(Tm, Sigma, feat_dim):
  """gmm_tmat.py:1578-1589 -> (T_invS [tv, MD], T_invS_Tt [M, t2])."""
  tv = Tm.shape[0]
  nmix = Tm.shape[1] // feat_dim
  itril = np.tril_indices(tv)
  T_invS = Tm / (Sigma + EPS)
  T_invS2 = Tm / (np.sqrt(Sigma) + EPS)
  T_invS_Tt = np.empty((nmix, tv * (tv + 1) // 2), dtype=Tm.dtype)
  for mix in range(nmix):
    blk = T_invS2[:, feat_dim * mix:feat_dim * (mix + 1)]
    T_invS_Tt[mix] = blk.dot(blk.T)[itril]
  return T_invS, T_invS_Tt


def expectation(Z, F, T_invS, T_invS_Tt):
  """gmm_tmat.py:1694-1725 (numpy branch) -> LU [M, t2], RU [tv, MD], llk, nframes."""
  tv = T_invS.shape[0]
  dtype = T_invS.dtype
  itril = np.tril_indices(tv)
  Im = np.eye(tv, dtype=dtype)
  nframes = np.ceil(Z.sum())
  nfiles = F.shape[0]
  L1 = np.dot(Z, T_invS_Tt)
  B1 = np.dot(F, T_invS.T)
  Ex = np.empty((nfiles, tv), dtype=dtype)
  Exx = np.empty((nfiles, tv * (tv + 1) // 2), dtype=dtype)
  llk = np.empty((nfiles, 1), dtype=dtype)
  for ix in range(nfiles):
    L = np.zeros((tv, tv), dtype=dtype)
    L[itril] = L1[ix]
    L = L + np.tril(L, k=-1).T + Im
    Cxx = linalg.inv(L)
    B = B1[ix][:, np.newaxis]
    this_Ex = np.dot(Cxx, B)
    Ex[ix] = this_Ex.T
    llk[ix] = -0.5 * this_Ex.T.dot(B - this_Ex) + this_Ex.T.dot(B)
    Exx[ix] = (Cxx + this_Ex.dot(this_Ex.T))[itril]
  RU = np.dot(Ex.T, F)
  LU = np.dot(Z.T, Exx)
  return LU, RU, llk.sum(), nframes


def maximization(LU, RU, nframes, feat_dim, min_div_est=True, orthogonalize=True):
  """gmm_tmat.py:1818-1865 (numpy branch) -> new Tm [tv, MD]."""
  tv = RU.shape[0]
  nmix = LU.shape[0]
  itril = np.tril_indices(tv)
  Tm = np.empty_like(RU)
  for mix in range(nmix):
    lu = np.zeros((tv, tv), dtype=RU.dtype)
    lu[itril] = LU[mix, :]
    lu += np.tril(lu, -1).T
    Tm[:, feat_dim * mix:feat_dim * (mix + 1)] = linalg.solve(lu, RU[:, feat_dim * mix:feat_dim * (mix + 1)])
  if min_div_est:
    lu = np.zeros((tv, tv))
    lu[itril] = LU.sum(0) / nframes
    lu += np.tril(lu, -1).T
    Tm = np.dot(linalg.cholesky(lu), Tm)      # scipy default: UPPER factor, lu = U^T U
  if orthogonalize:
    _, s_, V_ = linalg.svd(Tm, full_matrices=False)
    Tm = np.diag(s_).dot(V_)
  return Tm.astype(RU.dtype)


def ivector(Z, F, T_invS, T_invS_Tt):
  """gmm_tmat.py:1898-1942 for one utterance (Z [1, M], F [1, MD]) or a batch -> [n, tv]."""
  tv = T_invS.shape[0]
  itril = np.tril_indices(tv)
  out = np.empty((Z.shape[0], tv), dtype=T_invS.dtype)
  for i in range(Z.shape[0]):
    L = np.zeros((tv, tv), dtype=T_invS.dtype)
    L[itril] = np.dot(Z[i:i + 1], T_invS_Tt)
    L += np.tril(L, -1).T + np.eye(tv, dtype=T_invS.dtype)
    out[i] = np.dot(linalg.inv(L), np.dot(T_invS, F[i:i + 1].T)).T
  return out


def fit(Z, F, tv_dim, gmm_sigma, niter, seed=1234, dtype=np.float64):
  """gmm_tmat.py:2044-2090 given (Z, F): niter x (expectation, maximization).  Returns
  (Tm, T_invS, T_invS_Tt, [llk / nfiles per iteration])."""
  D = gmm_sigma.shape[0]
  Sigma = sigma_row(gmm_sigma, dtype)
  Z = np.asarray(Z)
  F = np.asarray(F)
  Tm = init_T(tv_dim, Sigma, seed, dtype)
  T_invS, T_invS_Tt = refresh(Tm, Sigma, D)
  hist = []
  for _ in range(niter):
    LU, RU, llk, nframes = expectation(Z, F, T_invS, T_invS_Tt)
    Tm = maximization(LU, RU, nframes, D)
    T_invS, T_invS_Tt = refresh(Tm, Sigma, D)
    hist.append(llk / Z.shape[0])
  return Tm, T_invS, T_invS_Tt, hist


def sign_normalise(A):
  """Rows of a T-matrix / columns of i-vectors are defined up to a sign (the SVD of
  gmm_tmat.py:1857-1859): flip every row so that its largest-magnitude entry is positive."""
  A = np.array(A, copy=True)
  j = np.argmax(np.abs(A), axis=1)
  sgn = np.sign(A[np.arange(A.shape[0]), j])
  sgn[sgn == 0] = 1.0
  return A * sgn[:, None], sgn
