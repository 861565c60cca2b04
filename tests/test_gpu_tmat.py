"""GPU parity: the total-variability path (odin_tmat_* through odin_b200.ml.Tmatrix) against the
reference's golden outputs (tests/golden/tmat.npz, produced by the real Tmatrix class) and against
the oracle (oracle/tmatrix.py) on fresh problems.

fp64 end to end, so the tolerance is set by conditioning, not by the arithmetic: 1e-8 on statistics,
1e-6 on the T-matrix after EM (row signs normalised: the orthogonalisation's SVD fixes rows only up to a
sign, see odin_b200/ml/tmat.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relmax
from oracle import tmatrix as OT

pytestmark = pytest.mark.gpu


class _FittedGMM(object):
  pass


def _gmm_stub(sigma):
  """A fitted odin_b200 GMM carrying the given variances (only feat_dim / nmix / sigma are read)."""
  from odin_b200.ml import GMM
  D, M = sigma.shape
  g = GMM(nmix=M, nmix_start=M)
  g.initialize(np.zeros((4, D), dtype=np.float32))
  g.sigma = sigma.copy()
  assert g.is_initialized and g.is_fitted
  return g


def _problem(seed, D, M, n_files, rank=3):
  from oracle.make_golden import _tmat_problem
  return _tmat_problem(seed=seed, D=D, M=M, n_files=n_files)


def test_golden_reference_run():
  from odin_b200.ml import Tmatrix
  g = np.load(os.path.join(GOLDEN, "tmat.npz"))
  sigma, Z, F, tv = g["sigma"], g["Z"], g["F"], int(g["tv_dim"])
  t = Tmatrix(tv, _gmm_stub(sigma), niter=3)
  assert np.array_equal(t.Tm, g["T0"])                      # host RNG contract (gmm_tmat.py:1469-1471)
  assert relmax(t.T_invS, g["T_invS0"]) < 1e-13 and relmax(t.T_invS_Tt, g["T_invS_Tt0"]) < 1e-13
  LU, RU, llk, nframes = t.expectation(Z, F)
  assert relmax(LU, g["LU0"]) < 1e-9 and relmax(RU, g["RU0"]) < 1e-9
  assert abs(llk - float(g["llk0"])) < 1e-9 * abs(float(g["llk0"])) and nframes == float(g["nframes0"])
  for it in range(3):
    t.expectation_maximization(Z, F)
    a, _ = OT.sign_normalise(t.Tm)
    b, _ = OT.sign_normalise(g["T%d" % (it + 1)])
    assert relmax(a, b) < 1e-6, (it, relmax(a, b))
  assert np.allclose(t._llk_hist, g["llk_hist"], rtol=1e-8)
  # i-vectors of the training files: coordinates follow the row signs of T
  _, sg_ours = OT.sign_normalise(t.Tm)
  _, sg_ref = OT.sign_normalise(g["T3"])
  iv = t.transform((Z, F))
  assert relmax(iv * sg_ours[None, :], g["ivec"] * sg_ref[None, :]) < 1e-6


@pytest.mark.parametrize("tv,D,M,n_files", [(16, 12, 20, 150), (64, 60, 32, 300), (128, 20, 8, 90), (33, 7, 5, 70),
                                            (176, 24, 10, 230), (400, 24, 20, 450)])
def test_em_iteration_vs_oracle(tv, D, M, n_files):
  """One full EM iteration and i-vector extraction on fresh statistics, including tv_dim 128 (the largest whose
  systems fit shared memory next to the M-step's right-hand sides), an odd one (a bye in the Jacobi tournament) and
  176 (every factorisation on the global-memory route) and 400 (the NIST-SRE recipe's tv_dim)."""
  from odin_b200.ml import Tmatrix
  sigma, Z, F = _problem(tv + D, D, M, n_files)
  t = Tmatrix(tv, _gmm_stub(sigma), niter=1)
  Sigma = OT.sigma_row(sigma)
  T0 = OT.init_T(tv, Sigma)
  T_invS, T_invS_Tt = OT.refresh(T0, Sigma, D)
  LU, RU, llk, nframes = t.expectation(Z, F)
  oLU, oRU, ollk, onframes = OT.expectation(Z, F, T_invS, T_invS_Tt)
  assert relmax(LU, oLU) < 1e-9 and relmax(RU, oRU) < 1e-9
  assert abs(llk - ollk) < 1e-9 * abs(ollk) and nframes == onframes
  # M-step pieces separately, then combined
  t.maximization(oLU, oRU, onframes, min_div_est=False, orthogonalize=False)
  assert relmax(t.Tm, OT.maximization(oLU, oRU, onframes, D, False, False)) < 1e-8
  t.maximization(oLU, oRU, onframes, min_div_est=True, orthogonalize=False)
  assert relmax(t.Tm, OT.maximization(oLU, oRU, onframes, D, True, False)) < 1e-8
  t.maximization(oLU, oRU, onframes, min_div_est=True, orthogonalize=True)
  T1 = OT.maximization(oLU, oRU, onframes, D, True, True)
  a, sa = OT.sign_normalise(t.Tm)
  b, sb = OT.sign_normalise(T1)
  assert relmax(a, b) < 1e-6, relmax(a, b)
  G = t.Tm.dot(t.Tm.T)                                        # rows orthogonal, singular values descending
  off = G - np.diag(np.diag(G))
  assert np.abs(off).max() < 1e-10 * np.abs(np.diag(G)).max()
  assert np.all(np.diff(np.diag(G)) <= 1e-12 * np.diag(G)[0])
  T_invS1, T_invS_Tt1 = OT.refresh(T1, Sigma, D)
  iv = t.transform((Z[:17], F[:17]))
  assert relmax(iv * sa[None, :], OT.ivector(Z[:17], F[:17], T_invS1, T_invS_Tt1) * sb[None, :]) < 1e-6


def test_fit_improves_likelihood_and_float32_stats():
  """fit() on float32 statistics (what GMM.transform_to_disk writes): the EM objective rises."""
  from odin_b200.ml import Tmatrix
  sigma, Z, F = _problem(5, 10, 16, 400)
  t = Tmatrix(8, _gmm_stub(sigma), niter=5)
  t.fit((Z.astype(np.float32), F.astype(np.float32)))
  h = np.array(t._llk_hist)
  assert len(h) == 5 and np.all(np.diff(h) > -1e-9 * np.abs(h[:-1]))
  assert t.is_fitted and t.transform((Z[:3].astype(np.float32), F[:3].astype(np.float32))).shape == (3, 8)


def test_limits_fail_loudly():
  from odin_b200 import _lib
  from odin_b200.ml import Tmatrix
  sigma, _, _ = _problem(1, 4, 3, 5)
  with pytest.raises(_lib.OdinError):
    Tmatrix(1025, _gmm_stub(sigma))


def test_ivector_wrapper(tmp_path):
  """odin.ml.Ivector surface (ivector.py:83-520): fit trains UBM + T-matrix and stores the models, transform returns
  the i-vectors of new files; the numbers equal driving GMM and Tmatrix by hand."""
  from odin_b200.ml import GMM, Ivector, Tmatrix
  rng = np.random.RandomState(3)
  D, n_utt = 8, 30
  cents = rng.randn(4, D) * 3
  lens = rng.randint(40, 90, size=n_utt)
  off = np.concatenate([[0], np.cumsum(lens)])
  X = np.concatenate([cents[rng.randint(0, 4, size=n)] + rng.randn(n, D) + 0.3 * rng.randn(1, D) for n in lens]).astype(np.float32)
  indices = [("utt%02d" % i, (int(off[i]), int(off[i + 1]))) for i in range(n_utt)]
  iv = Ivector(str(tmp_path / "ivec"), nmix=4, tv_dim=3, niter_gmm=3, niter_tmat=2)
  iv.fit(X, indices=indices, extract_ivecs=True, keep_stats=False)
  assert iv.is_fitted and os.path.exists(iv.gmm_path) and os.path.exists(iv.tmat_path)
  assert not os.path.exists(iv.z_path) and os.path.exists(iv.ivec_path)
  train_iv = np.load(iv.ivec_path)
  out = np.asarray(iv.transform(X, indices=indices, name="again", save_ivecs=True))
  assert out.shape == (n_utt, 3) and np.allclose(out, train_iv, rtol=1e-5, atol=1e-6)
  assert os.path.exists(iv.get_i_path("again"))
  # by hand: the same UBM schedule, statistics, T-matrix EM and extraction
  g = GMM(nmix=4, nmix_start=1, niter=3)
  g.fit((X, indices))
  pz, pf = str(tmp_path / "z.npy"), str(tmp_path / "f.npy")
  g.transform_to_disk(X, indices, pathZ=pz, pathF=pf)
  t = Tmatrix(3, g, niter=2)
  t.fit((np.load(pz), np.load(pf)))
  ref = t.transform((np.load(pz), np.load(pf)))
  assert np.allclose(out, ref.astype(np.float32), rtol=1e-4, atol=1e-5)
  # a reloaded wrapper finds its pickled models
  iv2 = Ivector(str(tmp_path / "ivec"), nmix=4, tv_dim=3)
  assert iv2.is_fitted
  assert np.allclose(np.asarray(iv2.transform(X, indices=indices)), out, rtol=1e-5, atol=1e-6)
