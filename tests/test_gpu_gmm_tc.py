"""GPU parity of the tcgen05 / 3xTF32 Baum-Welch kernels (odin_b200/csrc/gmm_tc.cu,
impl=2) against the fp64 oracle and against the fp32 CUDA-core kernels (impl=1).
Tolerance (north_star): <= 1e-3 on N/F/S as max|a-b| / max|b|; the kernels are
expected to sit near fp32 round-off, which TIGHT pins."""
import os

import numpy as np
import pytest

from conftest import relmax
from odin_b200 import synth
from oracle import gmm as OG

pytestmark = pytest.mark.gpu

TOL_STATS = 1e-3
TIGHT = 5e-5


def _gmm(M, mean, sigma, w, impl):
  from odin_b200.ml import GMM
  g = GMM(nmix=M, nmix_start=M, impl=impl)
  g.initialize(np.zeros((1, mean.shape[0]), dtype=np.float32))
  g.mean, g.sigma, g.w = mean.copy(), sigma.copy(), w.copy()
  return g


def _check(X, mean, sigma, w, sad=None, tol=TIGHT):
  M = mean.shape[1]
  z, f, s, l, n = OG.expectation(X, mean, sigma, w, sad=sad, compute_dtype=np.float64)
  Z, F, S, L = _gmm(M, mean, sigma, w, 2).expectation(X, sad=sad)
  errs = (relmax(Z, z), relmax(F, f), relmax(S, s))
  assert max(errs) < tol, errs
  assert abs(float(L) - float(l)) < 1e-4 * max(1.0, abs(float(l))), (L, l)
  assert abs(Z.sum() - n) < 1e-4 * max(n, 1)
  return Z, F, S, L


@pytest.mark.parametrize("D,M,N", [(60, 128, 32), (60, 128, 5000), (60, 256, 4099), (60, 2048, 3000),
                                   (60, 200, 1000), (40, 384, 2049), (24, 300, 64), (4, 96, 500)])
def test_tc_estep_vs_oracle(D, M, N):
  X = synth.gmm_features(N, D, 8, seed=D + M)
  mean, sigma, w = synth.gmm_params(D, M, seed=M)
  _check(X, mean, sigma, w)


def test_tc_matches_fp32_kernels_and_mask():
  D, M, N = 60, 512, 20000
  X = synth.gmm_features(N, D, 32, seed=7)
  mean, sigma, w = synth.gmm_params(D, M, seed=8)
  rng = np.random.RandomState(3)
  sad = (rng.rand(N) > 0.4).astype(np.uint8)
  Z2, F2, S2, L2 = _check(X, mean, sigma, w, sad=sad)
  Z1, F1, S1, L1 = _gmm(M, mean, sigma, w, 1).expectation(X, sad=sad)
  assert relmax(Z2, Z1) < TIGHT and relmax(F2, F1) < TIGHT and relmax(S2, S1) < TIGHT
  assert abs(float(L1) - float(L2)) < 1e-4 * abs(float(L1))
  # nothing selected -> exact zeros
  Z, F, S, L = _gmm(M, mean, sigma, w, 2).expectation(X, sad=np.zeros(N, dtype=np.uint8))
  assert np.all(Z == 0) and np.all(F == 0) and np.all(S == 0) and float(L) == 0.0


def test_tc_flush_interval_and_sub_batches(monkeypatch):
  """the fp32 TMEM accumulator is drained into fp64 every ODIN_TC_FLUSH_TILES tiles and
  pass 1 runs in ODIN_TC_SUB_BATCH-frame launches: neither may change the result."""
  D, M, N = 60, 256, 150000
  X = synth.gmm_features(N, D, 32, seed=17)
  mean, sigma, w = synth.gmm_params(D, M, seed=18)
  ref = None
  for flush, sub in (("512", str(1 << 20)), ("7", "40000"), ("100000", "33")):
    monkeypatch.setenv("ODIN_TC_FLUSH_TILES", flush)
    monkeypatch.setenv("ODIN_TC_SUB_BATCH", sub)
    out = _check(X, mean, sigma, w)
    if ref is None:
      ref = out
    else:
      assert relmax(out[0], ref[0]) < 1e-5 and relmax(out[1], ref[1]) < 1e-5 and relmax(out[2], ref[2]) < 1e-5


def test_tc_em_iterations_2048():
  """config-4 protocol at reduced size: 3 EM iterations of a 2048-mix UBM."""
  D, M, N = 60, 2048, 60000
  X = synth.gmm_features(N, D, 64, seed=27)
  rng = np.random.RandomState(5)
  mean = X[rng.choice(N, M, replace=False)].T.copy()
  sigma = np.tile(X.var(0)[:, None], (1, M)).astype(np.float32)
  w = np.full((1, M), 1.0 / M, dtype=np.float32)
  gm = _gmm(M, mean, sigma, w, 2)
  om, os_, ow = mean.astype(np.float64), sigma.astype(np.float64), w.astype(np.float64)
  for it in range(3):
    gm.expectation_maximization(X, print_progress=False)
    z, f, s, l, _ = OG.expectation(X, om, os_, ow, compute_dtype=np.float64)
    om, os_, ow, rb = OG.maximization(z, f, s, (om, os_, ow))
    assert not rb
  assert relmax(gm.mean, om) < TOL_STATS and relmax(gm.sigma, os_) < TOL_STATS and relmax(gm.w, ow) < TOL_STATS
  assert abs(gm._llk_hist[M][-1] - l) < 1e-3 * abs(l)


def test_tc_linearity_large():
  """size-independent property on a 3 M-frame shard at 2048 mixtures: stats(whole) =
  stats(first half) + stats(second half); sum(Z) = #frames."""
  import torch
  N, D, M = 3_000_000, 60, 2048
  g = torch.Generator(device="cuda")
  g.manual_seed(1)
  X = torch.randn(N, D, generator=g, device="cuda") * 2.0
  mean, sigma, w = synth.gmm_params(D, M, seed=22)
  gm = _gmm(M, mean, sigma * 4.0, w, 2)
  Z, F, S, L = gm.expectation(X)
  Za, Fa, Sa, La = gm.expectation(X[:N // 2])
  Zb, Fb, Sb, Lb = gm.expectation(X[N // 2:])
  assert relmax(Za + Zb, Z) < 1e-5 and relmax(Fa + Fb, F) < 1e-5 and relmax(Sa + Sb, S) < 1e-5
  assert abs(Z.sum() - N) < 1e-4 * N
