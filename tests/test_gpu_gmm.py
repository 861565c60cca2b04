"""GPU parity: GMM-UBM kernels (through the C-ABI, via odin_b200.ml.GMM) against
the oracle (oracle/gmm.py) and the golden vectors produced by the real
reference.  Tolerance (BASELINE.json north_star): <= 1e-3 on N/F/S and on the
UBM parameters after EM, measured as max|a-b| / max|b| per matrix."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relmax
from odin_b200 import synth
from oracle import gmm as OG

pytestmark = pytest.mark.gpu

TOL_STATS = 1e-3   # north_star
TIGHT = 2e-5       # what fp32 / 3xTF32 kernels actually reach on these sizes


def _gmm(M, mean, sigma, w, impl=0, **kw):
  from odin_b200.ml import GMM
  g = GMM(nmix=M, nmix_start=M, impl=impl, **kw)
  g.initialize(np.zeros((1, mean.shape[0]), dtype=np.float32))
  g.mean, g.sigma, g.w = mean.copy(), sigma.copy(), w.copy()
  return g


def _impls(D, M):
  import torch
  out = [1]
  # the tcgen05 path is taken when the library reports support for the shape
  try:
    g = _gmm(M, *synth.gmm_params(D, M), impl=2)
    g.expectation(synth.gmm_features(256, D, 4, seed=1))
    out.append(2)
  except Exception:
    pass
  return out


@pytest.mark.parametrize("tag", ["np2", "f32"])
def test_appendix_b_estep_mstep(tag):
  g = np.load(os.path.join(GOLDEN, "gmm_appendix_b.npz"))
  gm = _gmm(8, g["mean"], g["sigma"], g["w"], impl=1)
  Z, F, S, L = gm.expectation(g["X"])
  assert Z.shape == (1, 8) and F.shape == (6, 8) and S.shape == (6, 8)
  assert relmax(Z, g[tag + "_Z"]) < TIGHT and relmax(F, g[tag + "_F"]) < TIGHT
  assert relmax(S, g[tag + "_S"]) < TIGHT and abs(float(L) - float(g[tag + "_L"])) < 1e-5
  assert abs(Z.sum() - 1000.0) < 1e-2
  Zt, Ft = gm.transform(g["X"][:100])
  assert Ft.shape == (1, 48)
  assert relmax(Zt, g[tag + "_Zt"]) < TIGHT and relmax(Ft, g[tag + "_Ft"]) < 1e-4
  gm.maximization(Z, F, S)
  assert relmax(gm.mean, g[tag + "_mean1"]) < TIGHT and relmax(gm.sigma, g[tag + "_sigma1"]) < 1e-4
  assert relmax(gm.w, g[tag + "_w1"]) < TIGHT


@pytest.mark.parametrize("impl", [1, 2])
def test_d60_m64_golden(impl):
  g = np.load(os.path.join(GOLDEN, "gmm_d60_m64.npz"))
  if impl not in _impls(60, 64):
    pytest.skip("tcgen05 path not available for this shape")
  gm = _gmm(64, g["mean"], g["sigma"], g["w"], impl=impl)
  Z, F, S, L = gm.expectation(g["X"])
  for tag in ("np2", "f32"):
    assert relmax(Z, g[tag + "_Z"]) < TOL_STATS and relmax(F, g[tag + "_F"]) < TOL_STATS
    assert relmax(S, g[tag + "_S"]) < TOL_STATS and abs(float(L) - float(g[tag + "_L"])) < 1e-3
  assert relmax(Z, g["np2_Z"]) < TIGHT and relmax(F, g["np2_F"]) < TIGHT and relmax(S, g["np2_S"]) < TIGHT
  # with a SAD mask (gmm_tmat.py:162-164)
  Z, F, S, L = gm.expectation(g["X"], sad=g["sad"])
  assert relmax(Z, g["np2_Zsad"]) < TIGHT and relmax(F, g["np2_Fsad"]) < TIGHT
  assert relmax(S, g["np2_Ssad"]) < TIGHT and abs(float(L) - float(g["np2_Lsad"])) < 1e-4
  assert abs(Z.sum() - float(g["sad"].sum())) < 1e-2
  # per-utterance centred statistics
  idx = {"u%d" % i: (int(s), int(e)) for i, (s, e) in enumerate(g["utt_bounds"])}
  names = gm.transform_to_disk(g["X"], idx)
  Zu, Fu = gm.last_utt_stats_
  assert names == ["u0", "u1", "u2"] and Fu.shape == (3, 60 * 64)
  assert relmax(Zu, g["utt_Z"]) < TIGHT and relmax(Fu, g["utt_Fhat"]) < 1e-4


@pytest.mark.parametrize("D,M,N", [(6, 8, 777), (20, 16, 5000), (39, 100, 3001), (60, 512, 4096),
                                   (60, 2048, 2500), (24, 300, 64), (60, 64, 1)])
def test_estep_shapes_vs_oracle(D, M, N):
  X = synth.gmm_features(N, D, 8, seed=D + M)
  mean, sigma, w = synth.gmm_params(D, M, seed=M)
  z, f, s, l, n = OG.expectation(X, mean, sigma, w, compute_dtype=np.float64)
  for impl in _impls(D, M):
    gm = _gmm(M, mean, sigma, w, impl=impl)
    Z, F, S, L = gm.expectation(X)
    assert relmax(Z, z) < TOL_STATS and relmax(F, f) < TOL_STATS and relmax(S, s) < TOL_STATS, impl
    assert abs(float(L) - float(l)) < 1e-3 * max(1.0, abs(float(l)))
    assert abs(Z.sum() - N) < 1e-3 * N


def test_empty_and_all_masked():
  X = synth.gmm_features(300, 12, 4, seed=5)
  mean, sigma, w = synth.gmm_params(12, 8, seed=6)
  gm = _gmm(8, mean, sigma, w, impl=1)
  Z, F, S, L = gm.expectation(X, sad=np.zeros(300, dtype=np.uint8))
  assert np.all(Z == 0) and np.all(F == 0) and np.all(S == 0) and float(L) == 0.0
  sad = np.zeros(300, dtype=np.uint8)
  sad[17] = 1
  Z, F, S, L = gm.expectation(X, sad=sad)
  z, f, s, l, n = OG.expectation(X, mean, sigma, w, sad=sad, compute_dtype=np.float64)
  assert n == 1 and relmax(Z, z) < TIGHT and relmax(F, f) < TIGHT


def test_linearity_of_statistics_full_size():
  """size-independent property at config-2 scale (360 000 x 60, M=64): stats of
  the whole = sum of stats of two halves; sum(Z) = #frames."""
  import torch
  N, D, M = 360000, 60, 64
  X = torch.from_numpy(synth.gmm_features(N, D, 32, seed=21)).cuda()
  mean, sigma, w = synth.gmm_params(D, M, seed=22)
  gm = _gmm(M, mean, sigma, w)
  Z, F, S, L = gm.expectation(X)
  Za, Fa, Sa, La = gm.expectation(X[:N // 2])
  Zb, Fb, Sb, Lb = gm.expectation(X[N // 2:])
  # (each call picks its own power-of-two operand scales on the tensor-core path, which auto-dispatch now uses at every
  # mixture count: regrouping is exact to its split precision, ~2e-6, not to fp32 rounding)
  assert relmax(Za + Zb, Z) < 1e-5 and relmax(Fa + Fb, F) < 1e-5 and relmax(Sa + Sb, S) < 1e-5
  assert abs(0.5 * (float(La) + float(Lb)) - float(L)) < 1e-6 * abs(float(L))
  assert abs(Z.sum() - N) < 1e-4 * N


def test_scores_and_posteriors():
  X = synth.gmm_features(500, 20, 8, seed=31)
  mean, sigma, w = synth.gmm_params(20, 16, seed=32)
  gm = _gmm(16, mean, sigma, w)
  prec, mup, Cc = OG.posterior_constants(mean.astype(np.float64), sigma.astype(np.float64), w.astype(np.float64))
  X64 = X.astype(np.float64)
  lp = -0.5 * (Cc + X64**2 @ prec - 2 * X64 @ mup + 20 * np.log(2 * np.pi))
  llk = OG._lse(lp)
  assert np.max(np.abs(gm.logprob(X) - lp)) < 2e-3
  assert np.max(np.abs(gm.llk(X) - llk)) < 1e-3 and gm.score(X).shape == (500, 1)
  post = gm.postprob(X)
  assert np.max(np.abs(post - np.exp(lp - llk))) < 1e-4 and np.allclose(post.sum(1), 1.0, atol=1e-4)


def test_fit_matches_reference_schedule():
  """gmm_tmat.py:625-699: 1 -> 8 mixtures with the split schedule; parameters
  after EM within 1e-3 of the REAL reference (golden)."""
  from odin_b200.ml import GMM
  g = np.load(os.path.join(GOLDEN, "gmm_fit_d12_m8.npz"))
  gm = GMM(nmix=8, nmix_start=1, niter=4)
  gm.fit(g["X"])
  assert gm.is_fitted and gm.mean.shape == (12, 8)
  assert [len(gm._llk_hist[k]) for k in (1, 2, 4, 8)] == [1, 2, 4, 4]
  assert relmax(gm.mean, g["mean"]) < TOL_STATS and relmax(gm.sigma, g["sigma"]) < TOL_STATS
  assert relmax(gm.w, g["w"]) < TOL_STATS
  np.testing.assert_allclose([gm._llk_hist[k][-1] for k in (1, 2, 4, 8)], g["llk_last"], rtol=1e-4)
  # resume: a fitted model does nothing more (gmm_tmat.py:683-687)
  before = gm.mean.copy()
  gm.fit(g["X"])
  assert np.array_equal(before, gm.mean)


def test_ten_em_iterations_vs_oracle():
  """UBM parameters after 10 EM iterations (config-4 protocol, reduced size) <= 1e-3."""
  D, M, N = 60, 64, 20000
  X = synth.gmm_features(N, D, 32, seed=41)
  mean, sigma, w = synth.gmm_params(D, M, seed=42)
  sigma = sigma * 4.0
  gm = _gmm(M, mean, sigma, w)
  om, os_, ow = mean.astype(np.float64), sigma.astype(np.float64), w.astype(np.float64)
  for it in range(10):
    gm.expectation_maximization(X, print_progress=False)
    z, f, s, l, _ = OG.expectation(X, om, os_, ow, compute_dtype=np.float64)
    om, os_, ow, rb = OG.maximization(z, f, s, (om, os_, ow))
    assert not rb
  assert relmax(gm.mean, om) < TOL_STATS and relmax(gm.sigma, os_) < TOL_STATS and relmax(gm.w, ow) < TOL_STATS
  # oracle L is already the per-frame mean; fp32 parameters vs the fp64 oracle drift by ~1e-5 relative over 10 iterations
  assert abs(gm._llk_hist[M][-1] - l) < 1e-4 * abs(l)


def test_mixup_and_rollback():
  mean, sigma, w = synth.gmm_params(10, 4, seed=51)
  gm = _gmm(4, mean, sigma, w)
  gm._nmix = 8
  gm._handle = None
  m2, s2, w2 = OG.mixup(mean, sigma, w, 8)
  gm.gmm_mixup()
  assert gm.mean.shape == (10, 8) and relmax(gm.mean, m2) < 1e-6 and relmax(gm.sigma, s2) < 1e-7
  assert relmax(gm.w, w2) < 1e-7
  # truncated split (nmix not a power of two), gmm_tmat.py:1324-1331
  gm = _gmm(4, mean, sigma, w)
  gm._nmix = 6
  gm._handle = None
  gm.gmm_mixup()
  m3, s3, w3 = OG.mixup(mean, sigma, w, 6)
  assert gm.mean.shape == (10, 6) and relmax(gm.mean, m3) < 1e-6 and relmax(gm.w, w3) < 1e-7
  # rollback when a variance goes negative (gmm_tmat.py:1259-1266)
  gm = _gmm(4, mean, sigma, w)
  Z = np.ones((1, 4)); F = np.ones((10, 4)) * 3.0; S = np.ones((10, 4))  # S/Z - (F/Z)^2 < 0
  gm.maximization(Z, F, S)
  assert relmax(gm.mean, mean) < 1e-7 and relmax(gm.sigma, sigma) < 1e-7
  gm.allow_rollback = False
  gm.maximization(Z, F, S)
  assert np.all(gm.sigma == 0) and relmax(gm.mean, F / (Z + 1e-6)) < 1e-6


def test_host_array_streaming_equals_resident():
  """the chunked pinned-memory H2D path gives the same statistics as a resident tensor."""
  import torch
  from odin_b200.ml import gmm as G
  X = synth.gmm_features(50000, 60, 16, seed=61)
  mean, sigma, w = synth.gmm_params(60, 64, seed=62)
  gm = _gmm(64, mean, sigma, w)
  Zr, Fr, Sr, Lr = gm.expectation(torch.from_numpy(X).cuda())
  frames = G._DeviceFrames(X, chunk_frames=7000)
  Zh, Fh, Sh, Lh = gm.expectation(frames)
  # chunk boundaries regroup the fp32 partial sums that are flushed into the fp64 statistics
  # (and, on the tensor-core path, every chunk gets its own operand scales: equal to its split precision, ~2e-6)
  assert relmax(Zh, Zr) < 1e-5 and relmax(Fh, Fr) < 1e-5 and relmax(Sh, Sr) < 1e-5
  X16 = X.astype(np.float16)  # float16 stores are up-cast on load
  Z16, F16, S16, L16 = gm.expectation(X16)
  z, f, s, l, _ = OG.expectation(X16.astype(np.float32), mean, sigma, w, compute_dtype=np.float64)
  assert relmax(Z16, z) < TOL_STATS and relmax(S16, s) < TOL_STATS


def test_maximization_floor_const_prevents_rollback():
  """gmm_tmat.py:1254-1268: the variance floor is applied BEFORE the negative-variance test, so a positive
  `floor_const` keeps the update where the unfloored one would have been rolled back."""
  from odin_b200.ml import GMM
  D, M = 8, 4
  rng = np.random.RandomState(2)
  Z = rng.rand(1, M) * 50 + 10
  F = rng.randn(D, M) * Z
  S = (F / Z) ** 2 * Z + rng.rand(D, M) * Z            # variances in (0, 1)
  S[3, 1] = (F[3, 1] / Z[0, 1]) ** 2 * Z[0, 1] - 5.0   # one negative variance
  prev = (rng.randn(D, M).astype(np.float32), np.ones((D, M), np.float32), np.full((1, M), 0.25, np.float32))

  def run(floor):
    g = GMM(nmix=M, nmix_start=M, impl=1)
    g.initialize(np.zeros((1, D), np.float32))
    g.mean, g.sigma, g.w = [a.copy() for a in prev]
    g.maximization(Z, F, S, floor_const=floor)
    return g

  g0 = run(None)
  assert np.array_equal(g0.mean, prev[0]) and np.array_equal(g0.sigma, prev[1])     # rolled back (allow_rollback)
  g1 = run(0.1)
  iN = 1.0 / (Z + 1e-6)
  mu, sig = F * iN, S * iN - (F * iN) ** 2
  wgt = Z / Z.sum()
  ref = sig.clip(sig.dot(wgt.T) * 0.1)
  assert relmax(g1.mean, mu) < 1e-5 and relmax(g1.sigma, ref) < 1e-5 and relmax(g1.w, wgt) < 1e-5
