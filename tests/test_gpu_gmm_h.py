"""GPU parity of the 3xFP16 tcgen05 Baum-Welch kernels (odin_b200/csrc/gmm_h.cu, impl=3)
against the fp64 oracle, the fp32 CUDA-core kernels (impl=1) and the 3xTF32 kernels
(impl=2).  Tolerance (north_star): <= 1e-3 on N/F/S as max|a-b| / max|b|; the split
keeps ~22 significant bits, which TIGHT pins."""
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
  Z, F, S, L = _gmm(M, mean, sigma, w, 3).expectation(X, sad=sad)
  errs = (relmax(Z, z), relmax(F, f), relmax(S, s))
  assert max(errs) < tol, errs
  assert abs(float(L) - float(l)) < 1e-4 * max(1.0, abs(float(l))), (L, l)
  assert abs(Z.sum() - n) < 1e-4 * max(n, 1)
  return Z, F, S, L


@pytest.mark.parametrize("D,M,N", [(60, 256, 32), (60, 256, 5000), (60, 512, 4099), (60, 2048, 3000),
                                   (60, 300, 1000), (40, 384, 2049), (24, 300, 64), (4, 257, 500)])
def test_h_estep_vs_oracle(D, M, N):
  X = synth.gmm_features(N, D, 8, seed=D + M)
  mean, sigma, w = synth.gmm_params(D, M, seed=M)
  _check(X, mean, sigma, w)


def test_h_wide_dynamic_range():
  """fp16 halves only hold 5 exponent bits: features of very different scale per
  dimension (1e-3 .. 1e3) and tight / loose variances must survive the exact
  power-of-two column and row scaling."""
  D, M, N = 60, 256, 6000
  rng = np.random.RandomState(11)
  scale = (10.0 ** rng.uniform(-3, 3, size=D)).astype(np.float32)
  X = (synth.gmm_features(N, D, 8, seed=5) * scale[None, :]).astype(np.float32)
  mean, sigma, w = synth.gmm_params(D, M, seed=6)
  mean = (mean * scale[:, None]).astype(np.float32)
  sigma = (sigma * (scale[:, None] ** 2) * (10.0 ** rng.uniform(-1, 1, size=(D, M)))).astype(np.float32)
  _check(X, mean, sigma, w)


def test_h_matches_other_kernels_and_mask():
  D, M, N = 60, 512, 20000
  X = synth.gmm_features(N, D, 32, seed=7)
  mean, sigma, w = synth.gmm_params(D, M, seed=8)
  rng = np.random.RandomState(3)
  sad = (rng.rand(N) > 0.4).astype(np.uint8)
  Z3, F3, S3, L3 = _check(X, mean, sigma, w, sad=sad)
  for impl in (1, 2):
    Z1, F1, S1, L1 = _gmm(M, mean, sigma, w, impl).expectation(X, sad=sad)
    assert relmax(Z3, Z1) < TIGHT and relmax(F3, F1) < TIGHT and relmax(S3, S1) < TIGHT
    assert abs(float(L1) - float(L3)) < 1e-4 * abs(float(L1))
  # nothing selected -> exact zeros
  Z, F, S, L = _gmm(M, mean, sigma, w, 3).expectation(X, sad=np.zeros(N, dtype=np.uint8))
  assert np.all(Z == 0) and np.all(F == 0) and np.all(S == 0) and float(L) == 0.0


def test_h_flush_interval_and_sub_batches(monkeypatch):
  """the fp32 TMEM accumulator is drained into fp64 every ODIN_H_FLUSH_TILES tiles and the
  operand images are rebuilt per ODIN_H_SUB_BATCH frames: neither may change the result."""
  D, M, N = 60, 256, 150000
  X = synth.gmm_features(N, D, 32, seed=17)
  mean, sigma, w = synth.gmm_params(D, M, seed=18)
  ref = None
  for flush, sub in (("256", str(1 << 20)), ("7", "40000"), ("100000", "128")):
    monkeypatch.setenv("ODIN_H_FLUSH_TILES", flush)
    monkeypatch.setenv("ODIN_H_SUB_BATCH", sub)
    out = _check(X, mean, sigma, w)
    if ref is None:
      ref = out
    else:
      assert relmax(out[0], ref[0]) < 1e-5 and relmax(out[1], ref[1]) < 1e-5 and relmax(out[2], ref[2]) < 1e-5


def test_h_em_iterations_2048():
  """config-4 protocol at reduced size: 3 EM iterations of a 2048-mix UBM."""
  D, M, N = 60, 2048, 60000
  X = synth.gmm_features(N, D, 64, seed=27)
  rng = np.random.RandomState(5)
  mean = X[rng.choice(N, M, replace=False)].T.copy()
  sigma = np.tile(X.var(0)[:, None], (1, M)).astype(np.float32)
  w = np.full((1, M), 1.0 / M, dtype=np.float32)
  gm = _gmm(M, mean, sigma, w, 3)
  om, os_, ow = mean.astype(np.float64), sigma.astype(np.float64), w.astype(np.float64)
  for it in range(3):
    gm.expectation_maximization(X, print_progress=False)
    z, f, s, l, _ = OG.expectation(X, om, os_, ow, compute_dtype=np.float64)
    om, os_, ow, rb = OG.maximization(z, f, s, (om, os_, ow))
    assert not rb
  assert relmax(gm.mean, om) < TOL_STATS and relmax(gm.sigma, os_) < TOL_STATS and relmax(gm.w, ow) < TOL_STATS
  assert abs(gm._llk_hist[M][-1] - l) < 1e-3 * abs(l)


def test_h_linearity_large():
  """size-independent property on a 3 M-frame shard at 2048 mixtures: stats(whole) =
  stats(first half) + stats(second half); sum(Z) = #frames."""
  import torch
  N, D, M = 3_000_000, 60, 2048
  g = torch.Generator(device="cuda")
  g.manual_seed(1)
  X = torch.randn(N, D, generator=g, device="cuda") * 2.0
  mean, sigma, w = synth.gmm_params(D, M, seed=22)
  gm = _gmm(M, mean, sigma * 4.0, w, 3)
  Z, F, S, L = gm.expectation(X)
  Za, Fa, Sa, La = gm.expectation(X[:N // 2])
  Zb, Fb, Sb, Lb = gm.expectation(X[N // 2:])
  assert relmax(Za + Zb, Z) < 1e-5 and relmax(Fa + Fb, F) < 1e-5 and relmax(Sa + Sb, S) < 1e-5
  assert abs(Z.sum() - N) < 1e-4 * N


def test_h_prepared_frames_equal_per_call_images(monkeypatch):
  """odin_gmm_estep_frames (operand images of resident frames built once, reused across EM
  iterations) must give the statistics of odin_gmm_estep(impl 3) (to fp64 round-off), with and without a
  SAD mask, and after the model changed (the images depend on the data only)."""
  import torch
  D, M, N = 60, 512, 70001
  X = torch.from_numpy(synth.gmm_features(N, D, 32, seed=41)).cuda()
  mean, sigma, w = synth.gmm_params(D, M, seed=42)
  rng = np.random.RandomState(4)
  sad = (rng.rand(N) > 0.3).astype(np.uint8)
  from odin_b200.ml.gmm import _DeviceFrames
  frames = _DeviceFrames(X)
  g = _gmm(M, mean, sigma, w, 3)
  for mask in (None, sad):
    a = g.expectation(frames, sad=mask)
    assert frames._prepared is not None
    monkeypatch.setenv("ODIN_H_NO_PREPARED", "1")
    b = g.expectation(X, sad=mask)
    monkeypatch.delenv("ODIN_H_NO_PREPARED")
    for u, v in zip(a, b):   # same operands and tiles; only the order of the fp64 atomic adds may differ
      assert relmax(np.asarray(u), np.asarray(v)) < 1e-12
  g.mean, g.sigma, g.w = mean * 1.01, sigma * 1.1, w
  Z, F, S, L = g.expectation(frames)
  z, f, s, l, n = OG.expectation(X.cpu().numpy(), g.mean, g.sigma, g.w, compute_dtype=np.float64)
  assert max(relmax(Z, z), relmax(F, f), relmax(S, s)) < TIGHT


def test_h_per_utterance_statistics():
  """GMM.transform_to_disk on long utterances at a tensor-core-sized model: per-utterance Z / F-hat through the
  tcgen05 E-step (one accumulator per utterance, odin_gmm_utt_stats -> gmm_utt_stats_h) vs the oracle, with a SAD
  mask, an empty utterance and lengths that are not multiples of the 128-frame operand tile."""
  import tempfile
  from odin_b200.ml import GMM
  from oracle import gmm as OG
  rng = np.random.RandomState(21)
  D, M = 60, 256
  lens = [1500, 3001, 0, 1277, 2048, 4100, 1030]
  off = np.concatenate([[0], np.cumsum(lens)])
  cents = rng.randn(M, D).astype(np.float32) * 2
  X = (cents[rng.randint(0, M, size=off[-1])] + rng.randn(off[-1], D)).astype(np.float32)
  mean = cents.T.copy()
  sigma = (0.6 + rng.rand(D, M)).astype(np.float32)
  w = (rng.rand(1, M) + 0.5).astype(np.float32)
  w /= w.sum()
  g = GMM(nmix=M, nmix_start=M)
  g.initialize(X)
  g.mean, g.sigma, g.w = mean, sigma, w
  sad = (rng.rand(off[-1]) > 0.3).astype(np.uint8)
  indices = [("u%d" % i, (int(off[i]), int(off[i + 1]))) for i in range(len(lens)) if lens[i] > 0]
  for mask in (None, sad):
    with tempfile.TemporaryDirectory() as d:
      pz, pf = os.path.join(d, "z.npy"), os.path.join(d, "f.npy")
      names = g.transform_to_disk(X, indices, sad=mask, pathZ=pz, pathF=pf)
      Zu, Fu = np.load(pz), np.load(pf)
    nr, zr, fr = OG.utterance_stats(X, [(n, dict(indices)[n]) for n in names], mean, sigma, w, sad=mask,
                                    compute_dtype=np.float64)
    assert nr == names and Zu.shape == zr.shape and Fu.shape == fr.shape
    assert relmax(Zu, zr) < 5e-5 and relmax(Fu, fr) < 5e-5, (relmax(Zu, zr), relmax(Fu, fr))


@pytest.mark.parametrize("sub_batch", [None, "1024"])
def test_h_per_utterance_statistics_short_utterances_segmented(sub_batch, monkeypatch):
  """Thousands of SHORT utterances (the digits of config 5) through the tcgen05 kernels in segmented mode
  (odin_gmm_utt_stats -> gmm_utt_stats_hseg: utterances padded to whole 64-frame tiles, the accumulator drained at
  every utterance change): lengths around the tile size (1, 63, 64, 65, 128, 129 frames), an empty utterance, a SAD mask,
  a long utterance in the middle, M not a multiple of the 128-mixture chunk; against the oracle and the fp32 route.
  With ODIN_H_SUB_BATCH=1024 the batch is cut into many groups of whole utterances (the 700-frame one alone in its)."""
  import tempfile
  if sub_batch is not None:
    monkeypatch.setenv("ODIN_H_SUB_BATCH", sub_batch)
  from odin_b200.ml import GMM
  from oracle import gmm as OG
  rng = np.random.RandomState(5)
  D, M = 60, 320
  lens = [1, 63, 64, 65, 128, 129, 0, 700, 2] + list(rng.randint(30, 200, size=400))
  off = np.concatenate([[0], np.cumsum(lens)])
  cents = rng.randn(M, D).astype(np.float32) * 2
  X = (cents[rng.randint(0, M, size=off[-1])] + rng.randn(off[-1], D)).astype(np.float32)
  mean = cents.T.copy()
  sigma = (0.6 + rng.rand(D, M)).astype(np.float32)
  w = (rng.rand(1, M) + 0.5).astype(np.float32)
  w /= w.sum()
  sad = (rng.rand(off[-1]) > 0.3).astype(np.uint8)
  indices = [("u%d" % i, (int(off[i]), int(off[i + 1]))) for i in range(len(lens)) if lens[i] > 0]
  res = {}
  for impl in (0, 1):   # 0: auto -> segmented tensor-core route (>= 16 384 frames); 1: fp32 CUDA-core kernels
    g = GMM(nmix=M, nmix_start=M)
    g.initialize(X)
    g.mean, g.sigma, g.w = mean, sigma, w
    g.impl = impl
    for tag, mask in (("all", None), ("sad", sad)):
      names = g.transform_to_disk(X, indices, sad=mask)
      res[impl, tag] = (names,) + tuple(np.array(a) for a in g.last_utt_stats_)
  for tag, mask in (("all", None), ("sad", sad)):
    names, Zu, Fu = res[0, tag]
    nr, zr, fr = OG.utterance_stats(X, [(n, dict(indices)[n]) for n in names], mean, sigma, w, sad=mask,
                                    compute_dtype=np.float64)
    assert nr == names and Zu.shape == zr.shape and Fu.shape == fr.shape
    assert relmax(Zu, zr) < 5e-5 and relmax(Fu, fr) < 5e-5, (tag, relmax(Zu, zr), relmax(Fu, fr))
    # row by row (a short utterance's row is tiny next to the matrix maximum)
    rowmax = np.abs(zr).max(1)
    assert np.all(np.abs(Zu - zr).max(1) <= 1e-4 * np.maximum(rowmax, 1e-3)), tag
    assert res[1, tag][0] == names
    assert relmax(res[1, tag][1], Zu) < 5e-5 and relmax(res[1, tag][2], Fu) < 5e-5


def test_h_segmented_statistics_add_up_at_config5_scale():
  """size-independent property of the segmented route at config 5's size (3 000 utterances of 60-200 frames, 512
  mixtures) and with utterances of config 3's lengths at 2048 mixtures: the per-utterance statistics add up to the
  statistics of the whole call (Z, and F after un-centring F-hat + Z mean), every row sums to its utterance's length.
  Tolerance 1e-4: the whole-call kernel drains its fp32 TMEM accumulator every 256 tiles (16 384 frames), whose
  truncating accumulation leaves its totals 2-4e-5 low (sum Z = N (1 - 2e-5) here); the per-utterance rows, drained every
  <= 32 tiles, agree with the fp32 CUDA-core kernels to 3e-6."""
  import torch
  for M, lo, hi, n_utt in ((512, 60, 200, 3000), (2048, 500, 6000, 300)):
    D = 60
    rng = np.random.RandomState(M)
    lens = rng.randint(lo, hi, size=n_utt)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(M)
    X = torch.randn(int(off[-1]), D, generator=gen, device="cuda") * 2.0
    mean, sigma, w = synth.gmm_params(D, M, seed=M + 3)
    gm = _gmm(M, mean, sigma * 4.0, w, 0)
    Z, F, S, L = gm.expectation(X)
    Zu, Fu = gm._utt_stats_device(X, None, off)
    Zu, Fu = Zu.double().cpu().numpy(), Fu.double().cpu().numpy().reshape(n_utt, M, D)
    assert np.abs(Zu.sum(1) - lens).max() < 1e-4 * lens.max()
    assert relmax(Zu.sum(0)[None, :], np.asarray(Z, dtype=np.float64).reshape(1, M)) < 1e-4
    Ftot = (Fu + Zu[:, :, None] * mean.T.astype(np.float64)[None]).sum(0)       # [M, D]
    assert relmax(Ftot.T, np.asarray(F, dtype=np.float64)) < 1e-4


@pytest.mark.parametrize("D,M", [(39, 256), (13, 64), (57, 512)])
def test_h_feature_dims_not_multiple_of_four(D, M):
  """D = 39 (13 MFCC + deltas) and friends: the frames are carried with zero columns up to the next multiple of four
  (GMM._kdim) so they run on the tcgen05 kernels; statistics, log-likelihood, scores, per-utterance statistics, the
  M-step and the mix-up must be those of the D-dimensional model."""
  from odin_b200 import _lib
  from odin_b200.ml import GMM
  N = 6000
  X = synth.gmm_features(N, D, 8, seed=D)
  mean, sigma, w = synth.gmm_params(D, M, seed=M + 1)
  g = _gmm(M, mean, sigma, w, 0)
  assert g._kdim == (D + 3) // 4 * 4 and g._kdim != D
  sad = (np.random.RandomState(1).rand(N) > 0.25).astype(np.uint8)
  Z, F, S, L = g.expectation(X, sad=sad)
  z, f, s, l, n = OG.expectation(X, mean, sigma, w, sad=sad, compute_dtype=np.float64)
  assert F.shape == (D, M) and S.shape == (D, M)
  assert max(relmax(Z, z), relmax(F, f), relmax(S, s)) < TIGHT
  assert abs(float(L) - float(l)) < 1e-4 * abs(float(l))
  im = _lib.C.c_int32()
  a, b = _lib.C.c_float(), _lib.C.c_float()
  _lib.check(_lib.load().odin_gmm_last_estep_ms(g._handle, _lib.C.byref(a), _lib.C.byref(b), _lib.C.byref(im)))
  assert im.value == 3                                   # it did run on the 3xFP16 tensor-core kernels
  # scoring surface: same constant correction
  prec, mup, Cc = OG.posterior_constants(mean.astype(np.float64), sigma.astype(np.float64), w.astype(np.float64))
  X64 = X[:500].astype(np.float64)
  lp = -0.5 * (Cc + X64**2 @ prec - 2 * X64 @ mup + D * np.log(2 * np.pi))
  assert np.max(np.abs(g.logprob(X[:500]) - lp)) < 2e-3 * max(1.0, np.max(np.abs(lp)) / 100)
  assert np.max(np.abs(g.llk(X[:500]) - OG._lse(lp))) < 2e-3 and np.allclose(g.postprob(X[:500]).sum(1), 1.0, atol=1e-4)
  # per-utterance statistics: F-hat index m * D + d
  idx = [("a", (0, 2500)), ("b", (2500, 6000))]
  g.transform_to_disk(X, idx, sad=None)
  Zu, Fu = g.last_utt_stats_
  nr, zr, fr = OG.utterance_stats(X, idx, mean, sigma, w, compute_dtype=np.float64)
  assert Fu.shape == (2, M * D) and relmax(Zu, zr) < TIGHT and relmax(Fu, fr) < 2e-4
  # EM iterations and a mix-up follow the oracle (model started on data points: no empty mixtures, whose
  # variances of exactly zero +- rounding would make the roll-back decision a coin toss)
  M0 = M // 2
  rng = np.random.RandomState(D)
  g2 = GMM(nmix=M, nmix_start=M0, niter=3)
  g2.initialize(X)
  g2.mean = X[rng.choice(N, M0, replace=False)].T.copy()
  g2.sigma = np.tile(X.var(0)[:, None], (1, M0)).astype(np.float32)
  g2.w = np.full((1, M0), 1.0 / M0, dtype=np.float32)
  om, os_, ow = [a.astype(np.float64) for a in (g2.mean, g2.sigma, g2.w)]
  for _ in range(2):
    g2.expectation_maximization(X, print_progress=False)
    zz, ff, ss, ll, _ = OG.expectation(X, om, os_, ow, compute_dtype=np.float64)
    om, os_, ow, rb = OG.maximization(zz, ff, ss, (om, os_, ow))
    assert not rb
  assert relmax(g2.mean, om) < TOL_STATS and relmax(g2.sigma, os_) < TOL_STATS and relmax(g2.w, ow) < TOL_STATS
  g2.gmm_mixup()
  m2, s2, w2 = OG.mixup(om, os_, ow, M)
  assert g2.mean.shape == (D, M) and relmax(g2.mean, m2) < TOL_STATS and relmax(g2.sigma, s2) < TOL_STATS
  g2.expectation_maximization(X, print_progress=False)
  zz, ff, ss, ll, _ = OG.expectation(X, m2, s2, w2, compute_dtype=np.float64)
  m3, s3, w3, rb = OG.maximization(zz, ff, ss, (m2, s2, w2))
  assert not rb and relmax(g2.mean, m3) < TOL_STATS and relmax(g2.sigma, s3) < TOL_STATS


def test_h_float16_store_is_widened_on_the_device():
  """Frames kept as float16 (AsType('float16')) cross PCIe as float16 and are widened by
  odin_feat_convert: the statistics equal those of the same values handed over as float32."""
  import torch
  D, M, N = 60, 256, 300000
  X16 = synth.gmm_features(N, D, 8, seed=3).astype(np.float16)
  mean, sigma, w = synth.gmm_params(D, M, seed=4)
  a = _gmm(M, mean, sigma, w, 3).expectation(X16)
  b = _gmm(M, mean, sigma, w, 3).expectation(X16.astype(np.float32))
  c = _gmm(M, mean, sigma, w, 3).expectation(torch.from_numpy(X16).pin_memory())
  # (the same float32 values reach the kernels on all three routes; the fp64 atomics of the drain commute only up to
  #  rounding, so "equal" is 1e-12, not bit for bit)
  for other in (b, c):
    for x, y in zip(a, other):
      assert relmax(x, y) < 1e-12
  z, f, s, l, n = OG.expectation(X16.astype(np.float64), mean, sigma, w, compute_dtype=np.float64)
  assert max(relmax(a[0], z), relmax(a[1], f), relmax(a[2], s)) < TIGHT


def test_h_config4_protocol_ten_iterations_2048():
  """Config 4's protocol at reduced frame count: TEN EM iterations of a 2048-mixture UBM on resident frames (operand
  images built once) against the fp64 oracle.

  At ~60 frames per mixture the EM map itself amplifies rounding: perturbing the oracle's OWN statistics by 2e-6
  (the accuracy of any float32-input E-step) moves its parameters by 4e-4 after one iteration (sigma = S/N - mu^2
  cancels) and by 5e-3 .. 4e-2 after ten.  So the ten iterations are checked twice: free-running, through the
  log-likelihood trajectory (what the iterations optimise; insensitive to that drift), and iteration by iteration
  from the oracle's parameters (teacher forcing), where the north star's 1e-3 on (mu, sigma^2, w) applies."""
  D, M, N = 60, 2048, 120000
  X = synth.gmm_features(N, D, 64, seed=31)
  rng = np.random.RandomState(6)
  mean = X[rng.choice(N, M, replace=False)].T.copy()
  sigma = np.tile(X.var(0)[:, None], (1, M)).astype(np.float32)
  w = np.full((1, M), 1.0 / M, dtype=np.float32)
  from odin_b200.ml.gmm import _DeviceFrames
  fr = _DeviceFrames(X)
  fr.cache_on_device()
  fr.reuse = True
  free = _gmm(M, mean, sigma, w, 3)      # free-running
  forced = _gmm(M, mean, sigma, w, 3)    # restarted from the oracle's parameters before every iteration
  om, os_, ow = mean.astype(np.float64), sigma.astype(np.float64), w.astype(np.float64)
  llk_oracle = []
  for it in range(10):
    free.expectation_maximization(fr, print_progress=False)
    forced.mean, forced.sigma, forced.w = om.astype(np.float32), os_.astype(np.float32), ow.astype(np.float32)
    z, f, s, l, _ = OG.expectation(X, forced.mean.astype(np.float64), forced.sigma.astype(np.float64),
                                   forced.w.astype(np.float64), compute_dtype=np.float64)
    m1, s1, w1, rb = OG.maximization(z, f, s, (om, os_, ow))
    assert not rb
    forced.expectation_maximization(fr, print_progress=False)
    errs = (relmax(forced.mean, m1), relmax(forced.sigma, s1), relmax(forced.w, w1))
    assert max(errs) < TOL_STATS, (it, errs)
    assert abs(forced._llk_hist[M][-1] - l) < 1e-5 * abs(l)
    llk_oracle.append(float(l))
    om, os_, ow = m1, s1, w1
  hist = free._llk_hist[M]
  assert len(hist) == 10
  assert all(abs(a - b) < 1e-4 * abs(b) for a, b in zip(hist, llk_oracle)), (hist, llk_oracle)
  assert all(b >= a - 1e-6 * abs(a) for a, b in zip(hist, hist[1:]))   # EM never decreases the log-likelihood
