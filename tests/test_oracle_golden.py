"""CPU: the oracle (oracle/*.py) against the committed golden vectors, which
were produced by the REAL reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relmax
from oracle import frontend as F
from oracle import gmm as OG
from oracle.make_golden import FE_CONFIGS


@pytest.mark.parametrize("name", sorted(FE_CONFIGS))
def test_frontend_golden(name):
  cfg = FE_CONFIGS[name]
  g = np.load(os.path.join(GOLDEN, "fe_%s.npz" % name))
  for i in range(int(g["n_utt"])):
    pcm = g["u%d_pcm" % i]
    o = F.extract(pcm, cfg["sr"], cfg["frame_length"], cfg["step_length"], cfg["n_fft"],
                  n_mels=cfg["n_mels"], fmin=cfg["fmin"], fmax=cfg["fmax"], n_ceps=cfg["n_ceps"],
                  vad="gmm")
    L, hop = F.frame_step_length(cfg["sr"], cfg["frame_length"], cfg["step_length"])
    T = g["u%d_mfcc" % i].shape[0]
    assert F.num_frames(len(pcm), L, hop) == T == 1 + (len(pcm) - L) // hop
    assert np.array_equal(o["raw"], g["u%d_raw" % i])              # f32 stages bit-exact
    assert np.array_equal(o["stft_energy"], g["u%d_energy" % i])
    assert relmax(o["spec"][0], g["u%d_spec_row0" % i]) < 1e-12
    assert relmax(o["mspec"], g["u%d_mspec" % i]) < 1e-12
    assert relmax(o["mfcc"], g["u%d_mfcc" % i]) < 1e-12
    assert relmax(o["mfcc_energy"], g["u%d_c0" % i]) < 1e-12
    assert np.array_equal(o["sad"], g["u%d_sad_gmm" % i])
    assert abs(o["sad_threshold"] - float(g["u%d_sad_gmm_threshold" % i])) < 1e-9
    s2, t2 = F.sad_threshold(o["mfcc_energy"])
    assert np.array_equal(s2.astype(np.uint8), g["u%d_sad_thr" % i])
    assert abs(t2 - float(g["u%d_sad_thr_threshold" % i])) < 1e-9


def test_frontend_appendix_b():
  """SURVEY.md Appendix B self-check values, RNG-free input."""
  g = np.load(os.path.join(GOLDEN, "fe_appendix_b.npz"))
  o = F.extract(g["pcm"], 16000, vad="gmm", fmax=8000)
  assert o["mfcc"].shape == (28, 60)
  np.testing.assert_allclose(o["raw"][:3], [-0.78729165, 1718.9764, 1719.5463], rtol=1e-6)
  np.testing.assert_allclose(o["mspec"][0, :4], [-10.710571033, -4.716102060, -6.872628058,
                                                29.096165021], rtol=1e-8)
  np.testing.assert_allclose([o["mspec"].max(), o["mspec"].min()], [40.946535784, -39.053464216],
                             rtol=1e-9)
  np.testing.assert_allclose(o["mfcc"][0, :3], [67.745275783, 17.927752977, 8.993522132], rtol=1e-8)
  np.testing.assert_allclose(o["mfcc"][10, 40:42], [-0.709244549, -0.416948587], rtol=1e-6)
  np.testing.assert_allclose(o["mfcc"][0, 40:42], [-0.766150832, -0.146533281], rtol=1e-6)
  assert "".join(map(str, o["sad"])) == "0000000000000001111111111111"
  assert abs(o["sad_threshold"] - 1.02282264) < 1e-7
  s2, t2 = F.sad_threshold(o["mfcc_energy"])
  assert "".join(map(str, s2.astype(int))) == "0000000000111111110000000000"
  assert abs(t2 - 0.73292095) < 1e-7
  assert relmax(o["mfcc"], g["mfcc"]) < 1e-12


def test_smooth_golden():
  g = np.load(os.path.join(GOLDEN, "smooth.npz"))
  for i in range(int(g["n"])):
    x = g["x%d" % i]
    assert np.array_equal((F.smooth_flat(x.astype(bool), 3) >= 2. / 3).astype(np.uint8), g["bool3_%d" % i])
    assert np.array_equal((F.smooth_flat(x, 5) >= 2. / 5).astype(np.uint8), g["u8_5_%d" % i])
  # SURVEY.md 8.1-Q2
  x = np.array([0, 1, 1, 0, 0, 0, 0, 1, 1, 0], dtype=np.uint8)
  assert "".join(map(str, (F.smooth_flat(x, 5) >= 0.4).astype(int))) == "1111001111"
  assert "".join(map(str, (F.smooth_flat(x.astype(bool), 5) >= 0.4).astype(int))) == "0111001110"


def test_delta_closed_form():
  """order-2 deltas are delayed by 4 frames."""
  rng = np.random.RandomState(0)
  x = rng.randn(50, 3)
  d1, d2 = F.deltas(x, 9, 2)
  T = x.shape[0]
  cl = lambda t: min(max(t, 0), T - 1)
  dext = lambda u: sum((m / 60.0) * x[cl(u + m)] for m in range(-4, 5))
  for t in (0, 1, 5, 20, 49):
    np.testing.assert_allclose(d1[t], dext(t), atol=1e-6)
  for t in (3, 10, 30, 49):
    ref = sum(((4 - k) / 60.0) * dext(t - k) for k in range(9))
    np.testing.assert_allclose(d2[t], ref, atol=1e-6)


@pytest.mark.parametrize("tag,dt", [("np2", None), ("f32", np.float32)])
def test_gmm_appendix_b(tag, dt):
  g = np.load(os.path.join(GOLDEN, "gmm_appendix_b.npz"))
  Z, F_, S, L, n = OG.expectation(g["X"], g["mean"], g["sigma"], g["w"], compute_dtype=dt)
  assert np.array_equal(Z, g[tag + "_Z"]) and np.array_equal(F_, g[tag + "_F"])
  assert np.array_equal(S, g[tag + "_S"]) and float(L) == float(g[tag + "_L"])
  Zt, Ft = OG.transform(g["X"][:100], g["mean"], g["sigma"], g["w"], compute_dtype=dt)
  assert np.array_equal(Zt, g[tag + "_Zt"]) and np.array_equal(Ft, g[tag + "_Ft"])
  m1, s1, w1, rb = OG.maximization(Z, F_, S, (g["mean"], g["sigma"], g["w"]))
  assert not rb
  assert np.array_equal(m1, g[tag + "_mean1"]) and np.array_equal(s1, g[tag + "_sigma1"])
  assert np.array_equal(w1, g[tag + "_w1"])
  if tag == "np2":
    np.testing.assert_allclose(Z[0, :3], [37.721307852, 76.638436937, 396.629982095], rtol=1e-9)
    assert abs(float(L) - (-15.783989418815855)) < 1e-12
    np.testing.assert_allclose(Ft[0, :4], [2.100726751, -10.095958170, -17.500275960, -8.928262013],
                               rtol=1e-8)


@pytest.mark.parametrize("tag,dt", [("np2", None), ("f32", np.float32)])
def test_gmm_d60_m64(tag, dt):
  g = np.load(os.path.join(GOLDEN, "gmm_d60_m64.npz"))
  Z, F_, S, L, n = OG.expectation(g["X"], g["mean"], g["sigma"], g["w"], compute_dtype=dt)
  assert np.array_equal(Z, g[tag + "_Z"]) and np.array_equal(F_, g[tag + "_F"])
  assert np.array_equal(S, g[tag + "_S"]) and float(L) == float(g[tag + "_L"])
  Z, F_, S, L, n = OG.expectation(g["X"], g["mean"], g["sigma"], g["w"], sad=g["sad"],
                                  compute_dtype=dt)
  assert n == int(g["sad"].sum())
  assert np.array_equal(Z, g[tag + "_Zsad"]) and np.array_equal(S, g[tag + "_Ssad"])
  assert float(L) == float(g[tag + "_Lsad"])
  # the float64 checker sits within 1e-5 of both reference modes
  Z64, F64, S64, L64, _ = OG.expectation(g["X"], g["mean"], g["sigma"], g["w"], sad=g["sad"],
                                         compute_dtype=np.float64)
  assert relmax(Z64, Z) < 1e-5 and relmax(F64, F_) < 1e-5 and relmax(S64, S) < 1e-5


def test_gmm_utterance_stats():
  g = np.load(os.path.join(GOLDEN, "gmm_d60_m64.npz"))
  idx = [("u%d" % i, (int(s), int(e))) for i, (s, e) in enumerate(g["utt_bounds"])]
  names, Z, Fh = OG.utterance_stats(g["X"], idx, g["mean"], g["sigma"], g["w"])
  assert names == ["u0", "u1", "u2"]
  assert np.array_equal(Z, g["utt_Z"]) and np.array_equal(Fh, g["utt_Fhat"])
  assert Fh.shape == (3, 60 * 64)


def test_gmm_fit_schedule():
  g = np.load(os.path.join(GOLDEN, "gmm_fit_d12_m8.npz"))
  mean, sigma, w, hist = OG.fit(g["X"], 8, niter=4)
  assert [len(hist[k]) for k in (1, 2, 4, 8)] == list(g["niters"]) == [1, 2, 4, 4]
  assert np.array_equal(mean, g["mean"]) and np.array_equal(sigma, g["sigma"])
  assert np.array_equal(w, g["w"])
  np.testing.assert_allclose([hist[k][-1] for k in (1, 2, 4, 8)], g["llk_last"], rtol=1e-12)


def test_minibatch_and_batch_size():
  assert OG.minibatch_ranges(10, 4) == [(0, 4), (4, 8), (8, 10)]
  assert OG.default_batch_size(60, 64) == 52428
  assert OG.default_batch_size(60, 2048) == 13107     # / floor(2^2)


CMVN_TAGS = {
    "mvn": dict(),
    "mvn_novar": dict(var_norm=False),
    "wmvn": dict(windowed_mean_var_norm=True, win_length=51),
    "wonly": dict(mean_var_norm=False, windowed_mean_var_norm=True, win_length=31),
    "recipe": dict(windowed_mean_var_norm=True, win_length=301),
    "wmvn_sad": dict(windowed_mean_var_norm=True, win_length=51),
}


def test_cmvn_golden():
  """oracle.acoustic_norm against AcousticNorm of the real reference (tests/golden/cmvn.npz)."""
  g = np.load(os.path.join(GOLDEN, "cmvn.npz"))
  for k in range(int(g["n_case"])):
    for tag, kw in CMVN_TAGS.items():
      for f in ("mfcc", "mspec"):
        sad = g["c%d_sad" % k] if tag == "wmvn_sad" else None
        y = F.acoustic_norm(g["c%d_in_%s" % (k, f)], sad=sad, **kw)
        r = g["c%d_%s_%s" % (k, tag, f)]
        assert np.array_equal(np.isnan(y), np.isnan(r))
        ok = np.isfinite(r)
        assert np.max(np.abs(y[ok] - r[ok])) < 1e-12


def _spectra_inputs(g, i):
  """The signal SpectraExtractor saw in fixture case i (oracle/make_golden.py SPECTRA_CASES)."""
  from oracle.make_golden import SPECTRA_CASES
  sr, _, kw, front = SPECTRA_CASES[i]
  pcm = g["c%d_pcm" % i]
  y = F.pre_emphasis(F.read_audio(pcm, True), 0.97) if front else pcm.astype(np.float32)
  return sr, kw, front, pcm, y


def test_spectra_golden():
  """SpectraExtractor (speech.py:849-929) and stft(padding=True): oracle vs the reference's outputs."""
  g = np.load(os.path.join(GOLDEN, "spectra.npz"))
  for i in range(int(g["n"])):
    sr, kw, _, _, y = _spectra_inputs(g, i)
    o = F.spectra(y, sr, **kw)
    for k in ("spec", "energy", "mspec", "mfcc"):
      key = "c%d_%s" % (i, k)
      if key in g.files:
        assert o[k].dtype == np.float32 and o[k].shape == g[key].shape, (i, k)
        assert relmax(o[k], g[key]) < 1e-6, (i, k)   # f32 casts of f64 pipelines
      else:
        assert o[k] is None
  o = F.extract(g["pad_pcm"], 16000, vad="gmm", fmax=8000, padding=True)
  assert o["mfcc"].shape == g["pad_mfcc"].shape == (1 + (9600 + 2 * 200 - 400) // 160, 60)
  assert np.array_equal(o["stft_energy"], g["pad_energy"])
  assert relmax(o["mspec"], g["pad_mspec"]) < 1e-12 and relmax(o["mfcc"], g["pad_mfcc"]) < 1e-12
  assert np.array_equal(o["sad"], g["pad_sad_gmm"])


def test_tmatrix_golden():
  """oracle/tmatrix.py against the reference's Tmatrix (gmm_tmat.py:1343-2090): initial statistics, one
  E-step, three EM iterations and the i-vectors of the training files."""
  from oracle import tmatrix as OT
  g = np.load(os.path.join(GOLDEN, "tmat.npz"))
  sigma, Z, F, tv = g["sigma"], g["Z"], g["F"], int(g["tv_dim"])
  D = sigma.shape[0]
  Sigma = OT.sigma_row(sigma)
  T0 = OT.init_T(tv, Sigma)
  assert np.array_equal(T0, g["T0"])
  T_invS, T_invS_Tt = OT.refresh(T0, Sigma, D)
  assert relmax(T_invS, g["T_invS0"]) < 1e-14 and relmax(T_invS_Tt, g["T_invS_Tt0"]) < 1e-14
  LU, RU, llk, nframes = OT.expectation(Z, F, T_invS, T_invS_Tt)
  assert relmax(LU, g["LU0"]) < 1e-12 and relmax(RU, g["RU0"]) < 1e-12
  assert abs(llk - float(g["llk0"])) < 1e-9 * abs(float(g["llk0"])) and nframes == float(g["nframes0"])
  Tm, T_invS, T_invS_Tt, hist = OT.fit(Z, F, tv, sigma, 3)
  assert relmax(Tm, g["T3"]) < 1e-9
  assert np.allclose(hist, g["llk_hist"], rtol=1e-9)
  assert relmax(OT.ivector(Z, F, T_invS, T_invS_Tt), g["ivec"]) < 1e-9


def test_feature_variants_golden():
  """Framing / CalculateEnergy / RASTAfilter (+ shifted deltas) / StackFeatures: oracle vs the reference's outputs."""
  g = np.load(os.path.join(GOLDEN, "variants.npz"))
  y = F.pre_emphasis(F.read_audio(g["fr_pcm"], True), 0.97)
  for tag, padding in (("nopad", False), ("pad", True)):
    fr, scale = F.framing(y, 16000, 0.025, 0.010, "hamm", padding)
    assert fr.shape == g["fr_%s_frames" % tag].shape
    assert relmax(fr, g["fr_%s_frames" % tag]) < 1e-6 and abs(scale - float(g["fr_%s_scale" % tag])) < 1e-15
    assert np.array_equal(F.frame_energy(fr), g["fr_%s_energy" % tag])
  for i in range(int(g["n_mat"])):
    x = g["m%d_x" % i]
    assert relmax(F.rasta_sdc(x, True, 1), g["m%d_rasta_sdc" % i]) < 1e-6
    assert relmax(F.rasta_sdc(x, True, 0), g["m%d_rasta" % i]) < 1e-6
    assert relmax(F.rasta_sdc(x, False, 2), g["m%d_sdc2" % i]) < 1e-6
    assert np.array_equal(F.stack_context(x, 3), g["m%d_stack3" % i])
