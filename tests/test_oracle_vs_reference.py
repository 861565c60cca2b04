"""CPU, build container only: the oracle against the REAL reference executed
under oracle/ref_shim.py on fresh seeded inputs.  Skipped where the reference
checkout is absent (the GPU box)."""
import warnings

import numpy as np
import pytest

from conftest import relmax
from odin_b200 import synth
from oracle import frontend as F
from oracle import gmm as OG
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout absent")


def _ref_chain(raw, sr, cfg):
  from oracle.make_golden import _chain
  pp, _ = ref_shim.load_frontend()
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    return ref_shim.run_pipeline(_chain(pp.speech, pp.base, cfg), {"raw": raw, "sr": sr})


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_frontend_matches_reference(seed):
  from oracle.make_golden import FE_CONFIGS
  name = ["cfg1", "cfg3", "cfg5", "cfg1"][seed]
  cfg = FE_CONFIGS[name]
  raw = synth.speech_like(seed, 0.9 + 0.4 * seed, cfg["sr"], seed=31)
  R = _ref_chain(raw, cfg["sr"], cfg)
  o = F.extract(raw, cfg["sr"], cfg["frame_length"], cfg["step_length"], cfg["n_fft"],
                n_mels=cfg["n_mels"], fmin=cfg["fmin"], fmax=cfg["fmax"], vad="gmm")
  assert o["mfcc"].shape == R["mfcc"].shape
  assert np.array_equal(o["raw"], R["raw"])
  assert np.array_equal(o["stft_energy"], R["stft_energy"])
  assert relmax(o["mspec"], R["mspec"]) < 1e-12
  assert relmax(o["mfcc"], R["mfcc"]) < 1e-12
  assert np.array_equal(o["sad"], R["sad_gmm"])
  assert abs(o["sad_threshold"] - R["sad_gmm_threshold"]) < 1e-9
  s2, t2 = F.sad_threshold(o["mfcc_energy"])
  assert np.array_equal(s2, R["sad_thr"]) and abs(t2 - R["sad_thr_threshold"]) < 1e-9


def test_frontend_long_utterance_matches_reference():
  """Config 3 at the benchmark's utterance lengths (60 s: 5 998 frames, utterance-global top_db and SADgmm over thousands
  of frames) and on the float32 samples soundfile would hand over -- pins the oracle the GPU test
  `test_config3_full_length_utterances_vs_oracle` compares against."""
  from oracle.make_golden import FE_CONFIGS
  cfg = FE_CONFIGS["cfg3"]
  raw = synth.speech_like(700, 60.0, cfg["sr"], seed=99)
  for x in (raw, (raw[:cfg["sr"] * 8].astype(np.float64) / 32768.0).astype(np.float32)):
    R = _ref_chain(x, cfg["sr"], cfg)
    o = F.extract(x, cfg["sr"], cfg["frame_length"], cfg["step_length"], cfg["n_fft"],
                  n_mels=cfg["n_mels"], fmin=cfg["fmin"], fmax=cfg["fmax"], vad="gmm")
    assert o["mfcc"].shape == R["mfcc"].shape
    assert np.array_equal(o["stft_energy"], R["stft_energy"])
    assert relmax(o["mspec"], R["mspec"]) < 1e-12 and relmax(o["mfcc"], R["mfcc"]) < 1e-12
    assert np.array_equal(o["sad"], R["sad_gmm"])


def test_mel_and_dct_tables():
  _, S = ref_shim.load_frontend()
  for sr, n_fft, n_mels, fmin, fmax in [(16000, 512, 40, 64, 8000), (16000, 1024, 80, 64, 8000),
                                        (8000, 512, 24, 64, 4000), (8000, 256, 24, 0, 3800)]:
    assert np.allclose(F.mel_filterbank(sr, n_fft, n_mels, fmin, fmax),
                       S.mel_filters(sr, n_fft, n_mels, fmin, fmax), rtol=0, atol=1e-15)
  assert np.allclose(F.dct_basis(21, 40), S.dct_filters(21, 40), rtol=0, atol=1e-14)
  for name in ("hamm", "hann"):
    assert np.allclose(F.window_table(name, 400), S.get_window(name, 400), rtol=0, atol=1e-15)


def test_smooth_exhaustive_short():
  """every 0/1 sequence of length 5..9 through the reference smooth(), both
  dtype routes."""
  _, S = ref_shim.load_frontend()
  for n in range(5, 10):
    for code in range(2**n):
      x = np.array([(code >> i) & 1 for i in range(n)], dtype=np.uint8)
      for win, thr in ((3, 2. / 3), (5, 2. / 5)):
        assert np.array_equal(S.smooth(x, win=win, window="flat") >= thr,
                              F.smooth_flat(x, win) >= thr)
        assert np.array_equal(S.smooth(x.astype(bool), win=win, window="flat") >= thr,
                              F.smooth_flat(x.astype(bool), win) >= thr)


def test_gmm_estep_and_fit_match_reference():
  X = synth.gmm_features(5000, 20, 8, seed=77)
  mean, sigma, w = synth.gmm_params(20, 16, seed=78)
  for mode, dt in ((False, None), (True, np.float32)):
    g = ref_shim.make_ref_gmm(16, float32_mode=mode)
    ref_shim.ref_gmm_initialize(g, X)
    g.mean, g.sigma, g.w = mean.copy(), sigma.copy(), w.copy()
    g._resfresh_cpu_posterior()
    Z, Fs, S, L = g.expectation(X)
    z, f, s, l, _ = OG.expectation(X, mean, sigma, w, compute_dtype=dt)
    assert np.array_equal(Z, z) and np.array_equal(Fs, f) and np.array_equal(S, s)
    assert float(L) == float(l)
  g = ref_shim.make_ref_gmm(4, nmix_start=1, niter=3)
  g.fit(X)
  m, s, ww, hist = OG.fit(X, 4, niter=3)
  assert np.array_equal(g.mean, m) and np.array_equal(g.sigma, s) and np.array_equal(g.w, ww)


@pytest.mark.parametrize("padding", [False, True])
def test_spectra_matches_reference(padding):
  """oracle.spectra against the reference's SpectraExtractor (speech.py:849-929) on fresh audio."""
  pp, _ = ref_shim.load_frontend()
  raw = synth.speech_like(5, 0.8, 16000, seed=41).astype(np.float32)
  kw = dict(frame_length=0.025, step_length=0.010, n_fft=512, window="hann", n_mels=30, n_ceps=12, padding=padding)
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    R = ref_shim.run_pipeline([pp.speech.SpectraExtractor(**kw)], {"raw": raw, "sr": 16000})
  o = F.spectra(raw, 16000, **kw)
  for k in ("spec", "energy", "mspec", "mfcc"):
    assert o[k].shape == R[k].shape and o[k].dtype == R[k].dtype
    assert relmax(o[k], R[k]) < 1e-6, k


def test_tmatrix_matches_reference():
  """oracle/tmatrix.py against the reference class on a fresh problem (tv 7, 5 mixtures, 2 EM iterations)."""
  from oracle import tmatrix as OT
  from oracle.make_golden import _tmat_problem, make_ref_tmatrix
  sigma, Z, F = _tmat_problem(seed=11, D=4, M=5, n_files=25)
  t = make_ref_tmatrix(7, sigma, niter=2)
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for _ in range(2):
      t.expectation_maximization(Z, F, device="cpu", print_progress=False)
  Tm, T_invS, T_invS_Tt, hist = OT.fit(Z, F, 7, sigma, 2)
  assert relmax(Tm, t.Tm) < 1e-9 and relmax(T_invS_Tt, t.T_invS_Tt) < 1e-9
  assert np.allclose(hist, t._llk_hist, rtol=1e-9)
  iv = OT.ivector(Z[:3], F[:3], T_invS, T_invS_Tt)
  ref = np.concatenate([t.transform((Z[i:i + 1], F[i:i + 1])) for i in range(3)], 0)
  assert relmax(iv, ref) < 1e-9


@pytest.mark.parametrize("seed,mx,mn,thr", [(5, 5, None, 0.6), (6, 3, 1.0, 0.5)])
def test_vad_split_audio_matches_reference(seed, mx, mn, thr):
  _, S = ref_shim.load_frontend()
  s = np.concatenate(synth.utterance_batch(3, 4.0, 6.0, sr=8000, seed=seed)).astype(np.float32)
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    segs, vad, voices, cut = S.vad_split_audio(s, 8000, maximum_duration=mx, minimum_duration=mn, frame_length=128,
                                               nb_mixtures=3, threshold=thr, return_vad=True, return_voices=True,
                                               return_cut=True)
  o = F.vad_split_audio(s, 8000, mx, mn, 128, 3, thr)
  assert len(segs) == len(o[0]) and all(np.array_equal(a, b) for a, b in zip(segs, o[0]))
  assert np.array_equal(vad, o[1]) and np.array_equal(voices, o[2]) and np.array_equal(cut, o[3])
