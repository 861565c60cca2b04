"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/*.npz from the REAL
reference (``/root/reference`` executed under ``oracle/ref_shim.py``).

    python -m oracle.make_golden            # in the build container

The reference cannot travel to the GPU box, so its outputs on small seeded
inputs are committed as fixtures together with this script.  Inputs are stored
inside the fixtures so they stay self-contained.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from odin_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _chain(sp, base, cfg):
  steps = [
      sp.AudioReader(remove_dc=True),
      sp.PreEmphasis(0.97),
      sp.STFTExtractor(cfg["frame_length"], cfg["step_length"], n_fft=cfg["n_fft"],
                       window="hamm", energy=True),
      sp.PowerSpecExtractor(2.0, output_name="spec"),
      sp.MelsSpecExtractor(cfg["n_mels"], fmin=cfg["fmin"], fmax=cfg["fmax"]),
      sp.MFCCsExtractor(cfg["n_ceps"], remove_first_coef=True, first_coef_energy=True),
      base.DeltaExtractor("mfcc", order=(0, 1, 2)),
      sp.SADgmm(3, smooth_window=3, input_name="stft_energy", output_name="sad_gmm"),
      sp.SADthreshold(input_name="mfcc_energy", output_name="sad_thr"),
  ]
  return steps


FE_CONFIGS = {
    # SURVEY.md 8d config 1
    "cfg1": dict(sr=16000, frame_length=0.025, step_length=0.010, n_fft=512, n_mels=40,
                 fmin=64, fmax=8000, n_ceps=20, durations=[1.0, 0.64, 1.37]),
    # config 3 front-end
    "cfg3": dict(sr=16000, frame_length=0.025, step_length=0.010, n_fft=1024, n_mels=80,
                 fmin=64, fmax=None, n_ceps=20, durations=[1.21]),
    # config 5 (FSDD recipe, examples/fsdd_ivec.py:80-106)
    "cfg5": dict(sr=8000, frame_length=0.025, step_length=0.005, n_fft=512, n_mels=24,
                 fmin=64, fmax=4000, n_ceps=20, durations=[0.35, 0.8, 0.52]),
}


def frontend_fixtures():
  pp, _ = ref_shim.load_frontend()
  sp, base = pp.speech, pp.base
  for name, cfg in FE_CONFIGS.items():
    blob = {}
    for i, dur in enumerate(cfg["durations"]):
      raw = synth.speech_like(100 * len(name) + i, dur, cfg["sr"], seed=977)
      X = ref_shim.run_pipeline(_chain(sp, base, cfg), {"raw": raw, "sr": cfg["sr"]})
      blob["u%d_pcm" % i] = raw
      blob["u%d_raw" % i] = X["raw"]                       # f32 after DC + pre-emphasis
      blob["u%d_energy" % i] = X["stft_energy"]            # f32 [T,1]
      blob["u%d_mspec" % i] = X["mspec"]                   # f64 [T,n_mels]
      blob["u%d_mfcc" % i] = X["mfcc"]                     # f64 [T,60]
      blob["u%d_c0" % i] = X["mfcc_energy"]                # f64 [T]
      blob["u%d_spec_row0" % i] = X["spec"][0]             # f64 [n_bins]
      blob["u%d_sad_gmm" % i] = X["sad_gmm"].astype(np.uint8)
      blob["u%d_sad_gmm_threshold" % i] = np.float64(X["sad_gmm_threshold"])
      blob["u%d_sad_thr" % i] = X["sad_thr"].astype(np.uint8)
      blob["u%d_sad_thr_threshold" % i] = np.float64(X["sad_thr_threshold"])
    blob["n_utt"] = np.int64(len(cfg["durations"]))
    np.savez_compressed(os.path.join(OUT, "fe_%s.npz" % name), **blob)
    print("wrote fe_%s.npz" % name)
  # SURVEY.md Appendix B input (RNG-free)
  n = np.arange(4800)
  raw = np.round(10000 * np.sin(2 * np.pi * 440 * n / 16000) +
                 3000 * np.sin(2 * np.pi * 1234.5 * n / 16000) * (n > 2400)).astype(np.int16)
  X = ref_shim.run_pipeline(_chain(sp, base, FE_CONFIGS["cfg1"]), {"raw": raw, "sr": 16000})
  np.savez_compressed(
      os.path.join(OUT, "fe_appendix_b.npz"), pcm=raw, raw=X["raw"], energy=X["stft_energy"],
      mspec=X["mspec"], mfcc=X["mfcc"], c0=X["mfcc_energy"],
      sad_gmm=X["sad_gmm"].astype(np.uint8), sad_gmm_threshold=np.float64(X["sad_gmm_threshold"]),
      sad_thr=X["sad_thr"].astype(np.uint8), sad_thr_threshold=np.float64(X["sad_thr_threshold"]))
  print("wrote fe_appendix_b.npz")
  # smoothing quirk: bool vs uint8 routes, reference smooth()
  _, S = ref_shim.load_frontend()
  rng = np.random.RandomState(5)
  cases = [(rng.rand(n_) < p).astype(np.uint8) for n_ in (5, 6, 9, 17, 40) for p in (0.2, 0.5, 0.8)]
  cases.append(np.array([0, 1, 1, 0, 0, 0, 0, 1, 1, 0], dtype=np.uint8))
  blob = {"n": np.int64(len(cases))}
  for i, x in enumerate(cases):
    blob["x%d" % i] = x
    blob["bool3_%d" % i] = (S.smooth(x.astype(bool), win=3, window="flat") >= 2. / 3).astype(np.uint8)
    blob["u8_5_%d" % i] = (S.smooth(x, win=5, window="flat") >= 2. / 5).astype(np.uint8)
  np.savez_compressed(os.path.join(OUT, "smooth.npz"), **blob)
  print("wrote smooth.npz")


SPECTRA_CASES = [
    # (sr, duration, SpectraExtractor kwargs, AudioReader + PreEmphasis in front)
    (16000, 0.9, dict(frame_length=0.025, step_length=0.010, n_fft=512, window="hann", n_mels=40, n_ceps=13), True),
    (16000, 0.5, dict(frame_length=0.025, step_length=0.010, n_fft=1024, window="hamm", log=False, padding=True),
     False),
    (8000, 0.7, dict(frame_length=0.025, step_length=0.005, n_fft=256, window="hann", n_ceps=20, padding=True,
                     fmin=100, fmax=3000), True),
]


def spectra_fixtures():
  """SpectraExtractor (speech.py:849-929) and STFTExtractor(padding=True) from the reference."""
  pp, _ = ref_shim.load_frontend()
  sp, base = pp.speech, pp.base
  blob = {"n": np.int64(len(SPECTRA_CASES))}
  for i, (sr, dur, kw, front) in enumerate(SPECTRA_CASES):
    raw = synth.speech_like(700 + i, dur, sr, seed=311)
    steps = ([sp.AudioReader(remove_dc=True), sp.PreEmphasis(0.97)] if front else []) + [sp.SpectraExtractor(**kw)]
    X = ref_shim.run_pipeline(steps, {"raw": raw if front else raw.astype(np.float32), "sr": sr})
    blob["c%d_pcm" % i] = raw
    for k in ("spec", "energy", "mspec", "mfcc"):
      if X.get(k) is not None:
        blob["c%d_%s" % (i, k)] = X[k]
  # the chained extractors with stft(padding=True)
  cfg = FE_CONFIGS["cfg1"]
  raw = synth.speech_like(731, 0.6, 16000, seed=311)
  steps = _chain(sp, base, cfg)
  steps[2] = sp.STFTExtractor(cfg["frame_length"], cfg["step_length"], n_fft=cfg["n_fft"], window="hamm",
                              energy=True, padding=True)
  X = ref_shim.run_pipeline(steps, {"raw": raw, "sr": 16000})
  blob.update(pad_pcm=raw, pad_energy=X["stft_energy"], pad_mspec=X["mspec"], pad_mfcc=X["mfcc"],
              pad_sad_gmm=X["sad_gmm"].astype(np.uint8))
  np.savez_compressed(os.path.join(OUT, "spectra.npz"), **blob)
  print("wrote spectra.npz")


def gmm_fixtures():
  # (a) Appendix-B RNG-free case, both arithmetic modes of the reference
  N, D, M = 1000, 6, 8
  n = np.arange(N)[:, None]
  d = np.arange(D)[None, :]
  m = np.arange(M)[None, :]
  dd = np.arange(D)[:, None]
  X = (2 * np.sin(0.37 * n + 1.3 * d) + 0.5 * np.cos(0.011 * n * (d + 1))).astype("f")
  mean = (1.5 * np.cos(0.5 * m + 0.2 * dd)).astype("f")
  sigma = (0.5 + 0.25 * (1 + np.sin(m + dd))).astype("f")
  w = ((1 + np.arange(M)) / 36)[None, :].astype("f")
  blob = dict(X=X, mean=mean, sigma=sigma, w=w)
  for tag, mode in (("np2", False), ("f32", True)):
    g = ref_shim.make_ref_gmm(M, float32_mode=mode)
    ref_shim.ref_gmm_initialize(g, X)
    g.mean, g.sigma, g.w = mean.copy(), sigma.copy(), w.copy()
    g._resfresh_cpu_posterior()
    Z, F, S, L = g.expectation(X)
    Zt, Ft = g.transform(X[:100])
    g.maximization(Z, F, S)
    blob.update({tag + "_Z": Z, tag + "_F": F, tag + "_S": S, tag + "_L": np.float64(L),
                 tag + "_Zt": Zt, tag + "_Ft": Ft, tag + "_mean1": g.mean,
                 tag + "_sigma1": g.sigma, tag + "_w1": g.w})
  np.savez_compressed(os.path.join(OUT, "gmm_appendix_b.npz"), **blob)
  print("wrote gmm_appendix_b.npz")

  # (b) seeded D=60, M=64 E-step (config-2 shape, 3000 frames) with and without SAD
  X = synth.gmm_features(3000, 60, 32, seed=11)
  mean, sigma, w = synth.gmm_params(60, 64, seed=12)
  sad = (np.random.RandomState(13).rand(3000) > 0.35).astype(np.uint8)
  blob = dict(X=X, mean=mean, sigma=sigma, w=w, sad=sad)
  for tag, mode in (("np2", False), ("f32", True)):
    g = ref_shim.make_ref_gmm(64, float32_mode=mode)
    ref_shim.ref_gmm_initialize(g, X)
    g.mean, g.sigma, g.w = mean.copy(), sigma.copy(), w.copy()
    g._resfresh_cpu_posterior()
    Z, F, S, L = g.expectation(X)
    Zs, Fs, Ss, Ls = g.expectation(X, sad=sad)
    blob.update({tag + "_Z": Z, tag + "_F": F, tag + "_S": S, tag + "_L": np.float64(L),
                 tag + "_Zsad": Zs, tag + "_Fsad": Fs, tag + "_Ssad": Ss,
                 tag + "_Lsad": np.float64(Ls)})
  # per-utterance statistics (transform) for three ragged "utterances"
  g = ref_shim.make_ref_gmm(64)
  ref_shim.ref_gmm_initialize(g, X)
  g.mean, g.sigma, g.w = mean.copy(), sigma.copy(), w.copy()
  g._resfresh_cpu_posterior()
  bounds = [(0, 700), (700, 1900), (1900, 3000)]
  Zs, Fs = zip(*[g.transform(X[s:e]) for s, e in bounds])
  blob.update(utt_bounds=np.array(bounds), utt_Z=np.concatenate(Zs, 0), utt_Fhat=np.concatenate(Fs, 0))
  np.savez_compressed(os.path.join(OUT, "gmm_d60_m64.npz"), **blob)
  print("wrote gmm_d60_m64.npz")

  # (c) full fit 1 -> 8 mixtures with the reference's split schedule (D=12)
  X = synth.gmm_features(6000, 12, 8, seed=3)
  g = ref_shim.make_ref_gmm(8, nmix_start=1, niter=4)
  g.fit(X)
  hist = dict(g._llk_hist)
  np.savez_compressed(
      os.path.join(OUT, "gmm_fit_d12_m8.npz"), X=X, mean=g.mean, sigma=g.sigma, w=g.w,
      llk_last=np.array([float(hist[k][-1]) for k in (1, 2, 4, 8)]),
      niters=np.array([len(hist[k]) for k in (1, 2, 4, 8)]))
  print("wrote gmm_fit_d12_m8.npz")


def cmvn_fixtures():
  """AcousticNorm (speech.py:1536-1610) of the REAL reference on the MFCC / log-mel fixtures."""
  pp, _ = ref_shim.load_frontend()
  sp = pp.speech
  blob = {}
  k = 0
  for name in ("cfg1", "cfg5"):
    g = np.load(os.path.join(OUT, "fe_%s.npz" % name))
    for i in range(int(g["n_utt"])):
      feat = {"mfcc": g["u%d_mfcc" % i].astype(np.float64), "mspec": g["u%d_mspec" % i].astype(np.float64),
              "sad": g["u%d_sad_gmm" % i].astype(bool)}
      for tag, kw in (("mvn", dict(mean_var_norm=True, windowed_mean_var_norm=False)),
                      ("mvn_novar", dict(mean_var_norm=True, windowed_mean_var_norm=False, var_norm=False)),
                      ("wmvn", dict(mean_var_norm=True, windowed_mean_var_norm=True, win_length=51)),
                      ("wonly", dict(mean_var_norm=False, windowed_mean_var_norm=True, win_length=31)),
                      ("recipe", dict(mean_var_norm=True, windowed_mean_var_norm=True, win_length=301)),
                      ("wmvn_sad", dict(mean_var_norm=True, windowed_mean_var_norm=True, win_length=51,
                                        sad_name="sad"))):
        e = sp.AcousticNorm(input_name=("mspec", "mfcc"), **kw)
        out = e.transform(dict(feat))
        blob["c%d_%s_mfcc" % (k, tag)] = np.asarray(out["mfcc"])
        blob["c%d_%s_mspec" % (k, tag)] = np.asarray(out["mspec"])
      blob["c%d_in_mfcc" % k] = feat["mfcc"]
      blob["c%d_in_mspec" % k] = feat["mspec"]
      blob["c%d_sad" % k] = feat["sad"]
      k += 1
  blob["n_case"] = np.array(k)
  np.savez_compressed(os.path.join(OUT, "cmvn.npz"), **blob)
  print("wrote cmvn.npz (%d cases)" % k)


def variants_fixtures():
  """Framing + CalculateEnergy, RASTAfilter (+ shifted deltas) and StackFeatures from the reference
  (speech.py:569-649, 1483-1533; base.py:724-771)."""
  pp, _ = ref_shim.load_frontend()
  sp, base = pp.speech, pp.base
  blob = {}
  raw = synth.speech_like(801, 0.45, 16000, seed=19)
  for tag, padding in (("nopad", False), ("pad", True)):
    X = ref_shim.run_pipeline([sp.AudioReader(remove_dc=True), sp.PreEmphasis(0.97),
                               sp.Framing(0.025, 0.010, window="hamm", padding=padding), sp.CalculateEnergy(log=True)],
                              {"raw": raw, "sr": 16000})
    blob["fr_%s_frames" % tag] = X["frames"].astype(np.float32)
    blob["fr_%s_energy" % tag] = X["energy"]
    blob["fr_%s_scale" % tag] = np.float64(X["scale"])
  blob["fr_pcm"] = raw
  rng = np.random.RandomState(23)
  for i, (T, F_) in enumerate([(61, 20), (5, 7), (200, 13)]):
    x = np.cumsum(rng.randn(T, F_), axis=0).astype(np.float32) * 0.3 + rng.randn(T, F_).astype(np.float32)
    blob["m%d_x" % i] = x
    blob["m%d_rasta_sdc" % i] = sp.RASTAfilter(rasta=True, sdc=1, input_name="mfcc").transform({"mfcc": x})["mfcc"]
    blob["m%d_rasta" % i] = sp.RASTAfilter(rasta=True, sdc=0, input_name="mfcc").transform({"mfcc": x})["mfcc"]
    blob["m%d_sdc2" % i] = sp.RASTAfilter(rasta=False, sdc=2, input_name="mfcc").transform({"mfcc": x})["mfcc"]
    blob["m%d_stack3" % i] = base.StackFeatures(3, input_name="mfcc").transform({"mfcc": x.copy()})["mfcc"]
  blob["n_mat"] = np.int64(3)
  np.savez_compressed(os.path.join(OUT, "variants.npz"), **blob)
  print("wrote variants.npz")


def _tmat_problem(seed=3, D=6, M=8, n_files=40):
  """Synthetic per-utterance statistics shaped like GMM.transform output (gmm_tmat.py:708-767)."""
  rng = np.random.RandomState(seed)
  sigma = (0.5 + rng.rand(D, M)).astype(np.float64)
  Z = (rng.gamma(2.0, 15.0, size=(n_files, M))).astype(np.float64)
  load = rng.randn(3, M * D) * 0.3                       # a low-rank speaker / session subspace
  F = (rng.randn(n_files, 3).dot(load) * np.repeat(Z, D, axis=1) +
       rng.randn(n_files, M * D) * np.sqrt(np.repeat(Z, D, axis=1))).astype(np.float64)
  return sigma, Z, F


def make_ref_tmatrix(tv_dim, sigma, niter=3):
  """The reference Tmatrix on a stand-in for a fitted GMM (only feat_dim / nmix / sigma are read,
  gmm_tmat.py:1419-1468)."""
  G = ref_shim.load_gmm()
  D, M = sigma.shape
  g = G.GMM(nmix=M, nmix_start=M, niter=1, device="cpu", ncpu=1)
  g._feat_dim, g._is_initialized, g._is_fitted = D, True, True
  g.sigma = sigma
  if not hasattr(G, "defaultdictkey"):
    raise RuntimeError("reference helper defaultdictkey missing")
  G.Tmatrix._refresh_gpu = lambda self: None
  return G.Tmatrix(tv_dim=tv_dim, gmm=g, niter=niter, dtype="float64", device="cpu", ncpu=1)


def tmat_fixtures():
  sigma, Z, F = _tmat_problem()
  tv = 5
  t = make_ref_tmatrix(tv, sigma)
  blob = dict(sigma=sigma, Z=Z, F=F, tv_dim=np.int64(tv), T0=t.Tm.copy(), T_invS0=t.T_invS.copy(),
              T_invS_Tt0=t.T_invS_Tt.copy())
  LU, RU, llk, nframes = t.expectation(Z, F, device="cpu", print_progress=False)
  blob.update(LU0=LU, RU0=RU, llk0=np.float64(llk), nframes0=np.float64(nframes))
  for it in range(3):
    t.expectation_maximization(Z, F, device="cpu", print_progress=False)
    blob["T%d" % (it + 1)] = t.Tm.copy()
  blob["llk_hist"] = np.array(t._llk_hist, dtype=np.float64)
  blob["ivec"] = np.concatenate([t.transform((Z[i:i + 1], F[i:i + 1])) for i in range(Z.shape[0])], 0)
  np.savez_compressed(os.path.join(OUT, "tmat.npz"), **blob)
  print("wrote tmat.npz")


class _IdArray(np.ndarray):
  __ne__ = lambda self, other: self is not other
  __eq__ = lambda self, other: self is other
  __hash__ = None


def downsample_fixtures():
  """Which frames GMM.expectation visits with downsample > 1 (gmm_tmat.py:135-232, 1043-1231), observed by running
  the REAL reference on a data matrix whose first column is the frame number, and the statistics it returns."""
  G = ref_shim.load_gmm()
  rng = np.random.RandomState(21)
  D, M, N = 6, 4, 12000
  X = rng.randn(N, D).astype(np.float32)
  lens = rng.randint(100, 1000, size=20)
  starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
  indices = [("u%02d" % i, (int(s), int(s + n))) for i, (s, n) in enumerate(zip(starts, lens))]
  sad = (rng.rand(N) > 0.3).astype(np.uint8)
  mean, sigma, w = synth.gmm_params(D, M, seed=5)
  blob = dict(X=X, sad=sad, starts=starts, lens=lens, mean=mean, sigma=sigma, w=w)
  for tag, kw in (("ds4", dict(downsample=4, stochastic_downsample=True)),
                  ("ds3det", dict(downsample=3, stochastic_downsample=False))):
    for use_idx in (0, 1):
      g = ref_shim.make_ref_gmm(M, niter=1, batch_size_cpu=700, seed=77, **kw)
      ref_shim.ref_gmm_initialize(g, X)
      g.mean, g.sigma, g.w = mean.copy(), sigma.copy(), w.copy()
      g._resfresh_cpu_posterior()
      g._llk_hist[M] = [0.0, 0.0]           # pretend two EM iterations were done: curr_niter = 2 enters the seed
      # The workers of `expectation` are forked processes, so the picks are observed by driving the reference's own
      # batch generators the way its single default job does (gmm_tmat.py:1165-1181: the job takes the files popped
      # from the end of the list) on a matrix whose first column is the frame number ...
      Xn = X.copy()
      Xn[:, 0] = np.arange(N)
      bs = int(g.batch_size_cpu / np.floor(np.power(2, M / 1024)))
      common = dict(batch_size=bs, downsample=g.downsample, stochastic=g.stochastic_downsample, seed=g._seed,
                    curr_nmix=M, curr_niter=2)
      if use_idx:
        it = G._create_batch_indices(Xn, sad, list(indices)[::-1], **common)
      else:
        it = G._create_batch(Xn, sad, 0, N, **common)
      rows = np.concatenate([y[:, 0] for y, _, _ in it if y is not None]).astype(np.int64)
      picked = np.zeros(N, dtype=np.uint8)
      picked[rows] = 1
      blob["%s_idx%d_mask" % (tag, use_idx)] = picked
      # ... and confirmed against the real call: its statistics are those of exactly these frames
      # (GMM.initialize picks the index list out of the tuple with `i != tmp` (gmm_tmat.py:566), which is ambiguous for a
      #  plain ndarray; the reference feeds its own MmapData there.  An ndarray view with identity comparison stands in.)
      Z, F, S, L = g.expectation((X.view(_IdArray), list(indices)) if use_idx else X, sad=sad, print_progress=True)
      z2, f2, s2, l2 = g._fast_expectation(X[picked.astype(bool)], True, True, True, True, on_gpu=False)
      assert np.allclose(Z, z2, rtol=1e-5) and np.allclose(F, f2, rtol=1e-4, atol=1e-3) and np.allclose(S, s2, rtol=1e-4, atol=1e-3)
      assert abs(float(L) - float(l2) / int(picked.sum())) < 1e-5 * abs(float(L))
      blob["%s_idx%d_Z" % (tag, use_idx)], blob["%s_idx%d_F" % (tag, use_idx)] = Z, F
      blob["%s_idx%d_S" % (tag, use_idx)], blob["%s_idx%d_L" % (tag, use_idx)] = S, L
  np.savez_compressed(os.path.join(OUT, "gmm_downsample.npz"), **blob)
  print("wrote gmm_downsample.npz")


def stages_fixtures():
  """The chain one stage at a time through the reference's free functions (signal.py:955-967, 1002-1066,
  1421-1562, 1623-1716): what `odin_b200.preprocessing.signal.*` and the stand-alone extractors must return."""
  pp, S = ref_shim.load_frontend()
  sr = 16000
  raw = synth.speech_like(4242, 0.6, sr, seed=977).astype(np.float32)
  blob = {"pcm": raw}
  pre = S.pre_emphasis(raw, 0.97)
  blob["pre"] = pre
  blob["pre2d"] = S.pre_emphasis(np.stack([raw[:4000], raw[4000:8000]]), 0.95)
  st, en = S.stft(pre, frame_length=400, step_length=160, n_fft=512, window='hamm', energy=True)
  blob["stft"], blob["stft_energy"] = st, en
  blob["stft_pad"] = S.stft(raw, frame_length=400, step_length=240, n_fft=1024, window='hann', padding=True)
  blob["stft_scale"] = S.stft(raw, frame_length=200, step_length=80, n_fft=256, window='hann', scale=0.5)
  spec = S.power_spectrogram(st, 2.0)
  blob["spec"], blob["spec_mag"] = spec, S.power_spectrogram(st, 1.0)
  blob["spec_real3"] = S.power_spectrogram(np.abs(st)[:5], 3.0)
  mspec = S.mels_spectrogram(spec, sr, 40, fmin=64, fmax=8000, top_db=80.0)
  blob["mspec"] = mspec
  blob["mspec_top20"] = S.mels_spectrogram(spec, sr, 24, fmin=100, fmax=None, top_db=20.0)
  blob["mfcc"] = S.ceps_spectrogram(mspec, 20, remove_first_coef=True)
  blob["mfcc_keep0"] = S.ceps_spectrogram(mspec, 13, remove_first_coef=False)
  d1, d2 = S.delta(blob["mfcc"], width=9, order=2, axis=0)
  blob["d1"], blob["d2"] = d1, d2
  blob["d1_w5"] = S.delta(blob["mfcc"], width=5, order=1, axis=0)
  blob["d1_vec"] = S.delta(blob["mfcc"][:, 3], width=9, order=1, axis=0)
  blob["energy_frames"] = S.get_energy(np.lib.stride_tricks.sliding_window_view(pre, 400)[::160], log=True)
  # the same stages as stand-alone extractors (speech.py:540-563, 655-831; base.py:433-484)
  sp, base = pp.speech, pp.base
  X = {"raw": raw, "sr": sr}
  X = sp.PreEmphasis(0.97).transform(X)
  X = sp.STFTExtractor(0.025, 0.010, n_fft=512, window='hamm', energy=True).transform(X)
  X = sp.PowerSpecExtractor(2.0).transform(X)
  X = sp.MelsSpecExtractor(40, fmin=64, fmax=8000).transform(X)
  X = sp.MFCCsExtractor(20, remove_first_coef=True, first_coef_energy=True).transform(X)
  X = base.DeltaExtractor('mfcc', order=(0, 1, 2)).transform(X)
  blob["x_mfcc"], blob["x_mfcc_energy"], blob["x_stft_energy"] = X["mfcc"], X["mfcc_energy"], X["stft_energy"]
  # stored at single precision (the tolerances are 1e-4 of the matrix maximum): keeps the fixture small
  blob = {k: (v.astype(np.complex64) if np.iscomplexobj(v) else v.astype(np.float32)) for k, v in blob.items()}
  np.savez_compressed(os.path.join(OUT, "stages.npz"), **blob)
  print("wrote stages.npz")


def _scoring_problem(seed=5, ncls=10, per=30, d=24):
  """Synthetic i-vector-like vectors: class means + within-class noise, ragged class sizes."""
  rng = np.random.RandomState(seed)
  mu = rng.randn(ncls, d) * 1.5
  y = np.concatenate([np.full(per + 3 * (c % 3), c) for c in range(ncls)])
  X = mu[y] + rng.randn(len(y), d)
  yt = np.repeat(np.arange(ncls), 7)
  Xt = mu[yt] + rng.randn(len(yt), d)
  return X, y, Xt, yt


def scoring_fixtures():
  """PLDA (plda.py:215-423) and Scorer / VectorNormalizer (scoring.py:95-364) of the REAL reference."""
  SC, PL = ref_shim.load_scoring()
  X, y, Xt, yt = _scoring_problem()
  blob = dict(X=X, y=y, Xt=Xt, yt=yt)
  p = PL.PLDA(n_phi=8, n_iter=12, centering=True, wccn=True, unit_length=True, random_state=1234)
  p.fit(X, y)
  blob.update(plda_scores=p.predict_log_proba(Xt), plda_Phi=p.Phi_, plda_Sigma=p.Sigma_, plda_Uk=p.Uk_,
              plda_Lambda=p.Lambda_, plda_Qhat=p.Q_hat_, plda_Xmodel=p.X_model_, plda_proj=p.transform(Xt),
              plda_llk=np.float64(p.compute_llk(p.normalizer.transform(X))))
  # (n_iter='auto' evaluates compute_llk every iteration, whose Cholesky fails on the non-symmetric Sigma_ of the
  #  first M-steps for every seed tried here: the reference raises LinAlgError; tests/test_scoring.py expects the same)
  try:
    PL.PLDA(n_phi=8, n_iter='auto', improve_threshold=1e-1, random_state=7).fit(X, y)
    blob["plda_auto_raises"] = np.int64(0)
  except np.linalg.LinAlgError:
    blob["plda_auto_raises"] = np.int64(1)
  pm = PL.PLDA(n_phi=8, random_state=3).fit_maximum_likelihood(X, y)
  blob["plda_ml_scores"] = pm.predict_log_proba(Xt)
  blob["plda_enroll_scores"] = p.predict_log_proba(Xt, X_model=SC.compute_class_avg(Xt, yt, np.unique(yt)))
  for lda in (True, False):
    s = SC.Scorer(centering=True, wccn=True, lda=lda, method='cosine').fit(X, y)
    blob["cos_scores_lda%d" % lda] = s.transform(Xt)
    blob["cos_enroll_lda%d" % lda] = s.normalizer.enroll_vecs
  vn = SC.VectorNormalizer(centering=True, wccn=True, unit_length=True, lda=True, concat=True).fit(X, y)
  blob["vn_concat"] = vn.transform(Xt)
  blob["vn_W"], blob["vn_mean"], blob["vn_vmin"], blob["vn_vmax"] = vn.W, vn.mean, vn.vmin, vn.vmax
  np.savez_compressed(os.path.join(OUT, "scoring.npz"), **blob)
  print("wrote scoring.npz")


if __name__ == "__main__":
  warnings.filterwarnings("ignore")
  os.makedirs(OUT, exist_ok=True)
  if len(sys.argv) > 1 and sys.argv[1] == "cmvn":
    cmvn_fixtures()
  elif len(sys.argv) > 1 and sys.argv[1] == "spectra":
    spectra_fixtures()
  elif len(sys.argv) > 1 and sys.argv[1] == "tmat":
    tmat_fixtures()
  elif len(sys.argv) > 1 and sys.argv[1] == "variants":
    variants_fixtures()
  elif len(sys.argv) > 1 and sys.argv[1] == "scoring":
    scoring_fixtures()
  elif len(sys.argv) > 1 and sys.argv[1] == "stages":
    stages_fixtures()
  elif len(sys.argv) > 1 and sys.argv[1] == "downsample":
    downsample_fixtures()
  else:
    downsample_fixtures()
    stages_fixtures()
    scoring_fixtures()
    variants_fixtures()
    tmat_fixtures()
    spectra_fixtures()
    frontend_fixtures()
    gmm_fixtures()
    cmvn_fixtures()
