"""Config 5 (BASELINE.json): FSDD-style end-to-end pipeline on the GPU path --
8 kHz synthetic digits -> MFCC (+c0 energy, deltas, SADthreshold, recipe of
examples/fsdd_ivec.py:80-106 minus AcousticNorm / AsType) -> FeatureProcessor store with
name -> (start, end) indices -> 512-mix UBM statistics -> per-utterance Z [n, 512] and
F-hat [n, 512*60] (the i-vector input of gmm_tmat.py:769-913) -> T-matrix EM and i-vectors (gmm_tmat.py:1343-2090).
Bit-exact: frame counts, `sad`, indices.  <= 1e-4 on MFCC, <= 1e-3 on Z / F-hat."""
import os

import numpy as np
import pytest

from conftest import relmax
from odin_b200 import synth
from oracle import frontend as F
from oracle import gmm as OG
from oracle.make_golden import FE_CONFIGS

pytestmark = pytest.mark.gpu


def test_fsdd_style_pipeline(tmp_path):
  from odin_b200 import preprocessing as pp
  from odin_b200.ml import GMM
  cfg = FE_CONFIGS["cfg5"]
  sr, n_utt, M = cfg["sr"], 48, 512
  utts = synth.utterance_batch(n_utt, 0.3, 1.0, sr=sr, seed=555)
  jobs = [{"raw": u, "sr": sr, "name": "digit%03d" % i} for i, u in enumerate(utts)]
  pipe = pp.make_pipeline([
      pp.AudioReader(remove_dc=True), pp.PreEmphasis(0.97),
      pp.STFTExtractor(cfg["frame_length"], cfg["step_length"], n_fft=cfg["n_fft"], window="hamm", energy=True),
      pp.PowerSpecExtractor(2.0, output_name="spec"),
      pp.MelsSpecExtractor(cfg["n_mels"], fmin=cfg["fmin"], fmax=cfg["fmax"]),
      pp.MFCCsExtractor(cfg["n_ceps"], remove_first_coef=True, first_coef_energy=True),
      pp.DeltaExtractor("mfcc", order=(0, 1, 2)),
      pp.SADthreshold(input_name="mfcc_energy")])       # the SAD is stored, not applied (fsdd_ivec.py:97-106,197)
  feats, indices = pp.FeatureProcessor(jobs, extractor=pipe, batch_utts=16).run()
  # ---- front-end vs oracle: bit-exact counts / mask / indices, <= 1e-4 features
  pos = 0
  ref_feat, ref_sad = [], []
  for j, u in zip(jobs, utts):
    r = F.extract(u, sr, cfg["frame_length"], cfg["step_length"], cfg["n_fft"], n_mels=cfg["n_mels"],
                  fmin=cfg["fmin"], fmax=cfg["fmax"], n_ceps=cfg["n_ceps"], vad="threshold", vad_smooth=5)
    s, e = indices["mfcc"][j["name"]]
    assert s == pos and e - s == r["mfcc"].shape[0] == 1 + (len(u) - int(sr * cfg["frame_length"])) // int(sr * cfg["step_length"])
    assert np.array_equal(feats["sad"][s:e].astype(np.uint8), r["sad"].astype(np.uint8)), "VAD mask differs"
    assert relmax(feats["mfcc"][s:e], r["mfcc"]) < 1e-4
    ref_feat.append(r["mfcc"])
    ref_sad.append(r["sad"])
    pos = e
  X = np.ascontiguousarray(feats["mfcc"], dtype=np.float32)
  assert X.shape == (pos, 60)
  # ---- 512-mix UBM: parameters from a short oracle fit on the same features (float64), loaded
  # into the GPU model; per-utterance statistics through transform_to_disk
  rng = np.random.RandomState(7)
  mean = X[rng.choice(len(X), M, replace=False)].T.astype(np.float64)
  sigma = np.tile(X.var(0)[:, None], (1, M)).astype(np.float64)
  w = np.full((1, M), 1.0 / M)
  for _ in range(2):
    z, f, s, l, n = OG.expectation(X, mean, sigma, w, compute_dtype=np.float64)
    mean, sigma, w, rb = OG.maximization(z, f, s, (mean, sigma, w))
  g = GMM(nmix=M, nmix_start=M)
  g.initialize(X)
  g.mean, g.sigma, g.w = mean.astype(np.float32), sigma.astype(np.float32), w.astype(np.float32)
  # one more EM iteration on the GPU (3xFP16 tcgen05 path at M = 512) vs the oracle
  Z, Fs, S, L = g.expectation(X)
  z, f, s, l, n = OG.expectation(X, g.mean, g.sigma, g.w, compute_dtype=np.float64)
  assert max(relmax(Z, z), relmax(Fs, f), relmax(S, s)) < 1e-3
  pz, pf = str(tmp_path / "Z.npy"), str(tmp_path / "F.npy")
  names = g.transform_to_disk(X, indices["mfcc"], pathZ=pz, pathF=pf)
  assert names == [j["name"] for j in jobs]                          # sorted by start == job order
  Zu, Fu = np.load(pz), np.load(pf)
  assert Zu.shape == (n_utt, M) and Fu.shape == (n_utt, M * 60)
  nr, zr, fr = OG.utterance_stats(X, [(nm, indices["mfcc"][nm]) for nm in names], g.mean, g.sigma, g.w,
                                  compute_dtype=np.float64)
  assert nr == names
  assert relmax(Zu, zr) < 1e-3 and relmax(Fu, fr) < 1e-3
  # F-hat layout: index m*D + d (column-major flatten of [D, M], gmm_tmat.py:754-757)
  s0, e0 = indices["mfcc"][names[0]]
  z0, f0 = OG.transform(X[s0:e0], g.mean, g.sigma, g.w, compute_dtype=np.float64)
  assert relmax(Fu[0].reshape(M, 60), np.asarray(f0).reshape(M, 60)) < 1e-3
  # ---- i-vector extractor on those statistics (examples/fsdd_ivec.py:229-240: Tmatrix.fit on (Z, F), then
  # transform): 3 EM iterations of a tv = 16 model on the GPU vs the oracle ON THE SAME statistics; rows of T
  # and i-vector coordinates are compared up to the sign the orthogonalisation leaves open
  from odin_b200.ml import Tmatrix
  from oracle import tmatrix as OT
  t = Tmatrix(16, g, niter=3)
  t.fit((Zu, Fu))
  Tm, T_invS, T_invS_Tt, hist = OT.fit(Zu.astype(np.float64), Fu.astype(np.float64), 16, np.asarray(g.sigma), 3)
  assert np.allclose(t._llk_hist, hist, rtol=1e-6)
  a, sa = OT.sign_normalise(t.Tm)
  b, sb = OT.sign_normalise(Tm)
  assert relmax(a, b) < 1e-5, relmax(a, b)
  iv = t.transform((Zu, Fu))
  ref_iv = OT.ivector(Zu.astype(np.float64), Fu.astype(np.float64), T_invS, T_invS_Tt)
  assert iv.shape == (n_utt, 16) and relmax(iv * sa[None, :], ref_iv * sb[None, :]) < 1e-5


def test_fsdd_recipe_with_normalisation_tail():
  """The complete extractor list of examples/fsdd_ivec.py:80-106 (Converter / Rename / Delete /
  AcousticNorm(mvn + wmvn) / AsType('float16') included) through make_pipeline, against the oracle."""
  from odin_b200 import preprocessing as pp
  cfg = FE_CONFIGS["cfg5"]
  sr = cfg["sr"]
  utts = synth.utterance_batch(5, 0.3, 1.0, sr=sr, seed=777) + \
      synth.utterance_batch(5, 1.7, 2.4, sr=sr, seed=778)          # > 1.53 s: w = 301 slides at a 5 ms hop
  jobs = [{"raw": u, "sr": sr, "path": "/data/fsdd/%d_jackson_%d.wav" % (i % 10, i)} for i, u in enumerate(utts)]
  pipe = pp.make_pipeline([
      pp.AudioReader(remove_dc=True), pp.PreEmphasis(coeff=0.97),
      pp.Converter(converter=lambda x: os.path.basename(x).split('.')[0], input_name='path', output_name='name'),
      pp.STFTExtractor(frame_length=0.025, step_length=0.005, n_fft=512, window='hamm', energy=False),
      pp.PowerSpecExtractor(power=2.0, output_name='spec'),
      pp.MelsSpecExtractor(n_mels=24, fmin=64, fmax=4000, input_name=('spec', 'sr'), output_name='mspec'),
      pp.MFCCsExtractor(n_ceps=20, remove_first_coef=True, first_coef_energy=True, input_name='mspec',
                        output_name='mfcc'),
      pp.DeltaExtractor(input_name='mfcc', order=(0, 1, 2)),
      pp.RenameFeatures(input_name='mfcc_energy', output_name='energy'),
      pp.SADthreshold(energy_threshold=0.55, smooth_window=5, input_name='energy', output_name='sad'),
      pp.DeleteFeatures(input_name=('stft', 'spec', 'sad_threshold')),
      pp.AcousticNorm(mean_var_norm=True, windowed_mean_var_norm=True, input_name=('mspec', 'mfcc')),
      pp.AsType(dtype='float16')])
  outs = pipe.transform_batch(jobs)
  slid = 0
  for j, u, o in zip(jobs, utts, outs):
    r = F.extract(u, sr, 0.025, 0.005, 512, n_mels=24, fmin=64, fmax=4000, n_ceps=20, vad="threshold", vad_smooth=5)
    assert o["name"] == os.path.basename(j["path"]).split('.')[0]
    assert "stft" not in o and "spec" not in o and "sad_threshold" not in o
    assert np.array_equal(np.asarray(o["sad"]).astype(np.uint8), r["sad"].astype(np.uint8))
    for f in ("mspec", "mfcc"):
      want = F.acoustic_norm(r[f], mean_var_norm=True, windowed_mean_var_norm=True, win_length=301)
      assert o[f].dtype == np.float16 and o[f].shape == want.shape
      # float16 storage: 2^-11 relative on top of the 1e-4 feature tolerance
      assert np.max(np.abs(o[f].astype(np.float64) - want)) <= 1e-3 * max(1.0, float(np.max(np.abs(want))))
    slid += int(r["mfcc"].shape[0] >= 301)
  assert slid >= 3, "the test must exercise the sliding window"
