"""Config 5 (BASELINE.json): FSDD-style end-to-end pipeline on the GPU path --
8 kHz synthetic digits -> MFCC (+c0 energy, deltas, SADthreshold, recipe of
examples/fsdd_ivec.py:80-106 minus AcousticNorm / AsType) -> FeatureProcessor store with
name -> (start, end) indices -> 512-mix UBM statistics -> per-utterance Z [n, 512] and
F-hat [n, 512*60] (the i-vector input of gmm_tmat.py:769-913).
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
