"""The recipe of the reference's examples/fsdd_ivec.py (feature extraction -> UBM -> Baum-Welch statistics ->
T-matrix -> i-vectors -> cosine / PLDA scoring) with the imports switched to odin_b200, on synthetic "digits" (there is no dataset or
network here): 8 kHz, 25 ms / 5 ms frames, 24 mel bands, 20 MFCC + c0 energy + deltas, SADthreshold, mean /
windowed-mean normalisation, float16 store (fsdd_ivec.py:80-106); 128-mixture UBM, tv_dim 64 (its defaults).

    python examples/fsdd_style_ivec.py [n_files]          # needs a B200
"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odin_b200 import preprocessing as pp   # reference: from odin import preprocessing as pp
from odin_b200 import ml                     # reference: from odin import ml
from odin_b200 import synth


def eer(scores, labels):
  """Equal error rate of the target / non-target trials of a [n_trials, n_classes] score matrix."""
  tgt = np.zeros(scores.shape, dtype=bool)
  tgt[np.arange(len(labels)), labels] = True
  s, t = scores.ravel(), tgt.ravel()
  order = np.argsort(-s)
  t = t[order]
  fa = np.cumsum(~t) / max(1, int((~t).sum()))        # false accepts above each threshold
  miss = 1.0 - np.cumsum(t) / max(1, int(t.sum()))    # misses below it
  i = int(np.argmin(np.abs(fa - miss)))
  return float(0.5 * (fa[i] + miss[i]))


def main(n_files=600, nmix=128, tv_dim=64):
  sr = 8000
  rng = np.random.RandomState(1)
  jobs = [{"raw": synth.speech_like(i, rng.uniform(0.3, 1.0), sr=sr, seed=1, speaker=i % 10), "sr": sr,
           "name": "spk%02d_%04d" % (i % 10, i)} for i in range(n_files)]   # ten synthetic "speakers"
  extractors = pp.make_pipeline(steps=[
      pp.AudioReader(sr=sr, remove_dc=True),
      pp.PreEmphasis(coeff=0.97),
      pp.DuplicateFeatures("sr", "sr_copy"),
      pp.STFTExtractor(frame_length=0.025, step_length=0.005, n_fft=512, window="hamm", energy=False),
      pp.PowerSpecExtractor(power=2.0, output_name="spec"),
      pp.MelsSpecExtractor(n_mels=24, fmin=64, fmax=4000, input_name=("spec", "sr"), output_name="mspec"),
      pp.MFCCsExtractor(n_ceps=20, remove_first_coef=True, first_coef_energy=True, input_name="mspec", output_name="mfcc"),
      pp.DeltaExtractor(input_name="mfcc", order=(0, 1, 2)),
      pp.RenameFeatures(input_name="mfcc_energy", output_name="energy"),
      pp.SADthreshold(energy_threshold=0.55, smooth_window=5, input_name="energy", output_name="sad"),
      pp.DeleteFeatures(input_name=("stft", "spec", "sad_threshold", "sr_copy")),
      pp.AcousticNorm(mean_var_norm=True, windowed_mean_var_norm=True, input_name=("mspec", "mfcc")),
      pp.AsType(dtype="float16"),
  ])
  t0 = time.perf_counter()
  feats, indices = pp.FeatureProcessor(jobs, extractor=extractors, batch_utts=256).run()
  X = np.ascontiguousarray(feats["mfcc"], dtype=np.float32)
  t1 = time.perf_counter()
  print("features: %d files, %d frames x %d dims in %.2f s" % (n_files, X.shape[0], X.shape[1], t1 - t0))
  with tempfile.TemporaryDirectory() as d:
    gmm = ml.GMM(nmix=nmix, nmix_start=1, niter=12, dtype="float32", allow_rollback=True, exit_on_error=True,
                 downsample=1, stochastic_downsample=True, device="gpu", ncpu=1, gpu_factor=3)
    gmm.fit((X, indices["mfcc"]))
    t2 = time.perf_counter()
    print("UBM: %d mixtures, final llk %.4f in %.2f s" % (nmix, gmm._llk_hist[nmix][-1], t2 - t1))
    zp, fp = os.path.join(d, "Z.npy"), os.path.join(d, "F.npy")
    names = gmm.transform_to_disk(X, indices["mfcc"], pathZ=zp, pathF=fp, dtype="float32")
    Z, F = np.load(zp), np.load(fp)
    t3 = time.perf_counter()
    print("statistics: Z %s, F %s in %.2f s" % (Z.shape, F.shape, t3 - t2))
    tmat = ml.Tmatrix(tv_dim=tv_dim, gmm=gmm, niter=16, dtype="float64")
    tmat.fit((Z, F))
    ivecs = tmat.transform_to_disk(Z, F, path=os.path.join(d, "ivec.npy"), dtype="float32")
    t4 = time.perf_counter()
    print("T-matrix: tv_dim %d, llk %.4f -> %.4f; i-vectors %s in %.2f s" %
          (tv_dim, tmat._llk_hist[0], tmat._llk_hist[-1], ivecs.shape, t4 - t3))
    # ten synthetic "speakers" (own pitch and formants): same-speaker i-vectors must be closer than different ones
    spk = np.array([int(n[3:5]) for n in names])
    iv = np.asarray(ivecs, dtype=np.float64)
    iv = iv / np.linalg.norm(iv, axis=1, keepdims=True)
    cos = iv.dot(iv.T)
    same = cos[spk[:, None] == spk[None, :]].mean()
    diff = cos[spk[:, None] != spk[None, :]].mean()
    print("mean cosine: same speaker %.3f, different speaker %.3f" % (same, diff))
    assert same > diff
    # scoring back-end (fsdd_ivec.py:270-330): alternate blocks of ten files (every speaker once) enrol / are trials
    blk = (np.arange(len(spk)) // 10) % 2
    tr, te = np.where(blk == 0)[0], np.where(blk == 1)[0]
    X_train, X_test = np.asarray(ivecs[tr], np.float64), np.asarray(ivecs[te], np.float64)
    for tag, scorer in (("cosine (centering + WCCN + LDA)", ml.Scorer(centering=True, wccn=True, lda=True, method="cosine")),
                        ("PLDA (n_phi %d, 12 iterations)" % (tv_dim // 2),
                         ml.PLDA(n_phi=tv_dim // 2, n_iter=12, centering=True, wccn=True, unit_length=True, random_state=1234))):
      t5 = time.perf_counter()
      scorer.fit(X_train, spk[tr])
      scores = scorer.predict_log_proba(X_test)
      acc = float(np.mean(np.argmax(scores, 1) == spk[te]))
      print("%s: accuracy %.3f, EER %.3f  (fit + score %.1f ms)" % (tag, acc, eer(scores, spk[te]), 1e3 * (time.perf_counter() - t5)))
      assert acc > 0.3   # chance is 0.1: the synthetic "speakers" differ in pitch and formants only, utterances are < 1 s
  print("total %.2f s" % (time.perf_counter() - t0))


if __name__ == "__main__":
  main(int(sys.argv[1]) if len(sys.argv) > 1 else 600)
