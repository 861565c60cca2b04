"""GPU parity: the fused speech front-end (through the C-ABI, via
odin_b200.preprocessing) against the golden vectors produced by the real
reference and against the oracle on fresh seeded audio.

Tolerances (BASELINE.json north_star): frame counts, VAD masks and frame
indexing bit-exact; log-mel / MFCC / deltas <= 1e-4 as max|a-b| / max|b| per
matrix."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relmax
from odin_b200 import synth
from oracle import frontend as F
from oracle.make_golden import FE_CONFIGS

pytestmark = pytest.mark.gpu

TOL_FEAT = 1e-4


def _pipeline(cfg, vad="gmm", apply_sad=None, delta=True):
  from odin_b200 import preprocessing as pp
  steps = [pp.AudioReader(remove_dc=True), pp.PreEmphasis(0.97),
           pp.STFTExtractor(cfg["frame_length"], cfg["step_length"], n_fft=cfg["n_fft"], window="hamm",
                            energy=True),
           pp.PowerSpecExtractor(2.0, output_name="spec"),
           pp.MelsSpecExtractor(cfg["n_mels"], fmin=cfg["fmin"], fmax=cfg["fmax"]),
           pp.MFCCsExtractor(cfg["n_ceps"], remove_first_coef=True, first_coef_energy=True)]
  if delta:
    steps.append(pp.DeltaExtractor("mfcc", order=(0, 1, 2)))
  if vad == "gmm":
    steps.append(pp.SADgmm(3, smooth_window=3, input_name="stft_energy"))
  elif vad == "threshold":
    steps.append(pp.SADthreshold(input_name="mfcc_energy"))
  if apply_sad:
    steps.append(pp.ApplyingSAD(apply_sad, keep_unvoiced=False))
  return pp.make_pipeline(steps)


@pytest.mark.parametrize("name", sorted(FE_CONFIGS))
@pytest.mark.parametrize("vad", ["gmm", "threshold"])
def test_golden_configs(name, vad):
  cfg = FE_CONFIGS[name]
  g = np.load(os.path.join(GOLDEN, "fe_%s.npz" % name))
  n = int(g["n_utt"])
  pipe = _pipeline(cfg, vad)
  jobs = [{"raw": g["u%d_pcm" % i], "sr": cfg["sr"], "name": "u%d" % i} for i in range(n)]
  outs = pipe.transform_batch(jobs)              # one ragged batch
  singles = [pipe.transform(j) for j in jobs]    # and one by one: identical
  for i, (o, s) in enumerate(zip(outs, singles)):
    T = g["u%d_mfcc" % i].shape[0]
    assert o["mfcc"].shape == (T, 3 * cfg["n_ceps"]) and o["mspec"].shape == (T, cfg["n_mels"])  # frame counts
    assert o["stft_energy"].shape == (T, 1) and o["sad"].shape == (T,)
    assert relmax(o["mspec"], g["u%d_mspec" % i]) < TOL_FEAT
    assert relmax(o["mfcc"][:, :20], g["u%d_mfcc" % i][:, :20]) < TOL_FEAT
    assert relmax(o["mfcc"][:, 20:40], g["u%d_mfcc" % i][:, 20:40]) < TOL_FEAT
    assert relmax(o["mfcc"][:, 40:], g["u%d_mfcc" % i][:, 40:]) < TOL_FEAT
    assert relmax(o["mfcc_energy"], g["u%d_c0" % i]) < TOL_FEAT
    assert np.max(np.abs(o["stft_energy"] - g["u%d_energy" % i])) < 2e-6
    key = "sad_gmm" if vad == "gmm" else "sad_thr"
    assert np.array_equal(o["sad"].astype(np.uint8), g["u%d_%s" % (i, key)]), "VAD mask differs"
    assert abs(o["sad_threshold"] - float(g["u%d_%s_threshold" % (i, key)])) < 1e-5
    for k in ("mspec", "mfcc", "sad", "stft_energy"):
      assert np.array_equal(o[k], s[k]), "batched != single for %s" % k
    assert o["name"] == "u%d" % i and o["sr"] == cfg["sr"]


def test_appendix_b_tones():
  """SURVEY.md Appendix B input: pure tones, the top_db clip is active."""
  g = np.load(os.path.join(GOLDEN, "fe_appendix_b.npz"))
  cfg = FE_CONFIGS["cfg1"]
  o = _pipeline(cfg, "gmm").transform({"raw": g["pcm"], "sr": 16000})
  assert o["mfcc"].shape == (28, 60)
  assert abs(o["mspec"].max() - 40.946535784) < 1e-3 and abs(o["mspec"].min() + 39.053464216) < 1e-3
  assert relmax(o["mspec"], g["mspec"]) < 5e-4        # 80 dB in-frame dynamic range in float32
  assert relmax(o["mfcc"][:, :20], g["mfcc"][:, :20]) < 5e-4
  assert "".join(map(str, o["sad"])) == "0000000000000001111111111111"
  assert abs(o["sad_threshold"] - 1.02282264) < 1e-5
  o2 = _pipeline(cfg, "threshold").transform({"raw": g["pcm"], "sr": 16000})
  assert "".join(map(str, o2["sad"].astype(int))) == "0000000000111111110000000000"
  assert abs(o2["sad_threshold"] - 0.73292095) < 1e-5


@pytest.mark.parametrize("name,n_utt", [("cfg1", 40), ("cfg3", 6), ("cfg5", 30)])
def test_fresh_audio_vs_oracle(name, n_utt):
  """config-1/3/5 shaped ragged batches against the oracle (which is pinned to the reference)."""
  cfg = FE_CONFIGS[name]
  sr = cfg["sr"]
  durs = {"cfg1": (3.0, 3.0), "cfg3": (5.0, 12.0), "cfg5": (0.3, 1.0)}[name]
  utts = synth.utterance_batch(n_utt, durs[0], durs[1], sr=sr, seed=4242)
  for vad in ("gmm", "threshold"):
    outs = _pipeline(cfg, vad).transform_batch([{"raw": u, "sr": sr} for u in utts])
    flips = 0
    total = 0
    for u, o in zip(utts, outs):
      r = F.extract(u, sr, cfg["frame_length"], cfg["step_length"], cfg["n_fft"], n_mels=cfg["n_mels"],
                    fmin=cfg["fmin"], fmax=cfg["fmax"], n_ceps=cfg["n_ceps"], vad=vad,
                    vad_smooth=3 if vad == "gmm" else 5)
      assert o["mfcc"].shape == r["mfcc"].shape
      assert relmax(o["mspec"], r["mspec"]) < TOL_FEAT and relmax(o["mfcc"], r["mfcc"]) < TOL_FEAT
      assert relmax(o["mfcc"][:, 40:], r["mfcc"][:, 40:]) < TOL_FEAT
      flips += int(np.sum(o["sad"].astype(np.uint8) != r["sad"].astype(np.uint8)))
      total += len(r["sad"])
    assert flips == 0, "%d of %d VAD decisions differ" % (flips, total)


def test_config3_full_length_utterances_vs_oracle():
  """Config 3 at the utterance lengths the benchmark runs (U[5, 60] s): a 60 s and a 37.3 s utterance (6 000 and 3 729
  frames: many tiles per utterance, utterance-global top_db and SADgmm over thousands of frames), n_fft 1024, 80 mel,
  both SADs, int16 PCM, against the oracle; and the WAV-file route (soundfile-style float32 in [-1, 1))."""
  import wave
  import tempfile
  cfg = FE_CONFIGS["cfg3"]
  sr = cfg["sr"]
  utts = [synth.speech_like(700, 60.0, sr=sr, seed=99), synth.speech_like(701, 37.3, sr=sr, seed=99)]
  for vad in ("gmm", "threshold"):
    outs = _pipeline(cfg, vad).transform_batch([{"raw": u, "sr": sr} for u in utts])
    for u, o in zip(utts, outs):
      r = F.extract(u, sr, cfg["frame_length"], cfg["step_length"], cfg["n_fft"], n_mels=cfg["n_mels"], fmin=cfg["fmin"],
                    fmax=cfg["fmax"], n_ceps=cfg["n_ceps"], vad=vad, vad_smooth=3 if vad == "gmm" else 5)
      assert o["mfcc"].shape == r["mfcc"].shape == (1 + (len(u) - 400) // 160, 60)
      assert relmax(o["mspec"], r["mspec"]) < TOL_FEAT and relmax(o["mfcc"], r["mfcc"]) < TOL_FEAT
      assert relmax(o["mfcc"][:, 20:40], r["mfcc"][:, 20:40]) < TOL_FEAT and relmax(o["mfcc"][:, 40:], r["mfcc"][:, 40:]) < TOL_FEAT
      assert np.array_equal(o["sad"].astype(np.uint8), r["sad"].astype(np.uint8)), "VAD mask differs (%s)" % vad
  # the same audio as a 16-bit WAV file: AudioReader normalises like soundfile (int16 / 32768 as float32)
  with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "u.wav")
    with wave.open(path, "wb") as f:
      f.setnchannels(1); f.setsampwidth(2); f.setframerate(sr)
      f.writeframes(utts[1][:sr * 8].astype("<i2").tobytes())
    o = _pipeline(cfg, "gmm").transform(path)
    xf = (utts[1][:sr * 8].astype(np.float64) / 32768.0).astype(np.float32)
    r = F.extract(xf, sr, cfg["frame_length"], cfg["step_length"], cfg["n_fft"], n_mels=cfg["n_mels"], fmin=cfg["fmin"],
                  fmax=cfg["fmax"], n_ceps=cfg["n_ceps"], vad="gmm", vad_smooth=3)
    assert o["path"] == os.path.abspath(path) and o["sr"] == sr
    assert relmax(o["mspec"], r["mspec"]) < TOL_FEAT and relmax(o["mfcc"], r["mfcc"]) < TOL_FEAT
    assert np.array_equal(o["sad"].astype(np.uint8), r["sad"].astype(np.uint8))


def test_ragged_edge_cases():
  cfg = FE_CONFIGS["cfg1"]
  L, hop = 400, 160
  rng = np.random.RandomState(3)
  lens = [L, L + 1, L + hop - 1, L + hop, L + 31 * hop, L + 32 * hop, L + 33 * hop, L + 127 * hop,
          L + 128 * hop, 7 * 16000 + 13]
  utts = [(rng.randn(n) * 3000).astype(np.int16) for n in lens]
  pipe = _pipeline(cfg, "gmm")
  outs = pipe.transform_batch([{"raw": u, "sr": 16000} for u in utts])
  for u, o in zip(utts, outs):
    T = 1 + (len(u) - L) // hop
    assert o["mfcc"].shape == (T, 60)
    r = F.extract(u, 16000, vad=None, fmax=8000)
    assert relmax(o["mspec"], r["mspec"]) < TOL_FEAT and relmax(o["mfcc"], r["mfcc"]) < TOL_FEAT
  # float32 input in [-1, 1] (soundfile-style) goes through the same kernels
  uf = (synth.speech_like(5, 1.0) / 32768.0).astype(np.float32)
  o = pipe.transform({"raw": uf, "sr": 16000})
  r = F.extract(uf, 16000, vad=None, fmax=8000)
  assert relmax(o["mspec"], r["mspec"]) < TOL_FEAT and relmax(o["mfcc"], r["mfcc"]) < TOL_FEAT
  # too-short utterance -> signal for that job only; others unaffected
  from odin_b200 import preprocessing as pp
  outs = pipe.transform_batch([{"raw": utts[3], "sr": 16000}, {"raw": np.zeros(100, np.int16), "sr": 16000},
                               {"raw": np.zeros(16000, np.int16), "sr": 16000}])
  assert isinstance(outs[1], pp.ExtractorSignal) and outs[0]["mfcc"].shape == (2, 60)
  # digital silence: energy falls back to float32 eps (signal.py:1436).  The reference's
  # float32 mean of identical values is off by one ulp, so the standardised energy is a
  # constant -1 and SADgmm marks EVERY frame as speech (threshold about -1.002): same here.
  r = F.extract(np.zeros(16000, np.int16), 16000, vad="gmm", fmax=8000)
  assert np.allclose(outs[2]["stft_energy"], np.log(np.finfo(np.float32).eps))
  assert np.array_equal(outs[2]["sad"], r["sad"]) and r["sad"].sum() == 98
  assert abs(outs[2]["sad_threshold"] - r["sad_threshold"]) < 1e-9
  assert np.all(np.isfinite(outs[2]["mfcc"])) and relmax(outs[2]["mspec"], r["mspec"]) < TOL_FEAT


def test_applying_sad_compaction_and_processor():
  """speech.py:1732-1756 + processor.py:640-651: compacted rows and name->(start,end) indices."""
  from odin_b200 import preprocessing as pp
  cfg = FE_CONFIGS["cfg5"]
  sr = cfg["sr"]
  utts = synth.utterance_batch(12, 0.3, 1.0, sr=sr, seed=99)
  jobs = [{"raw": u, "sr": sr, "name": "d%02d" % i} for i, u in enumerate(utts)]
  jobs.insert(4, {"raw": np.zeros(4000, np.int16), "sr": sr, "name": "silent"})  # dropped by ApplyingSAD
  full = _pipeline(cfg, "threshold").transform_batch(jobs)
  pipe = _pipeline(cfg, "threshold", apply_sad=("mfcc",))
  outs = pipe.transform_batch(jobs)
  assert isinstance(outs[4], pp.ExtractorSignal)
  for f, o in zip(full, outs):
    if isinstance(o, pp.ExtractorSignal):
      continue
    assert np.array_equal(o["mfcc"], f["mfcc"][f["sad"].astype(bool)])   # index-exact compaction
    assert o["mspec"].shape[0] == f["mspec"].shape[0]                    # only the named feature is cut
  proc = pp.FeatureProcessor(jobs, extractor=pipe, batch_utts=5)
  feats, indices = proc.run()
  assert "silent" not in indices["mfcc"] and len(indices["mfcc"]) == 12
  pos = 0
  for j, o in zip(jobs, outs):
    if isinstance(o, pp.ExtractorSignal):
      continue
    s, e = indices["mfcc"][j["name"]]
    assert s == pos and e - s == o["mfcc"].shape[0]                      # job order
    assert np.array_equal(feats["mfcc"][s:e], o["mfcc"])
    pos = e
  # the streaming on-disk store holds the same rows (appended batch by batch, never the whole corpus in memory)
  import tempfile
  with tempfile.TemporaryDirectory() as d:
    disk, ind2 = pp.FeatureProcessor(jobs, path=d, extractor=pipe, batch_utts=5).run()
    assert ind2 == indices and sorted(disk) == sorted(feats)
    for k in feats:
      assert isinstance(disk[k], np.memmap) and disk[k].dtype == feats[k].dtype
      assert np.array_equal(np.load(os.path.join(d, k + ".npy")), feats[k])
    rows = [l.strip().split(",") for l in open(os.path.join(d, "indices_mfcc.csv"))]
    assert [(r[0], (int(r[1]), int(r[2]))) for r in rows] == list(indices["mfcc"].items())


def test_device_tables_match_oracle():
  import ctypes as C
  from odin_b200 import _lib
  pipe = _pipeline(FE_CONFIGS["cfg3"], None)
  h, cfg = pipe.plan[0]._handle(16000)
  lib = _lib.load()

  def table(which, n):
    buf = np.zeros(n, dtype=np.float64)
    got = lib.odin_fe_get_table(h, which, buf.ctypes.data_as(C.POINTER(C.c_double)), n)
    assert got == n
    return buf

  assert np.allclose(table(0, 400), F.window_table("hamm", 400), rtol=0, atol=1e-15)
  assert np.allclose(table(1, 80 * 513).reshape(80, 513), F.mel_filterbank(16000, 1024, 80, 64, 8000), rtol=0,
                     atol=1e-15)
  assert np.allclose(table(2, 21 * 80).reshape(21, 80), F.dct_basis(21, 80), rtol=0, atol=1e-14)


def test_round_trip_properties_large():
  """size-independent checks on a config-3 sized shard (about 30 min of audio):
  frame indexing identities, batching invariance, energy/VAD consistency."""
  import torch
  sr = 16000
  cfg = FE_CONFIGS["cfg3"]
  pool = synth.utterance_batch(8, 5.0, 20.0, sr=sr, seed=777)
  utts = [pool[i % 8] for i in range(160)]
  pcm, off = synth.pack_utterances(utts)
  fe = _pipeline(cfg, "gmm").plan[0]
  out = fe.run_packed(torch.from_numpy(pcm).cuda(), off, sr)
  fo = out["frame_offsets"]
  assert list(np.diff(fo)) == [1 + (len(u) - 400) // 160 for u in utts]
  feat = out["feat"].cpu().numpy()
  sad = out["sad"].cpu().numpy()
  assert np.all(np.isfinite(feat))
  for i in range(8, 160):   # the same utterance gives the same rows wherever it sits in the batch
    j = i % 8
    assert np.array_equal(feat[fo[i]:fo[i + 1]], feat[fo[j]:fo[j + 1]])
    assert np.array_equal(sad[fo[i]:fo[i + 1]], sad[fo[j]:fo[j + 1]])
  mspec = out["mspec"].cpu().numpy()
  for i in range(8):        # utterance-global top_db clip: min >= max - 80 exactly
    m = mspec[fo[i]:fo[i + 1]]
    assert m.min() >= m.max() - 80.0 - 1e-4


def test_spectra_extractor_and_padding():
  """SpectraExtractor and STFTExtractor(padding=True) against the reference's golden outputs."""
  from odin_b200 import preprocessing as pp
  from oracle.make_golden import SPECTRA_CASES
  g = np.load(os.path.join(GOLDEN, "spectra.npz"))
  for i, (sr, _, kw, front) in enumerate(SPECTRA_CASES):
    pcm = g["c%d_pcm" % i]
    steps = ([pp.AudioReader(remove_dc=True), pp.PreEmphasis(0.97)] if front else []) + [pp.SpectraExtractor(**kw)]
    pipe = pp.make_pipeline(steps)
    X = pipe.transform({"raw": pcm if front else pcm.astype(np.float32), "sr": sr})
    for k in ("spec", "energy", "mspec", "mfcc"):
      key = "c%d_%s" % (i, k)
      if key in g.files:
        assert X[k].dtype == np.float32 and X[k].shape == g[key].shape, (i, k, X[k].shape)
        assert relmax(X[k], g[key]) < TOL_FEAT, (i, k, relmax(X[k], g[key]))
      else:
        assert k not in X
    assert "stft_energy" not in X and X["sr"] == sr
  # chained extractors with centred frames
  cfg = FE_CONFIGS["cfg1"]
  steps = [pp.AudioReader(remove_dc=True), pp.PreEmphasis(0.97),
           pp.STFTExtractor(cfg["frame_length"], cfg["step_length"], n_fft=cfg["n_fft"], window="hamm", energy=True,
                            padding=True),
           pp.PowerSpecExtractor(2.0, output_name="spec"),
           pp.MelsSpecExtractor(cfg["n_mels"], fmin=cfg["fmin"], fmax=cfg["fmax"]),
           pp.MFCCsExtractor(cfg["n_ceps"], remove_first_coef=True, first_coef_energy=True),
           pp.DeltaExtractor("mfcc", order=(0, 1, 2)), pp.SADgmm(3, smooth_window=3, input_name="stft_energy")]
  X = pp.make_pipeline(steps).transform({"raw": g["pad_pcm"], "sr": 16000})
  assert X["mfcc"].shape == g["pad_mfcc"].shape
  assert relmax(X["stft_energy"], g["pad_energy"]) < 1e-6
  assert relmax(X["mspec"], g["pad_mspec"]) < TOL_FEAT and relmax(X["mfcc"], g["pad_mfcc"]) < TOL_FEAT
  assert np.array_equal(X["sad"], g["pad_sad_gmm"])
  # batch of ragged utterances, spectrum vs the oracle (dB clip is per utterance)
  utts = synth.utterance_batch(5, 0.2, 1.1, sr=16000, seed=77)
  ex = pp.SpectraExtractor(0.025, 0.010, n_fft=512, n_mels=40, n_ceps=13, padding=True)
  outs = ex.transform_batch([{"raw": u.astype(np.float32), "sr": 16000} for u in utts])
  for u, o in zip(utts, outs):
    r = F.spectra(u.astype(np.float32), 16000, 0.025, 0.010, n_fft=512, n_mels=40, n_ceps=13, padding=True)
    for k in ("spec", "mspec", "mfcc"):
      assert o[k].shape == r[k].shape and relmax(o[k], r[k]) < TOL_FEAT, k


def test_run_host_packed_matches_run_packed():
  """The pipelined host-buffer call (chunks of whole utterances over three streams) returns exactly what one
  resident call does: per-utterance results do not depend on how the batch is cut."""
  import torch
  from odin_b200 import preprocessing as pp
  cfg = FE_CONFIGS["cfg1"]
  fe = _pipeline(cfg, vad="gmm").plan[0]
  utts = synth.utterance_batch(23, 0.3, 2.5, sr=16000, seed=5)
  pcm, off = synth.pack_utterances(utts)
  ref = fe.run_packed(torch.from_numpy(pcm).cuda(), off, 16000)
  pinned = torch.from_numpy(pcm).pin_memory()
  for n_chunks in (1, 4, 50):
    out = fe.run_host_packed(pinned, off, 16000, want=("feat", "sad", "mspec", "energy", "c0"), n_chunks=n_chunks)
    assert np.array_equal(out["frame_offsets"], ref["frame_offsets"])
    for k in ("feat", "sad", "mspec", "energy", "c0"):
      assert torch.equal(out[k], ref[k].cpu()), (n_chunks, k)


@pytest.mark.parametrize("n_fft,frame_length,force_stockham", [(2048, 0.064, False), (512, 0.025, True), (256, 0.016, False)])
def test_other_fft_sizes_vs_oracle(n_fft, frame_length, force_stockham, monkeypatch):
  """n_fft 2048 runs on the three-pass Stockham kernel (also forced at 512 for an A/B of the two FFT kernels),
  256 on the four-step kernel with four frame pairs per warp."""
  from odin_b200 import preprocessing as pp
  if force_stockham:
    monkeypatch.setenv("ODIN_FE_STOCKHAM", "1")
  pipe = pp.make_pipeline([pp.AudioReader(remove_dc=True), pp.PreEmphasis(0.97),
                           pp.STFTExtractor(frame_length, 0.010, n_fft=n_fft, window="hamm", energy=True),
                           pp.PowerSpecExtractor(2.0), pp.MelsSpecExtractor(40, fmin=64, fmax=8000),
                           pp.MFCCsExtractor(20, first_coef_energy=True), pp.DeltaExtractor("mfcc", order=(0, 1, 2)),
                           pp.SADgmm(3, smooth_window=3, input_name="stft_energy")])
  utts = synth.utterance_batch(7, 0.4, 1.6, sr=16000, seed=91)
  outs = pipe.transform_batch([{"raw": u, "sr": 16000} for u in utts])
  for u, o in zip(utts, outs):
    r = F.extract(u, 16000, frame_length, 0.010, n_fft, n_mels=40, fmin=64, fmax=8000, vad="gmm")
    assert o["mfcc"].shape == r["mfcc"].shape
    assert relmax(o["mspec"], r["mspec"]) < TOL_FEAT and relmax(o["mfcc"], r["mfcc"]) < TOL_FEAT
    assert relmax(o["stft_energy"], r["stft_energy"]) < 1e-6
    assert np.array_equal(o["sad"], r["sad"])


def test_growing_batches_on_one_handle():
  """The per-call staging of a handle grows with the batch: a second, larger batch (more utterances, more frames)
  on the same pipeline must give the same per-utterance results as the first."""
  cfg = FE_CONFIGS["cfg1"]
  pipe = _pipeline(cfg, vad="gmm")
  utts = synth.utterance_batch(40, 0.3, 0.9, sr=16000, seed=321)
  small = pipe.transform_batch([{"raw": u, "sr": 16000} for u in utts[:3]])
  big = pipe.transform_batch([{"raw": u, "sr": 16000} for u in utts])
  again = pipe.transform_batch([{"raw": u, "sr": 16000} for u in utts[:3]])
  for a, b, c in zip(small, big[:3], again):
    for k in ("mfcc", "mspec", "sad", "stft_energy"):
      assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k]), k


def test_sadgmm_many_short_utterances_warp_kernel():
  """More than 4 x SM-count short utterances switch SADgmm to the warp-per-utterance kernel (and a few long ones in the
  same batch stay on the cluster kernel): masks must still equal the oracle's bit for bit, and equal what the cluster
  kernel gives for the same utterances in a small batch."""
  cfg = FE_CONFIGS["cfg1"]
  pipe = _pipeline(cfg, vad="gmm")
  short = synth.utterance_batch(30, 0.3, 2.0, sr=16000, seed=777)
  long_ = synth.utterance_batch(2, 6.0, 8.0, sr=16000, seed=778)
  utts = [short[i % 30] for i in range(700)] + long_
  outs = pipe.transform_batch([{"raw": u, "sr": 16000} for u in utts])
  small = pipe.transform_batch([{"raw": u, "sr": 16000} for u in short + long_])
  for i in range(30):
    r = F.extract(short[i], 16000, vad="gmm", fmax=8000)
    assert np.array_equal(outs[i]["sad"], r["sad"]), i
    assert abs(outs[i]["sad_threshold"] - r["sad_threshold"]) < 1e-9
    assert np.array_equal(outs[i]["sad"], small[i]["sad"]) and np.array_equal(outs[i + 30]["sad"], small[i]["sad"])
  for j in range(2):
    assert np.array_equal(outs[700 + j]["sad"], small[30 + j]["sad"])
    assert np.array_equal(outs[700 + j]["sad"], F.extract(long_[j], 16000, vad="gmm", fmax=8000)["sad"])
