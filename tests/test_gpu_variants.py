"""GPU parity of the feature-matrix extractors of SURVEY 8f-3 (Framing, CalculateEnergy, RASTAfilter with
shifted deltas, StackFeatures) against the reference's golden outputs (tests/golden/variants.npz) and the
oracle on ragged batches."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relmax
from oracle import frontend as F

pytestmark = pytest.mark.gpu


def test_framing_and_energy_golden():
  from odin_b200 import preprocessing as pp
  g = np.load(os.path.join(GOLDEN, "variants.npz"))
  for tag, padding in (("nopad", False), ("pad", True)):
    pipe = pp.make_pipeline([pp.AudioReader(remove_dc=True), pp.PreEmphasis(0.97),
                             pp.Framing(0.025, 0.010, window="hamm", padding=padding), pp.CalculateEnergy(log=True)])
    X = pipe.transform({"raw": g["fr_pcm"], "sr": 16000})
    assert X["frames"].shape == g["fr_%s_frames" % tag].shape and X["frames"].dtype == np.float32
    assert relmax(X["frames"], g["fr_%s_frames" % tag]) < 1e-6
    assert abs(X["scale"] - float(g["fr_%s_scale" % tag])) < 1e-15
    assert X["energy"].shape == g["fr_%s_energy" % tag].shape
    assert relmax(X["energy"], g["fr_%s_energy" % tag]) < 1e-6
  # the fused kernel's own energy output (odin_fe_frames with d_energy) agrees with CalculateEnergy
  lin = pp.CalculateEnergy(log=False).transform({"frames": g["fr_nopad_frames"]})["energy"]
  assert relmax(np.log(lin), g["fr_nopad_energy"]) < 1e-6


def test_rasta_sdc_stack_golden_and_ragged():
  from odin_b200 import preprocessing as pp
  g = np.load(os.path.join(GOLDEN, "variants.npz"))
  n = int(g["n_mat"])
  xs = [g["m%d_x" % i] for i in range(n)]
  for key, ex in (("rasta_sdc", pp.RASTAfilter(True, 1, "mfcc")), ("rasta", pp.RASTAfilter(True, 0, "mfcc")),
                  ("sdc2", pp.RASTAfilter(False, 2, "mfcc")), ("stack3", pp.StackFeatures(3, "mfcc"))):
    # one at a time and as one ragged batch (matrices of different widths go to different launches)
    single = [ex.transform({"mfcc": x})["mfcc"] for x in xs]
    batch = [o["mfcc"] for o in ex.transform_batch([{"mfcc": x} for x in xs])]
    for i in range(n):
      ref = g["m%d_%s" % (i, key)]
      assert single[i].shape == ref.shape and single[i].dtype == np.float32, (key, i)
      assert relmax(single[i], ref) < 1e-6, (key, i, relmax(single[i], ref))
      assert np.array_equal(single[i], batch[i]), (key, i)
  # same width, ragged lengths, against the oracle
  rng = np.random.RandomState(4)
  mats = [rng.randn(T, 12).astype(np.float32) for T in (1, 2, 4, 5, 9, 150, 33)]
  outs = pp.RASTAfilter(True, 1, "mfcc").transform_batch([{"mfcc": m} for m in mats])
  for m, o in zip(mats, outs):
    assert relmax(o["mfcc"], F.rasta_sdc(m, True, 1)) < 1e-6
  outs = pp.StackFeatures(2, "mfcc").transform_batch([{"mfcc": m} for m in mats])
  for m, o in zip(mats, outs):
    assert np.array_equal(o["mfcc"], F.stack_context(m, 2))


def test_standalone_sad_extractors():
  """SADgmm / SADthreshold on an energy feature that is not produced by the fused STFT (here: CalculateEnergy on
  explicit frames, and a c0-like feature): bit-exact masks vs the oracle, single and ragged batch."""
  from odin_b200 import preprocessing as pp
  from odin_b200 import synth
  utts = synth.utterance_batch(9, 0.4, 2.2, sr=16000, seed=31)
  pipe = pp.make_pipeline([pp.AudioReader(remove_dc=True), pp.PreEmphasis(0.97), pp.Framing(0.025, 0.010, window="hamm"),
                           pp.CalculateEnergy(log=True), pp.SADgmm(3, smooth_window=3, input_name="energy")])
  assert [type(x).__name__ for x in pipe.plan] == ["Framing", "CalculateEnergy", "SADgmm"]
  outs = pipe.transform_batch([{"raw": u, "sr": 16000} for u in utts])
  for u, o in zip(utts, outs):
    r = F.extract(u, 16000, vad="gmm", fmax=8000)
    assert relmax(o["energy"], r["stft_energy"]) < 1e-6
    sad, thr = F.sad_gmm(o["energy"], 3, 25, 3)            # the oracle on the energies the extractor saw
    assert np.array_equal(o["sad"], sad) and abs(o["sad_threshold"] - thr) < 1e-9
    one = pp.SADgmm(3, smooth_window=3, input_name="energy").transform({"energy": o["energy"]})
    assert np.array_equal(one["sad"], o["sad"])
  rng = np.random.RandomState(8)
  es = [np.cumsum(rng.randn(n)).astype(np.float32) for n in (7, 60, 333, 5, 1200)]
  th = pp.SADthreshold(input_name="e", output_name="v")
  outs = th.transform_batch([{"e": e} for e in es])
  for e, o in zip(es, outs):
    s2, t2 = F.sad_threshold(e)
    assert o["v"].dtype == bool and np.array_equal(o["v"], s2.astype(bool)) and abs(o["v_threshold"] - t2) < 1e-7


@pytest.mark.parametrize("seed,mx,mn,thr", [(5, 5, None, 0.6), (6, 3, 1.0, 0.5), (7, 4, None, 0.8)])
def test_vad_split_audio(seed, mx, mn, thr):
  """signal.vad_split_audio (signal.py:341-478) on the GPU kernels vs the oracle (itself checked against the
  reference in tests/test_oracle_vs_reference.py): identical segments, smoothed VAD curve, voiced and cut indicators."""
  from odin_b200 import synth
  from odin_b200.preprocessing import signal
  s = np.concatenate(synth.utterance_batch(3, 4.0, 6.0, sr=8000, seed=seed)).astype(np.float32)
  segs, vad, voices, cut = signal.vad_split_audio(s, 8000, mx, mn, 128, 3, thr, return_vad=True, return_voices=True,
                                                  return_cut=True)
  o_segs, o_vad, o_voices, o_cut = F.vad_split_audio(s, 8000, mx, mn, 128, 3, thr)
  assert np.array_equal(vad, o_vad) and np.array_equal(voices, o_voices) and np.array_equal(cut, o_cut)
  assert len(segs) == len(o_segs) and all(np.array_equal(a, b) for a, b in zip(segs, o_segs))
  assert signal.vad_split_audio(s[:8000], 8000, 5) [0] is not None and len(signal.vad_split_audio(s[:8000], 8000, 5)) == 1
