"""GPU parity of AcousticNorm / odin_fe_cmvn against the golden vectors of the
real reference (speech.py:1536-1610 through oracle/make_golden.py) and against the oracle on
ragged batches.  Tolerance: <= 1e-4 as max|a-b| / max|b| per matrix; NaN patterns (empty SAD
selection) must coincide."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relmax
from oracle import frontend as F

pytestmark = pytest.mark.gpu

TAGS = {
    "mvn": dict(),
    "mvn_novar": dict(var_norm=False),
    "wmvn": dict(windowed_mean_var_norm=True, win_length=51),
    "wonly": dict(mean_var_norm=False, windowed_mean_var_norm=True, win_length=31),
    "recipe": dict(windowed_mean_var_norm=True, win_length=301),
    "wmvn_sad": dict(windowed_mean_var_norm=True, win_length=51, sad_name="sad"),
}


def _same(y, r, tol=1e-4):
  assert y.shape == r.shape
  assert np.array_equal(np.isnan(y), np.isnan(r)), "NaN pattern differs"
  ok = np.isfinite(r)
  if ok.any():
    assert float(np.max(np.abs(y[ok] - r[ok]))) <= tol * max(float(np.max(np.abs(r[ok]))), 1e-30)


def test_cmvn_golden_batched():
  from odin_b200 import preprocessing as pp
  g = np.load(os.path.join(GOLDEN, "cmvn.npz"))
  n = int(g["n_case"])
  jobs = [{"mfcc": g["c%d_in_mfcc" % k], "mspec": g["c%d_in_mspec" % k], "sad": g["c%d_sad" % k], "name": "u%d" % k}
          for k in range(n)]
  for tag, kw in TAGS.items():
    e = pp.AcousticNorm(input_name=("mspec", "mfcc"), **kw)
    outs = e.transform_batch(jobs)                # one ragged launch per feature
    one = e.transform(jobs[2])                    # and a single utterance: identical
    for k, o in enumerate(outs):
      for f in ("mfcc", "mspec"):
        _same(o[f], g["c%d_%s_%s" % (k, tag, f)])
        assert o[f].dtype == np.float32
      assert o["name"] == "u%d" % k and np.array_equal(o["sad"], jobs[k]["sad"])   # other keys pass through
    assert np.array_equal(one["mfcc"], outs[2]["mfcc"], equal_nan=True)


def test_cmvn_long_ragged_vs_oracle():
  """recipe setting (w = 301) on utterances longer and shorter than the window, with and without SAD."""
  from odin_b200 import preprocessing as pp
  rng = np.random.RandomState(5)
  lens = [40, 300, 301, 302, 1500, 4000, 7]
  jobs = []
  for i, n in enumerate(lens):
    x = (rng.randn(n, 60) * rng.uniform(0.5, 20.0, size=60) + rng.uniform(-50, 50, size=60)).astype(np.float32)
    x += np.cumsum(rng.randn(n, 60) * 0.05, 0).astype(np.float32)          # slow drift: the window matters
    sad = rng.rand(n) > 0.35
    if i == 0:
      sad[:] = False                                                        # nothing selected -> NaN everywhere
    jobs.append({"mfcc": x, "sad": sad})
  for kw in (dict(mean_var_norm=True, windowed_mean_var_norm=True, win_length=301),
             dict(mean_var_norm=True, windowed_mean_var_norm=True, win_length=301, sad_name="sad"),
             dict(mean_var_norm=False, windowed_mean_var_norm=True, win_length=3, sad_name="sad"),
             dict(mean_var_norm=True, windowed_mean_var_norm=False, var_norm=True, sad_name="sad")):
    e = pp.AcousticNorm(input_name="mfcc", **kw)
    outs = e.transform_batch(jobs)
    okw = {k: v for k, v in kw.items() if k != "sad_name"}
    for j, o in zip(jobs, outs):
      r = F.acoustic_norm(j["mfcc"].astype(np.float64), sad=j["sad"] if "sad_name" in kw else None, **okw)
      _same(o["mfcc"], r)


def test_cmvn_argument_errors():
  from odin_b200 import preprocessing as pp
  with pytest.raises(ValueError):
    pp.AcousticNorm("mfcc", win_length=300)
  with pytest.raises(ValueError):
    pp.AcousticNorm("mfcc", win_length=1)
  e = pp.AcousticNorm("mfcc")
  sig = e.transform({"mspec": np.zeros((4, 3), np.float32)})     # missing input -> error signal (base.py:310-316)
  assert isinstance(sig, pp.ExtractorSignal)
