"""Seeded synthetic inputs for tests and benchmarks.

Not part of the reference: the reference ships no audio fixtures, so parity is
checked on these generated signals.  Pure numpy, deterministic per
``(seed, index)``; runs on the GPU box without any reference file.
"""
import numpy as np


def speech_like(index, duration_s, sr=16000, seed=1234, speaker=None):
  """int16 PCM: harmonics of f0 in U[90,250] Hz with 1/k roll-off, shaped by 3
  random formant resonances, gated by a random on/off envelope (segments
  0.1-0.6 s, about 55 % on) over -45 dB white noise plus a small DC offset;
  peak about 0.5 * 32767.  The gating makes the energy VAD, the top_db clip
  and the DC pre-pass all do real work.  `speaker` (int) ties pitch and formants to a
  speaker (with a few per cent of per-utterance jitter) while envelope, phases and
  noise stay per utterance -- for the scoring end of the recipe example; the default
  (None) draws everything per utterance, as all fixtures do."""
  rng = np.random.RandomState(seed + int(index))
  n = int(round(duration_s * sr))
  t = np.arange(n, dtype=np.float64) / sr
  f0 = rng.uniform(90.0, 250.0)
  vib = 1.0 + 0.02 * np.sin(2 * np.pi * rng.uniform(3.0, 7.0) * t)
  phase = 2 * np.pi * np.cumsum(f0 * vib) / sr
  formants = rng.uniform([300.0, 900.0, 2200.0], [900.0, 2200.0, min(3600.0, 0.45 * sr)])
  bw = rng.uniform(60.0, 200.0, size=3)
  if speaker is not None:   # (after the per-utterance draws, so the default stream is untouched)
    srng = np.random.RandomState(100003 + int(speaker))
    jit = np.random.RandomState(seed + 31 * int(index) + 7)
    f0 = srng.uniform(90.0, 250.0) * (1.0 + 0.03 * jit.randn())
    formants = srng.uniform([300.0, 900.0, 2200.0], [900.0, 2200.0, min(3600.0, 0.45 * sr)]) * (1.0 + 0.02 * jit.randn(3))
    bw = srng.uniform(60.0, 200.0, size=3)
    phase = 2 * np.pi * np.cumsum(f0 * vib) / sr
  nharm = int(min(0.45 * sr, 5000.0) // f0)
  y = np.zeros(n)
  for k in range(1, nharm + 1):
    fk = k * f0
    gain = (1.0 / k) * (0.05 + np.sum(1.0 / (1.0 + ((fk - formants) / bw)**2)))
    y += gain * np.sin(k * phase + rng.uniform(0, 2 * np.pi))
  # on/off envelope with 10 ms raised-cosine ramps
  env = np.zeros(n)
  pos, on = 0, bool(rng.rand() < 0.55)
  while pos < n:
    seg = int(rng.uniform(0.1, 0.6) * sr)
    if on:
      env[pos:pos + seg] = rng.uniform(0.3, 1.0)
    pos += seg
    on = (rng.rand() < 0.55)
  r = max(1, int(0.010 * sr))
  ramp = 0.5 - 0.5 * np.cos(np.pi * (np.arange(2 * r) + 0.5) / (2 * r))
  kern = np.diff(np.concatenate([[0.0], ramp]))
  env = np.convolve(env, kern, mode="same")
  y = y / (np.max(np.abs(y)) + 1e-12) * env
  y = y + 10**(-45 / 20.0) * rng.randn(n)
  y = y / np.max(np.abs(y)) * 0.5 * 32767.0 + rng.uniform(-20.0, 20.0)
  return np.round(y).astype(np.int16)


def utterance_batch(n_utt, min_s, max_s, sr=16000, seed=1234, first_index=0):
  """List of int16 utterances with durations U[min_s, max_s]."""
  out = []
  for i in range(n_utt):
    rng = np.random.RandomState(seed + 7919 * (first_index + i))
    dur = min_s if max_s <= min_s else rng.uniform(min_s, max_s)
    out.append(speech_like(first_index + i, dur, sr=sr, seed=seed))
  return out


def pack_utterances(utts):
  """Concatenate ragged utterances -> (pcm [sum n], sample_offsets int64 [n_utt+1])."""
  lens = np.array([len(u) for u in utts], dtype=np.int64)
  off = np.zeros(len(utts) + 1, dtype=np.int64)
  np.cumsum(lens, out=off[1:])
  return np.concatenate(utts), off


def gmm_features(n_frames, dim=60, n_true=32, seed=1234, dtype=np.float32):
  """Features drawn from a seeded random diagonal mixture (means N(0,3^2),
  variances U[0.5,1.5]) so posteriors are neither one-hot nor uniform."""
  rng = np.random.RandomState(seed)
  mu = rng.randn(n_true, dim) * 3.0
  var = rng.uniform(0.5, 1.5, size=(n_true, dim))
  comp = rng.randint(0, n_true, size=n_frames)
  X = mu[comp] + np.sqrt(var[comp]) * rng.randn(n_frames, dim)
  return np.ascontiguousarray(X.astype(dtype))


def gmm_params(dim=60, nmix=64, seed=4321, dtype=np.float32):
  """Seeded UBM parameters in the reference layout: mean [D,M], sigma
  (= variance) [D,M], w [1,M]."""
  rng = np.random.RandomState(seed)
  mean = (rng.randn(dim, nmix) * 3.0).astype(dtype)
  sigma = rng.uniform(0.5, 1.5, size=(dim, nmix)).astype(dtype)
  w = rng.uniform(0.5, 1.5, size=(1, nmix))
  w = (w / w.sum()).astype(dtype)
  return mean, sigma, w
