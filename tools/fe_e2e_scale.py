"""End-to-end front-end throughput (pinned host PCM in, features + VAD out) against corpus size and chunk count."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odin_b200 import synth, preprocessing as pp  # noqa: E402

pipe = pp.make_pipeline([pp.AudioReader(), pp.PreEmphasis(0.97), pp.STFTExtractor(0.025, 0.010, n_fft=1024, window="hamm"),
                         pp.PowerSpecExtractor(), pp.MelsSpecExtractor(80, fmin=64, fmax=8000),
                         pp.MFCCsExtractor(20, first_coef_energy=True), pp.DeltaExtractor("mfcc", order=(0, 1, 2)),
                         pp.SADgmm(input_name="stft_energy")])
fe = pipe.plan[0]
pool = synth.utterance_batch(24, 5.0, 60.0, sr=16000, seed=4000)
lens = np.array([len(u) for u in pool])
one = np.concatenate(pool)
CASES = ((12.5, (16, 32)), (50.0, (32, 64, 128, 256)), (100.0, (128, 256))) if len(sys.argv) < 2 else ((float(sys.argv[1]), tuple(int(x) for x in sys.argv[2:])),)
for hours, chunk_list in CASES:
  reps = int(round(hours * 3600 * 16000 / lens.sum()))
  off = np.zeros(reps * 24 + 1, np.int64)
  np.cumsum(np.tile(lens, reps), out=off[1:])
  pin = torch.empty(int(off[-1]), dtype=torch.int16, pin_memory=True)
  v = pin.numpy()
  for r in range(reps):
    v[r * len(one):(r + 1) * len(one)] = one
  for sd in (None, "float16"):
    for nch in chunk_list:
      out, ts = None, []
      for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fe.run_host_packed(pin, off, 16000, want=("feat", "sad"), n_chunks=nch, out=out, store_dtype=sd)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
      T = out["feat"].shape[0]
      dt = min(ts[1:])
      print("hours %5.1f store %-7s chunks %3d: %6.1f M frames/s (%.3f s)  H2D %.1f GB/s  D2H %.1f GB/s" % (
          hours, sd or "float32", nch, T / dt / 1e6, dt, pin.numel() * 2 / dt / 1e9,
          (out["feat"].numel() * out["feat"].element_size() + T) / dt / 1e9), flush=True)
      del out
  del pin
