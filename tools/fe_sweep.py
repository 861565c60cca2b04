"""A/B sweep of the front-end launch knobs on the config-3 shard of bench.py (GPU only).
Prints per-kernel milliseconds (dc, frame, post, vad) for each setting of the environment knobs."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odin_b200 import _lib, synth  # noqa: E402
from odin_b200 import preprocessing as pp  # noqa: E402


def main():
  sr = 16000
  pipe = pp.make_pipeline([
      pp.AudioReader(), pp.PreEmphasis(0.97), pp.STFTExtractor(0.025, 0.010, n_fft=1024, window="hamm"),
      pp.PowerSpecExtractor(), pp.MelsSpecExtractor(80, fmin=64, fmax=8000),
      pp.MFCCsExtractor(20, first_coef_energy=True), pp.DeltaExtractor("mfcc", order=(0, 1, 2)),
      pp.SADgmm(input_name="stft_energy")])
  fe = pipe.plan[0]
  pool = synth.utterance_batch(24, 5.0, 60.0, sr=sr, seed=4000)
  utts = [pool[i % len(pool)] for i in range(9 * len(pool))]
  pcm_h, off = synth.pack_utterances(utts)
  pcm = torch.from_numpy(pcm_h).cuda()
  lib = _lib.load()
  h, _ = fe._handle(sr)
  ref = None
  if "--once" in sys.argv:   # for ncu captures: two plain runs
    for _ in range(2):
      fe.run_packed(pcm, off, sr)
    torch.cuda.synchronize()
    return
  settings = [{}] + [{"ODIN_FE_VAD_NCTA": str(n), "ODIN_FE_VAD_LPT": l} for n in (1, 2, 4, 8) for l in ("0", "1")]
  for env in settings:
    for k in ("ODIN_FE_VAD_NCTA", "ODIN_FE_VAD_LPT"):
      os.environ.pop(k, None)
    os.environ.update(env)
    for _ in range(3):
      out = fe.run_packed(pcm, off, sr)
    torch.cuda.synchronize()
    acc = np.zeros(4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
      out = fe.run_packed(pcm, off, sr)
    e1.record()
    torch.cuda.synchronize()
    buf = (C.c_float * 4)()
    _lib.check(lib.odin_fe_last_run_ms(h, buf))
    sad = out["sad"].cpu().numpy()
    if ref is None:
      ref = sad
    same = bool(np.array_equal(ref, sad))
    print("%-50s step %.3f ms  dc %.3f frame %.3f post %.3f vad %.3f  sad_same=%s" %
          (env, e0.elapsed_time(e1) / 5, buf[0], buf[1], buf[2], buf[3], same), flush=True)


if __name__ == "__main__":
  main()
