"""A/B of the packed-f32x2 frame kernel (fe_frame5_kernel) against the scalar four-step kernel (ODIN_FE_FRAME4=1) on
the three front-end configurations: per-kernel milliseconds and the largest difference of every output (GPU only)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odin_b200 import _lib, synth  # noqa: E402
from odin_b200 import preprocessing as pp  # noqa: E402


def chain(sr, n_fft, n_mels, fmax, frame=0.025, hop=0.010, vad="gmm"):
  steps = [pp.AudioReader(), pp.PreEmphasis(0.97), pp.STFTExtractor(frame, hop, n_fft=n_fft, window="hamm"),
           pp.PowerSpecExtractor(), pp.MelsSpecExtractor(n_mels, fmin=64, fmax=fmax),
           pp.MFCCsExtractor(20, first_coef_energy=True), pp.DeltaExtractor("mfcc", order=(0, 1, 2))]
  steps.append(pp.SADgmm(input_name="stft_energy") if vad == "gmm" else pp.SADthreshold(input_name="mfcc_energy"))
  return pp.make_pipeline(steps)


def run(name, pipe, utts, sr, reps=5):
  fe = pipe.plan[0]
  pcm_h, off = synth.pack_utterances(utts)
  pcm = torch.from_numpy(pcm_h).cuda()
  lib = _lib.load()
  h, _ = fe._handle(sr)
  res = {}
  for tag, env in (("frame4", "1"), ("frame5", "0")):
    os.environ["ODIN_FE_FRAME4"] = env
    for _ in range(3):
      out = fe.run_packed(pcm, off, sr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      out = fe.run_packed(pcm, off, sr)
    e1.record()
    torch.cuda.synchronize()
    buf = (C.c_float * 4)()
    _lib.check(lib.odin_fe_last_run_ms(h, buf))
    res[tag] = {k: v.cpu().numpy() for k, v in out.items() if torch.is_tensor(v)}
    nfr = res[tag]["sad"].shape[0]
    print("%-8s %-7s %8d frames  step %.3f ms  dc %.3f frame %.3f post %.3f vad %.3f  (%.1f M frames/s)" %
          (name, tag, nfr, e0.elapsed_time(e1) / reps, buf[0], buf[1], buf[2], buf[3],
           nfr / (e0.elapsed_time(e1) / reps) / 1e3), flush=True)
  a, b = res["frame4"], res["frame5"]
  for k in sorted(a):
    x, y = a[k].astype(np.float64), b[k].astype(np.float64)
    if a[k].dtype == np.uint8 or a[k].dtype == np.bool_:
      print("   %-12s differing entries: %d of %d" % (k, int((a[k] != b[k]).sum()), a[k].size))
    else:
      print("   %-12s max|d| %.3e  max|ref| %.3e  rel %.2e" % (k, np.abs(x - y).max(), np.abs(x).max(),
                                                            np.abs(x - y).max() / max(np.abs(x).max(), 1e-30)))


def main():
  pool = synth.utterance_batch(24, 5.0, 60.0, sr=16000, seed=4000)
  run("cfg3", chain(16000, 1024, 80, 8000), [pool[i % len(pool)] for i in range(9 * len(pool))], 16000)
  short = synth.utterance_batch(50, 3.0, 3.0, sr=16000, seed=7)
  run("cfg1", chain(16000, 512, 40, 8000), [short[i % 50] for i in range(2000)], 16000)
  dig = synth.utterance_batch(100, 0.3, 1.0, sr=8000, seed=11)
  run("cfg5", chain(8000, 512, 24, 4000, 0.025, 0.005, vad="thr"), [dig[i % 100] for i in range(3000)], 8000)
  run("n256", chain(8000, 256, 24, 4000, 0.025, 0.010), [dig[i % 100] for i in range(3000)], 8000)


if __name__ == "__main__":
  main()
