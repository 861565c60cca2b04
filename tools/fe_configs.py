"""Front-end timings at the other BASELINE.json configurations (GPU only): config 1 (100 x 3 s, n_fft 512, 40 mel,
SADgmm) and config 5 (3 000 digits of 0.3-1 s at 8 kHz, 25/5 ms, n_fft 512, 24 mel, SADthreshold), PCM resident."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odin_b200 import _lib, synth  # noqa: E402
from odin_b200 import preprocessing as pp  # noqa: E402


def run(name, pipe, utts, sr):
  fe = pipe.plan[0]
  pcm_h, off = synth.pack_utterances(utts)
  pcm = torch.from_numpy(pcm_h).cuda()
  lib = _lib.load()
  h, _ = fe._handle(sr)
  for _ in range(3):
    out = fe.run_packed(pcm, off, sr)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10):
    out = fe.run_packed(pcm, off, sr)
  e1.record()
  torch.cuda.synchronize()
  buf = (C.c_float * 4)()
  _lib.check(lib.odin_fe_last_run_ms(h, buf))
  T = int(out["frame_offsets"][-1])
  ms = e0.elapsed_time(e1) / 10
  print("%s: %d utterances, %d frames, %.3f ms per batch = %.1f M frames/s (dc %.3f frame %.3f post %.3f vad %.3f)" %
        (name, len(utts), T, ms, T / ms / 1e3, buf[0], buf[1], buf[2], buf[3]), flush=True)


def main():
  p1 = pp.make_pipeline([
      pp.AudioReader(), pp.PreEmphasis(0.97), pp.STFTExtractor(0.025, 0.010, n_fft=512, window="hamm"),
      pp.PowerSpecExtractor(), pp.MelsSpecExtractor(40, fmin=64, fmax=8000),
      pp.MFCCsExtractor(20, first_coef_energy=True), pp.DeltaExtractor("mfcc", order=(0, 1, 2)),
      pp.SADgmm(input_name="stft_energy")])
  pool = synth.utterance_batch(20, 3.0, 3.0, sr=16000, seed=11)
  run("config 1", p1, [pool[i % 20] for i in range(100)], 16000)
  run("config 1 x 20", p1, [pool[i % 20] for i in range(2000)], 16000)
  p5 = pp.make_pipeline([
      pp.AudioReader(), pp.PreEmphasis(0.97), pp.STFTExtractor(0.025, 0.005, n_fft=512, window="hamm"),
      pp.PowerSpecExtractor(), pp.MelsSpecExtractor(24, fmin=64, fmax=4000),
      pp.MFCCsExtractor(20, first_coef_energy=True), pp.DeltaExtractor("mfcc", order=(0, 1, 2)),
      pp.SADthreshold(input_name="mfcc_energy")])
  pool = synth.utterance_batch(60, 0.3, 1.0, sr=8000, seed=12)
  run("config 5", p5, [pool[i % 60] for i in range(3000)], 8000)


if __name__ == "__main__":
  main()
