"""Per-kernel milliseconds of the config-3 front-end step (the bench's MFCC leg) at a given corpus size, with checksums of
every output so that two builds can be compared bit for bit (GPU only).

  python tools/fe_step_scale.py [hours ...]          # default: 2 25 100
"""
import ctypes as C
import os
import sys
import zlib

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odin_b200 import _lib, synth  # noqa: E402
from odin_b200 import preprocessing as pp  # noqa: E402


def main():
  sr = 16000
  hours = [float(h) for h in sys.argv[1:]] or [2.0, 25.0, 100.0]
  pipe = pp.make_pipeline([
      pp.AudioReader(), pp.PreEmphasis(0.97), pp.STFTExtractor(0.025, 0.010, n_fft=1024, window="hamm"),
      pp.PowerSpecExtractor(), pp.MelsSpecExtractor(80, fmin=64, fmax=8000),
      pp.MFCCsExtractor(20, first_coef_energy=True), pp.DeltaExtractor("mfcc", order=(0, 1, 2)),
      pp.SADgmm(input_name="stft_energy")])
  fe = pipe.plan[0]
  pool = synth.utterance_batch(24, 5.0, 60.0, sr=sr, seed=4000)
  lens = np.array([len(u) for u in pool], dtype=np.int64)
  one = np.concatenate(pool)
  lib = _lib.load()
  h, _ = fe._handle(sr)
  for hr in hours:
    reps = max(1, int(round(hr * 3600.0 * sr / lens.sum())))
    off = np.zeros(reps * len(pool) + 1, dtype=np.int64)
    np.cumsum(np.tile(lens, reps), out=off[1:])
    pcm = torch.from_numpy(np.tile(one, reps)).cuda()
    for _ in range(3):
      out = fe.run_packed(pcm, off, sr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
      out = fe.run_packed(pcm, off, sr)
    e1.record()
    torch.cuda.synchronize()
    buf = (C.c_float * 4)()
    _lib.check(lib.odin_fe_last_run_ms(h, buf))
    T = int(out["frame_offsets"][-1])
    ms = e0.elapsed_time(e1) / n
    crc = {k: "%08x" % zlib.crc32(v[:4_000_000].contiguous().cpu().numpy().tobytes())
           for k, v in sorted(out.items()) if torch.is_tensor(v) and v.is_cuda}
    print("%6.1f h %5d utt %9d frames  step %.3f ms (%.1f M frames/s)  dc %.3f frame %.3f post %.3f vad %.3f  crc %s" %
          (hr, len(off) - 1, T, ms, T / ms / 1e3, buf[0], buf[1], buf[2], buf[3], crc), flush=True)
    del out, pcm
    torch.cuda.empty_cache()


if __name__ == "__main__":
  main()
