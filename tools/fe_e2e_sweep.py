import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from odin_b200 import synth, preprocessing as pp
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
pinned = torch.from_numpy(pcm_h).pin_memory()
for nc in (1, 4, 6, (1, 3, 3, 2, 1), (1, 2, 3, 3, 2, 1), (1, 2, 2, 2, 2, 1), (2, 3, 3, 2), (1, 2, 4, 4, 2, 1)):
  out = None
  ts = []
  for _ in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = fe.run_host_packed(pinned, off, sr, want=("feat", "sad"), n_chunks=nc, out=out)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
  T = int(out["frame_offsets"][-1])
  print("chunks %s: %.2f ms  %.1f M frames/s" % (str(nc), min(ts[1:]) * 1e3, T / min(ts[1:]) / 1e6), flush=True)
