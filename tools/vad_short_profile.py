import os, sys
sys.path.insert(0, os.getcwd())
import torch
from odin_b200 import synth, preprocessing as pp
p1 = pp.make_pipeline([
    pp.AudioReader(), pp.PreEmphasis(0.97), pp.STFTExtractor(0.025, 0.010, n_fft=512, window="hamm"),
    pp.PowerSpecExtractor(), pp.MelsSpecExtractor(40, fmin=64, fmax=8000),
    pp.MFCCsExtractor(20, first_coef_energy=True), pp.DeltaExtractor("mfcc", order=(0, 1, 2)),
    pp.SADgmm(input_name="stft_energy")])
pool = synth.utterance_batch(20, 3.0, 3.0, sr=16000, seed=11)
pcm_h, off = synth.pack_utterances([pool[i % 20] for i in range(2000)])
pcm = torch.from_numpy(pcm_h).cuda()
fe = p1.plan[0]
for _ in range(3):
  fe.run_packed(pcm, off, 16000)
torch.cuda.synchronize()
