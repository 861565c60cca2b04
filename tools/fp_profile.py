import os, sys, time, cProfile, pstats
sys.path.insert(0, os.getcwd())
import numpy as np
from odin_b200 import preprocessing as pp, synth
sr = 8000
pool = synth.utterance_batch(60, 0.3, 1.0, sr=sr, seed=1)
jobs = [{"raw": pool[i % 60], "sr": sr, "name": "f%04d" % i} for i in range(600)]
def mk():
  return pp.make_pipeline(steps=[
      pp.AudioReader(sr=sr, remove_dc=True), pp.PreEmphasis(coeff=0.97),
      pp.STFTExtractor(frame_length=0.025, step_length=0.005, n_fft=512, window="hamm", energy=False),
      pp.PowerSpecExtractor(power=2.0, output_name="spec"),
      pp.MelsSpecExtractor(n_mels=24, fmin=64, fmax=4000, input_name=("spec", "sr"), output_name="mspec"),
      pp.MFCCsExtractor(n_ceps=20, remove_first_coef=True, first_coef_energy=True, input_name="mspec", output_name="mfcc"),
      pp.DeltaExtractor(input_name="mfcc", order=(0, 1, 2)),
      pp.RenameFeatures(input_name="mfcc_energy", output_name="energy"),
      pp.SADthreshold(energy_threshold=0.55, smooth_window=5, input_name="energy", output_name="sad"),
      pp.DeleteFeatures(input_name=("stft", "spec", "sad_threshold")),
      pp.AcousticNorm(mean_var_norm=True, windowed_mean_var_norm=True, input_name=("mspec", "mfcc")),
      pp.AsType(dtype="float16")])
ex = mk()
for rep in range(3):
  t0 = time.perf_counter()
  feats, indices = pp.FeatureProcessor(jobs, extractor=ex, batch_utts=256).run()
  print("run %d: %.3f s" % (rep, time.perf_counter() - t0), flush=True)
pr = cProfile.Profile(); pr.enable()
feats, indices = pp.FeatureProcessor(jobs, extractor=ex, batch_utts=256).run()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
