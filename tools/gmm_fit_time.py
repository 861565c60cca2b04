"""Wall time of the reference's whole split-and-train schedule (GMM.fit, gmm_tmat.py:625-699: M = 1 -> 2048, 71 EM
iterations) on frames resident on one B200."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from odin_b200 import synth
from odin_b200.ml import GMM
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
X = torch.from_numpy(synth.gmm_features(N, 60, 64, seed=3)).cuda()
for rep in range(2):
  g = GMM(nmix=2048, nmix_start=1, niter=10)
  torch.cuda.synchronize(); t0 = time.perf_counter()
  g.fit(X)
  torch.cuda.synchronize(); dt = time.perf_counter() - t0
  nit = sum(len(v) for v in g._llk_hist.values())
  print("fit 2048-mix UBM on %d frames: %.2f s, %d EM iterations, final llk %.4f" % (N, dt, nit, g._llk_hist[2048][-1]), flush=True)
