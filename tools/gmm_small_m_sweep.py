"""E-step time and accuracy of the three kernel families at small mixture counts (GPU only): the tensor-core
paths pad M up to their tile (128 / 256 mixtures), the fp32 path does not.  Run with
ODIN_GMM_TC_MIN_M=1 ODIN_GMM_H_MIN_M=1 to lift the dispatch thresholds."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from odin_b200 import synth
from odin_b200.ml import GMM
from odin_b200.ml.gmm import _DeviceFrames
from oracle import gmm as OG

N, D = 1_000_000, 60
X = torch.from_numpy(synth.gmm_features(N, D, 16, seed=3)).cuda()
XF = _DeviceFrames(X)
Xs = X[:20000].cpu().numpy()
for M in (1, 2, 8, 32, 64, 128, 256):
  mean, sigma, w = synth.gmm_params(D, M, seed=5)
  ref = OG.expectation(Xs, mean, sigma, w, compute_dtype=np.float64)
  line = "M %4d:" % M
  for impl in (1, 2, 3):
    try:
      g = GMM(nmix=M, nmix_start=M, impl=impl)
      g.initialize(Xs)
      g.mean, g.sigma, g.w = mean, sigma, w
      Z, F, S, L = g.expectation(Xs)
      err = max(float(np.abs(a - b).max() / np.abs(b).max()) for a, b in ((Z, ref[0]), (F, ref[1]), (S, ref[2])))
      for _ in range(2):
        g._estep_device(XF, None, True)
      torch.cuda.synchronize(); t0 = time.perf_counter()
      for _ in range(5):
        g._estep_device(XF, None, True)
      torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
      line += "  impl %d %.2f ms (err %.1e)" % (impl, dt * 1e3, err)
    except Exception as e:
      line += "  impl %d n/a (%s)" % (impl, str(e)[:40])
  print(line, flush=True)
