"""GPU debug helper: E-step of impl 1/2/3 vs the fp64 oracle on a few shapes, printing errors."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odin_b200 import synth
from odin_b200.ml import GMM
from oracle import gmm as OG

def relmax(a, b):
  return float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(float(np.max(np.abs(b))), 1e-30))

def run(D, M, N, impls=(2, 3), sad=None):
  X = synth.gmm_features(N, D, 8, seed=D + M)
  mean, sigma, w = synth.gmm_params(D, M, seed=M)
  z, f, s, l, n = OG.expectation(X, mean, sigma, w, sad=sad, compute_dtype=np.float64)
  for impl in impls:
    try:
      g = GMM(nmix=M, nmix_start=M, impl=impl)
      g.initialize(np.zeros((1, D), dtype=np.float32))
      g.mean, g.sigma, g.w = mean.copy(), sigma.copy(), w.copy()
      Z, F, S, L = g.expectation(X, sad=sad)
      print("D=%d M=%d N=%d impl=%d: Z %.2e F %.2e S %.2e  L %.6f (ref %.6f)  sumZ %.3f (n %d)" % (
          D, M, N, impl, relmax(Z, z), relmax(F, f), relmax(S, s), float(L), float(l), Z.sum(), n), flush=True)
    except Exception as e:
      print("D=%d M=%d N=%d impl=%d: FAILED %s" % (D, M, N, impl, e), flush=True)
      raise

if __name__ == "__main__":
  run(60, 256, 128)
  run(60, 256, 64)
  run(60, 256, 5000)
  run(60, 512, 4099)
  run(60, 2048, 3000)
  run(40, 384, 2049)
  run(4, 300, 500)
  rng = np.random.RandomState(3)
  run(60, 512, 20000, sad=(rng.rand(20000) > 0.4).astype(np.uint8))
