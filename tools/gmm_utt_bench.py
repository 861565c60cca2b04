"""odin_gmm_utt_stats: fp32 CUDA-core route (impl 1) against the tcgen05 routes (impl 0: per-utterance launches for long
utterances, the segmented launch for short ones) -- long (6 000-18 000 frames) or short (config-5 digits) utterances.

  python tools/gmm_utt_bench.py [long|mid|short] [M]      (ODIN_GMM_UTT_SEG=1 forces the segmented route)"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from odin_b200.ml import GMM
from odin_b200 import _lib
kind = sys.argv[1] if len(sys.argv) > 1 else "long"
D, M, n_utt = 60, int(sys.argv[2]) if len(sys.argv) > 2 else (2048 if kind == "long" else 512), (200 if kind == "long" else (1000 if kind == "mid" else 3000))
rng = np.random.RandomState(0)
lens = rng.randint(6000, 18000, size=n_utt) if kind == "long" else (rng.randint(500, 6000, size=n_utt) if kind == "mid" else rng.randint(60, 200, size=n_utt))
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
X = torch.randn((int(off[-1]), D), device="cuda")
g = GMM(nmix=M, nmix_start=M)
g.initialize(X[:100].cpu().numpy())
g.mean = (rng.randn(D, M) * 2).astype(np.float32); g.sigma = (0.5 + rng.rand(D, M)).astype(np.float32); g.w = np.full((1, M), 1.0 / M, np.float32)
for impl in (1, 0):
  g.impl = impl
  for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Z, Fh = g._utt_stats_device(X, None, off)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
  print("impl %d: %d utterances, %d frames: %.1f ms = %.1f M frames/s" % (impl, n_utt, off[-1], dt * 1e3, off[-1] / dt / 1e6), flush=True)
  if impl == 1: Z1, F1 = Z.clone(), Fh.clone()
print("max rel diff Z %.2e F %.2e" % (float((Z - Z1).abs().max() / Z1.abs().max()), float((Fh - F1).abs().max() / F1.abs().max())))
