import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from odin_b200.ml import GMM
N, M, D = 6_000_000, 2048, 60
X, mean, sigma, w = bench.make_ubm_and_frames(torch, N, M, 1000, "cuda")
Xh = torch.empty((N, D), dtype=torch.float32).pin_memory(); Xh.copy_(X); del X
for ch in (1 << 20, 1 << 19, 1 << 18, 3 << 18):
  os.environ['ODIN_GMM_CHUNK_FRAMES'] = str(ch)
  g = GMM(nmix=M, nmix_start=M)
  g.initialize(Xh[:1000].numpy())
  ts = []
  for i in range(4):
    g.mean, g.sigma, g.w = mean, sigma, w
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g.expectation_maximization(Xh, print_progress=False)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
  print("chunk %8d: %.2f ms  %.1f M frames/s" % (ch, min(ts[1:]) * 1e3, N / min(ts[1:]) / 1e6), flush=True)
  del g
