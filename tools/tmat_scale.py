"""T-matrix EM iteration at growing tv_dim, up to the NIST-SRE recipe's scale (2048 mixtures, 60 dimensions, tv 400-600),
on statistics resident in HBM (the bench's T-matrix leg with the sizes as arguments): milliseconds per phase and the
fp64 rate of the E-step (GPU only).

  TMAT_M=2048 TMAT_FILES=2000 python tools/tmat_scale.py [tv ...]      # default tv: 64 128 200 400 600
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odin_b200 import _lib  # noqa: E402
from odin_b200.ml import GMM, Tmatrix  # noqa: E402


def main():
  tvs = [int(v) for v in sys.argv[1:]] or [64, 128, 200, 400, 600]
  M, D, n = int(os.environ.get("TMAT_M", 2048)), 60, int(os.environ.get("TMAT_FILES", 2000))
  rng = np.random.RandomState(0)
  g = GMM(nmix=M, nmix_start=M)
  g.initialize(np.zeros((4, D), dtype=np.float32))
  g.sigma = 0.5 + rng.rand(D, M)
  gen = torch.Generator(device="cuda").manual_seed(99)
  Z = torch.empty((n, M), dtype=torch.float64, device="cuda").uniform_(0.0, 4.0, generator=gen)
  F = torch.randn((n, M * D), dtype=torch.float64, device="cuda", generator=gen) * Z.repeat_interleave(D, 1).sqrt()
  lib = _lib.load()
  for tv in tvs:
    t = Tmatrix(tv, g, niter=1)
    acc = torch.zeros(t._acc_size, dtype=torch.float64, device="cuda")
    T0 = t.Tm
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    te = tm = 0.0
    reps = 2
    for it in range(1 + reps):
      t._upload(T0)
      acc.zero_()
      ev[0].record()
      _lib.check(lib.odin_tmat_estep(t._h, _lib.ptr(Z), _lib.ptr(F), n, _lib.ptr(acc), _lib.current_stream()))
      ev[1].record()
      t._mstep_device(acc, True, True)
      ev[2].record()
      torch.cuda.synchronize()
      if it > 0:
        te += ev[0].elapsed_time(ev[1]) / reps
        tm += ev[1].elapsed_time(ev[2]) / reps
    t2 = tv * (tv + 1) // 2
    flops_file = 2 * M * t2 * 2 + 2 * M * D * tv * 2 + tv ** 3   # the bench's count: L1 + LU, B1 + RU, factorise / invert / product
    print("tv %4d  M %d  D %d  files %d: E-step %8.1f ms (%.2f TFLOP/s fp64)  M-step %8.1f ms  -> %.0f files/s" %
          (tv, M, D, n, te, flops_file * n / te / 1e9, tm, n / (te + tm) * 1e3), flush=True)
    del t, acc
    torch.cuda.empty_cache()


if __name__ == "__main__":
  main()
