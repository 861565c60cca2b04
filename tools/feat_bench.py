"""HBM throughput of the feature-matrix kernels on a ragged batch resident on the device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from odin_b200 import _lib

lib = _lib.load()
_lib.require_cuda()
rng = np.random.RandomState(0)
n_utt, dim = 2000, 60
lens = rng.randint(500, 1500, size=n_utt)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
T = int(off[-1])
x = torch.randn((T, dim), device="cuda")
st = _lib.current_stream()


def timed(name, fn, nbytes):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10):
    fn()
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / 10
  print("%-34s %.3f ms  %.0f GB/s algorithmic (%d frames)" % (name, ms, nbytes / ms / 1e6, T), flush=True)


c = 5
y = torch.empty((T, (2 * c + 1) * dim), device="cuda")
timed("StackFeatures(n_context=5)", lambda: _lib.check(lib.odin_feat_stack(_lib.ptr(x), _lib.ptr(y), dim, _lib.as_i64_ptr(off), n_utt, c, st)),
      T * dim * 4 * (1 + 2 * c + 1))
y2 = torch.empty((T, dim), device="cuda")
timed("RASTAfilter(rasta, sdc=0)", lambda: _lib.check(lib.odin_feat_rasta_sdc(_lib.ptr(x), _lib.ptr(y2), dim, _lib.as_i64_ptr(off), n_utt, 1, 0, st)),
      T * dim * 4 * 2)
x20 = torch.randn((T, 20), device="cuda")
y3 = torch.empty((T, 20 + 400), device="cuda")
timed("RASTAfilter(rasta, sdc=1), 20 ceps", lambda: _lib.check(lib.odin_feat_rasta_sdc(_lib.ptr(x20), _lib.ptr(y3), 20, _lib.as_i64_ptr(off), n_utt, 1, 1, st)),
      T * 4 * (20 + 420))
fr = torch.randn((T // 4, 400), device="cuda")
e = torch.empty(T // 4, device="cuda")
timed("CalculateEnergy([T/4, 400])", lambda: _lib.check(lib.odin_feat_energy(_lib.ptr(fr), _lib.ptr(e), T // 4, 400, 1, st)),
      (T // 4) * 401 * 4)
sad = (torch.rand(T, device="cuda") > 0.4).to(torch.uint8)
yn = torch.empty((T, dim), device="cuda")
timed("AcousticNorm(mvn + wmvn 301)", lambda: _lib.check(lib.odin_fe_cmvn(_lib.ptr(x), _lib.ptr(yn), dim, _lib.as_i64_ptr(off), n_utt, None, 1, 1, 1, 301, st)),
      T * dim * 4 * 2)
