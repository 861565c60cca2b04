"""Worker of tests/test_gpu_multirank.py: launched by torch.distributed.run with one rank per GPU (NCCL).

Each rank first computes the whole-set answer alone on its own GPU (no process group yet), then joins the group
and repeats the computation the way a recipe would under torchrun: every rank passes the SAME global arrays and
GMM / Tmatrix take their share (sharding.rank_frame_ranges / shard_utterances).  Results go to
<outdir>/rank<k>.npz for the parent test to compare."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(outdir):
  import torch
  rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
  torch.cuda.set_device(local)
  from odin_b200 import synth
  from odin_b200.ml import GMM, Tmatrix
  D, M = 60, 64
  rng = np.random.RandomState(11)
  lens = rng.randint(200, 3000, size=40)
  off = np.concatenate([[0], np.cumsum(lens)])
  # frames from 16 well separated components; the 64-mixture model starts at perturbed copies of the true means, so
  # every mixture keeps occupancy (the T-matrix M-step needs every per-mixture system to be positive definite)
  mu = rng.randn(16, D) * 3.0
  X = (mu[rng.randint(0, 16, size=int(off[-1]))] + rng.randn(int(off[-1]), D)).astype(np.float32)
  sad = (rng.rand(int(off[-1])) > 0.2).astype(np.uint8)
  indices = [("utt%03d" % i, (int(off[i]), int(off[i + 1]))) for i in range(len(lens))]
  mean = (np.tile(mu, (4, 1)) + 0.5 * rng.randn(M, D)).T.astype(np.float32).copy()
  sigma = np.full((D, M), 1.5, dtype=np.float32)
  w = np.full((1, M), 1.0 / M, dtype=np.float32)

  def model():
    g = GMM(nmix=M, nmix_start=M, niter=2)
    g.initialize(X)
    g.mean, g.sigma, g.w = mean.copy(), sigma.copy(), w.copy()
    return g

  # ---- alone (no process group): whole-set statistics, two EM iterations, per-utterance statistics
  g1 = model()
  Z1, F1, S1, L1 = g1.expectation((X, indices), sad=sad)
  g1.expectation_maximization((X, indices), sad=sad, print_progress=False)
  g1.expectation_maximization((X, indices), sad=sad, print_progress=False)
  g1.transform_to_disk(X, indices, sad=sad)
  zu1, fu1 = g1.last_utt_stats_
  t1 = Tmatrix(4, g1, niter=1)
  t1.expectation_maximization(zu1.astype(np.float64), fu1.astype(np.float64))
  T1 = t1.Tm
  # ---- the same calls under NCCL
  import torch.distributed as td
  td.init_process_group("nccl", device_id=torch.device("cuda", local))
  g2 = model()
  Z2, F2, S2, L2 = g2.expectation((X, indices), sad=sad)
  g2.expectation_maximization((X, indices), sad=sad, print_progress=False)
  g2.expectation_maximization((X, indices), sad=sad, print_progress=False)
  zp, fp = os.path.join(outdir, "Z.npy"), os.path.join(outdir, "F.npy")
  names = g2.transform_to_disk(X, indices, sad=sad, pathZ=zp, pathF=fp)
  zu2, fu2 = np.load(zp), np.load(fp)
  t2 = Tmatrix(4, g2, niter=1)
  t2.expectation_maximization(zu2.astype(np.float64), fu2.astype(np.float64))
  # ---- features stay with the GPU that extracted them: FeatureProcessor shards the jobs, GMM(local_shard=True) fits on
  # each rank's own store; the model must be the one a single process gets from the whole job list
  from odin_b200 import preprocessing as pp
  pool = synth.utterance_batch(12, 0.6, 1.6, sr=8000, seed=5)
  jobs = [{"raw": pool[i], "sr": 8000, "name": "j%02d" % i} for i in range(12)]
  chain = lambda: pp.make_pipeline([pp.AudioReader(), pp.PreEmphasis(0.97), pp.STFTExtractor(0.025, 0.010, n_fft=256, window="hamm"),
                                    pp.PowerSpecExtractor(), pp.MelsSpecExtractor(24, fmin=64, fmax=4000), pp.MFCCsExtractor(12)])
  fdir = os.path.join(outdir, "feats")
  feats, idx = pp.FeatureProcessor(jobs, path=fdir, extractor=chain(), batch_utts=4).run()      # this rank's share
  gl = GMM(nmix=4, nmix_start=1, niter=2, local_shard=True)
  gl.fit((np.ascontiguousarray(feats["mfcc"], dtype=np.float32), idx["mfcc"]))
  local_names = sorted(idx["mfcc"])
  td.barrier()
  full, idx_all = pp.FeatureProcessor(jobs, extractor=chain(), batch_utts=4, shard=False).run()  # every job, in memory
  g_all = GMM(nmix=4, nmix_start=1, niter=2)                       # global input: the GMM takes its own share
  g_all.fit((np.ascontiguousarray(full["mfcc"], dtype=np.float32), idx_all["mfcc"]))
  # ---- the C-ABI collective (odin_gmm_allreduce) on an NCCL communicator of its own, against torch.distributed
  import ctypes as C
  from odin_b200 import _lib
  nccl = C.CDLL("libnccl.so.2")          # the library torch has already loaded

  class UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]

  uid = UniqueId()
  if rank == 0:
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
  t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device="cuda")
  td.broadcast(t, 0)
  C.memmove(C.byref(uid), bytes(t.cpu().numpy().tolist()), 128)
  comm = C.c_void_p()
  nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
  assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
  nst = (2 * g2._kdim + 1) * M + 2
  mine = torch.arange(nst, dtype=torch.float64, device="cuda") * (rank + 1) + 0.25 * rank
  want = mine.clone()
  td.all_reduce(want)
  _lib.check(_lib.load().odin_gmm_allreduce(g2._handle, _lib.ptr(mine), comm, _lib.current_stream()))
  torch.cuda.synchronize()
  ar_equal = bool(torch.equal(mine, want))
  nccl.ncclCommDestroy(comm)
  np.savez(os.path.join(outdir, "rank%d.npz" % rank), ar_equal=ar_equal, fp_names=np.array(local_names),
           fp_mean_local=gl.mean, fp_mean_global=g_all.mean, fp_sigma_local=gl.sigma, fp_sigma_global=g_all.sigma, Z1=Z1, F1=F1, S1=S1, L1=L1, Z2=Z2, F2=F2, S2=S2, L2=L2,
           mean1=g1.mean, sigma1=g1.sigma, w1=g1.w, mean2=g2.mean, sigma2=g2.sigma, w2=g2.w,
           zu1=zu1, fu1=fu1, zu2=zu2, fu2=fu2, T1=T1, T2=t2.Tm, names=np.array(names), world=world)
  td.barrier()
  td.destroy_process_group()


if __name__ == "__main__":
  main(sys.argv[1])
