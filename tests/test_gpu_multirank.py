"""On-hardware multi-rank parity: the same answer at 1 and 2 ranks.

Runs tests/_multirank_worker.py under torch.distributed.run with 2 ranks over NCCL (skipped on a box with one GPU):
every rank passes the same global (X, indices, sad); the all-reduced N/F/S must equal the single-GPU whole-set
statistics, the replicated M-step must be bit-identical across ranks, and the sharded per-utterance statistics /
T-matrix must match the single-GPU ones."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import relmax

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
  import torch
  return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_ranks_equal_one_rank(tmp_path):
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_multirank_worker.py"), str(tmp_path)]
  out = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
  assert out.returncode == 0, out.stdout[-4000:]
  r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(2)]
  for k in range(2):
    assert int(r[k]["world"]) == 2
    # all-reduced statistics of the sharded run == whole-set statistics of one GPU.  Not bit-identical: the fp16
    # operand scales of the tensor-core E-step are chosen per rank from its own frames, well inside 1e-5
    for a, b in (("Z2", "Z1"), ("F2", "F1"), ("S2", "S1")):
      assert relmax(r[k][a], r[k][b]) < 1e-5, (a, relmax(r[k][a], r[k][b]))
    assert abs(float(r[k]["L2"]) - float(r[k]["L1"])) < 1e-6 * abs(float(r[k]["L1"]))
    for a, b in (("mean2", "mean1"), ("sigma2", "sigma1"), ("w2", "w1")):
      assert relmax(r[k][a], r[k][b]) < 1e-4, a
    # per-utterance statistics under the two fitted models (which differ at the 1e-5 level after two EM iterations)
    assert relmax(r[k]["zu2"], r[k]["zu1"]) < 1e-4 and relmax(r[k]["fu2"], r[k]["fu1"]) < 1e-4
    sgn = np.sign(np.sum(r[k]["T2"] * r[k]["T1"], axis=1, keepdims=True))
    assert relmax(r[k]["T2"] * sgn, r[k]["T1"]) < 1e-4
  # replicated M-step / T-matrix update: bit-identical on both ranks
  for name in ("Z2", "F2", "S2", "mean2", "sigma2", "w2", "T2", "zu2", "fu2"):
    assert np.array_equal(r[0][name], r[1][name]), name
  assert list(r[0]["names"]) == ["utt%03d" % i for i in range(40)]
  # FeatureProcessor(shard=True): the two ranks split the 12 jobs, each wrote its own store; fitting on the rank-local
  # stores (local_shard) gives the model of fitting on the whole job list
  names = sorted(list(r[0]["fp_names"]) + list(r[1]["fp_names"]))
  assert names == ["j%02d" % i for i in range(12)] and len(r[0]["fp_names"]) > 0 and len(r[1]["fp_names"]) > 0
  for k in range(2):
    assert relmax(r[k]["fp_mean_local"], r[k]["fp_mean_global"]) < 1e-4
    assert relmax(r[k]["fp_sigma_local"], r[k]["fp_sigma_global"]) < 1e-3
  assert np.array_equal(r[0]["fp_mean_local"], r[1]["fp_mean_local"])
  # odin_gmm_allreduce (C-ABI, its own NCCL communicator) == torch.distributed.all_reduce, bit for bit
  assert bool(r[0]["ar_equal"]) and bool(r[1]["ar_equal"])
