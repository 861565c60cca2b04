"""CPU, world_size = 2 over gloo: the N > 1 host logic of SURVEY.md 8e -- utterance sharding,
the single all-reduce of the packed statistics, the replicated M-step and the job-order gather
of per-utterance rows.  The oracle stands in for the kernels (it is the checker here, as
everywhere in tests/); what is exercised is odin_b200.sharding, the code bench.py and
GMM._estep_device run between the kernels and NCCL."""
import os
import socket

import numpy as np
import pytest

from odin_b200 import sharding, synth
from oracle import gmm as OG

D, M, N_UTT = 12, 8, 23


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _data():
  rng = np.random.RandomState(3)
  lens = rng.randint(20, 400, size=N_UTT)
  X = synth.gmm_features(int(lens.sum()), D, 4, seed=9)
  off = np.concatenate([[0], np.cumsum(lens)])
  mean, sigma, w = synth.gmm_params(D, M, seed=10)
  return lens, off, X, mean, sigma, w


def _worker(rank, world, port, q):
  import torch
  import torch.distributed as td
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  td.init_process_group("gloo", rank=rank, world_size=world)
  try:
    lens, off, X, mean, sigma, w = _data()
    mine = sharding.shard_utterances(lens, world)[rank]
    rows = np.concatenate([np.arange(off[i], off[i + 1]) for i in mine])
    z, f, s, l, n = OG.expectation(X[rows], mean, sigma, w, compute_dtype=np.float64)
    stats = torch.from_numpy(sharding.pack_stats(z, f, s, l * n, n))   # L travels as a SUM, like the kernels
    sharding.allreduce_stats(stats)
    Z, F, S, Lsum, nfr = sharding.unpack_stats(stats.numpy(), D, M)
    m1, s1, w1, rb = OG.maximization(Z, F, S, (mean, sigma, w))         # replicated M-step
    # per-utterance rows gathered in job order
    zu = np.stack([OG.transform(X[off[i]:off[i + 1]], mean, sigma, w, compute_dtype=np.float64)[0].reshape(-1)
                   for i in mine])
    allz = sharding.gather_rows(zu, mine, N_UTT)
    q.put((rank, stats.numpy().copy(), m1, s1, w1, allz, mine))
  finally:
    td.destroy_process_group()


def test_shard_utterances_properties():
  rng = np.random.RandomState(0)
  lens = rng.randint(1, 6000, size=101)
  for world in (1, 2, 4, 8):
    sh = sharding.shard_utterances(lens, world)
    assert len(sh) == world
    flat = sorted(i for part in sh for i in part)
    assert flat == list(range(len(lens)))                       # a partition
    assert all(part == sorted(part) for part in sh)             # job order inside a rank
    loads = [int(lens[part].sum()) for part in sh]
    assert max(loads) - min(loads) <= int(lens.max())           # greedy LPT bound
    assert sh == sharding.shard_utterances(lens, world)         # deterministic
  assert sharding.shard_utterances([], 2) == [[], []]
  assert sharding.shard_utterances([5], 4) == [[0], [], [], []]


def test_pack_unpack_roundtrip():
  rng = np.random.RandomState(1)
  Z, F, S = rng.rand(1, M), rng.rand(D, M), rng.rand(D, M)
  p = sharding.pack_stats(Z, F, S, -12.5, 77)
  assert p.shape[0] == (2 * D + 1) * M + 2
  Z2, F2, S2, L2, n2 = sharding.unpack_stats(p, D, M)
  assert np.array_equal(Z, Z2) and np.array_equal(F, F2) and np.array_equal(S, S2) and L2 == -12.5 and n2 == 77


@pytest.mark.timeout(120)
def test_two_ranks_allreduce_matches_single_process():
  import torch.multiprocessing as mp
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  got = [q.get(timeout=100) for _ in procs]
  for p in procs:
    p.join(timeout=30)
    assert p.exitcode == 0
  got.sort(key=lambda t: t[0])
  lens, off, X, mean, sigma, w = _data()
  z, f, s, l, n = OG.expectation(X, mean, sigma, w, compute_dtype=np.float64)
  whole = sharding.pack_stats(z, f, s, l * n, n)
  for rank, stats, m1, s1, w1, allz, mine in got:
    assert np.allclose(stats, whole, rtol=1e-12, atol=1e-12)      # sum of shards == whole job
  # the replicated M-step gives bit-identical models on both ranks (no broadcast needed)
  assert np.array_equal(got[0][2], got[1][2]) and np.array_equal(got[0][3], got[1][3])
  assert np.array_equal(got[0][4], got[1][4])
  # job-order gather: same matrix on both ranks, equal to the single-process rows
  ref = np.stack([OG.transform(X[off[i]:off[i + 1]], mean, sigma, w, compute_dtype=np.float64)[0].reshape(-1)
                  for i in range(N_UTT)])
  assert np.array_equal(got[0][5], got[1][5]) and np.allclose(got[0][5], ref, rtol=1e-12)
  assert sorted(got[0][6] + got[1][6]) == list(range(N_UTT))
