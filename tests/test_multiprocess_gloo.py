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


# ---------------------------------------------------------------------------
# T-matrix E-step: contiguous file shards, one all-reduce of LU | RU | llk | nframes
# ---------------------------------------------------------------------------
def _tmat_data():
  from oracle import tmatrix as OT
  from oracle.make_golden import _tmat_problem
  sigma, Z, F = _tmat_problem(seed=17, D=5, M=6, n_files=31)
  Z = np.round(Z)                       # integer frame counts: ceil(sum Z) is then shard-independent
  Sigma = OT.sigma_row(sigma)
  T0 = OT.init_T(4, Sigma)
  T_invS, T_invS_Tt = OT.refresh(T0, Sigma, 5)
  return sigma, Z, F, T_invS, T_invS_Tt


def _tmat_worker(rank, world, port, q):
  import torch
  import torch.distributed as td
  from oracle import tmatrix as OT
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  td.init_process_group("gloo", rank=rank, world_size=world)
  try:
    sigma, Z, F, T_invS, T_invS_Tt = _tmat_data()
    lo, hi = sharding.file_shard(Z.shape[0], rank, world)
    LU, RU, llk, nfr = OT.expectation(Z[lo:hi], F[lo:hi], T_invS, T_invS_Tt)
    stats = torch.from_numpy(sharding.pack_tmat_stats(LU, RU, llk, nfr))
    sharding.allreduce_stats(stats)
    LU, RU, llk, nfr = sharding.unpack_tmat_stats(stats.numpy(), 6, 4, 5)
    T1 = OT.maximization(LU, RU, nfr, 5)                      # replicated M-step
    q.put((rank, stats.numpy().copy(), T1, (lo, hi)))
  finally:
    td.destroy_process_group()


def test_file_shard_properties():
  for n in (0, 1, 7, 31, 1000):
    for world in (1, 2, 3, 8):
      parts = [sharding.file_shard(n, r, world) for r in range(world)]
      assert parts[0][0] == 0 and parts[-1][1] == n
      assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))           # contiguous partition
      sizes = [hi - lo for lo, hi in parts]
      assert max(sizes) - min(sizes) <= 1
  rng = np.random.RandomState(2)
  LU, RU = rng.rand(6, 10), rng.rand(4, 30)
  p = sharding.pack_tmat_stats(LU, RU, -3.5, 1234)
  assert p.shape[0] == 6 * 10 + 4 * 30 + 2
  a, b, c, d = sharding.unpack_tmat_stats(p, 6, 4, 5)
  assert np.array_equal(a, LU) and np.array_equal(b, RU) and c == -3.5 and d == 1234


@pytest.mark.timeout(120)
def test_two_ranks_tmatrix_estep_matches_single_process():
  import torch.multiprocessing as mp
  from oracle import tmatrix as OT
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_tmat_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  got = [q.get(timeout=100) for _ in procs]
  for p in procs:
    p.join(timeout=30)
    assert p.exitcode == 0
  got.sort(key=lambda t: t[0])
  sigma, Z, F, T_invS, T_invS_Tt = _tmat_data()
  LU, RU, llk, nfr = OT.expectation(Z, F, T_invS, T_invS_Tt)
  whole = sharding.pack_tmat_stats(LU, RU, llk, nfr)
  for rank, stats, T1, (lo, hi) in got:
    assert np.allclose(stats, whole, rtol=1e-11, atol=1e-11)
  assert np.array_equal(got[0][2], got[1][2])                    # bit-identical replicated M-step
  assert got[0][3][1] == got[1][3][0] and got[1][3][1] == Z.shape[0]


def test_rank_frame_ranges_partition_the_global_frame_axis():
  """sharding.rank_frame_ranges is what GMM.fit / expectation use under torch.distributed to take a rank's share of a
  GLOBAL (X, indices): every frame of every utterance belongs to exactly one rank, frames outside `indices` to none."""
  rng = np.random.RandomState(4)
  lens = rng.randint(1, 500, size=57)
  gaps = rng.randint(0, 3, size=57)            # some utterances leave holes between them
  starts = np.cumsum(np.concatenate([[5], lens[:-1] + gaps[:-1]]))
  indices = {"u%d" % i: (int(s), int(s + n)) for i, (s, n) in enumerate(zip(starts, lens))}
  n_total = int(starts[-1] + lens[-1] + 7)
  covered = np.zeros(n_total, dtype=np.int32)
  for s, e in indices.values():
    covered[s:e] += 1
  for world in (1, 2, 3, 8):
    seen = np.zeros(n_total, dtype=np.int32)
    for rank in range(world):
      rr = sharding.rank_frame_ranges(n_total, indices, rank, world)
      assert rr == sorted(rr) and all(e > s for s, e in rr)
      for s, e in rr:
        seen[s:e] += 1
      x = np.arange(n_total)
      assert np.array_equal(sharding.take_ranges(x, rr), np.concatenate([x[s:e] for s, e in rr]) if rr else x[:0]) or not rr
    assert np.array_equal(seen, covered)
    seen = np.zeros(n_total, dtype=np.int32)
    for rank in range(world):
      (s, e), = sharding.rank_frame_ranges(n_total, None, rank, world)
      seen[s:e] += 1
    assert np.all(seen == 1)
