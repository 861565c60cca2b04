"""Multi-GPU plumbing of the hot path: one process per GPU, utterances
sharded by rank, ONE sum all-reduce of the packed statistics per EM iteration.

The reference fans the E-step out over forked workers and sums their partial
(Z, F, S, L) through a multiprocessing queue (gmm_tmat.py:249-265, 1199-1220);
here every rank accumulates its own frames on its GPU and the partials meet in a
single `all_reduce(SUM)` of the packed fp64 buffer  Z[M] | F[D,M] | S[D,M] | L | n
(NCCL over NVLink on GPUs; gloo in the CPU tests).  The M-step and the mixture
split are deterministic functions of that buffer, so they are simply replicated.
"""
import numpy as np


def shard_utterances(n_frames, world_size):
  """Greedy longest-first assignment of whole utterances to ranks by frame count.

  n_frames: sequence of per-utterance frame counts (job order).
  Returns a list of `world_size` index lists; every list is in ascending job order
  (the order in which a rank stores its rows), the assignment is deterministic
  (ties broken by job index) and each utterance appears exactly once.
  """
  n_frames = np.asarray(n_frames, dtype=np.int64)
  world_size = int(world_size)
  if world_size < 1:
    raise ValueError("world_size must be >= 1")
  order = sorted(range(len(n_frames)), key=lambda i: (-int(n_frames[i]), i))
  load = [0] * world_size
  out = [[] for _ in range(world_size)]
  for i in order:
    r = min(range(world_size), key=lambda k: (load[k], k))
    out[r].append(i)
    load[r] += int(n_frames[i])
  return [sorted(ix) for ix in out]


def rank_frame_ranges(n_frames_total, indices, rank, world_size):
  """The frame ranges [(start, end), ...] (ascending) of a GLOBAL [N, D] feature matrix that rank `rank` owns.

  This is the device-side stand-in for the reference's job fan-out (gmm_tmat.py:102-133 `_split_jobs`,
  :1165-1220): with `indices` (name -> (start, end), or a list of such pairs) whole utterances are dealt to the
  ranks longest-first (`shard_utterances`); without, the frames are cut into `world_size` contiguous
  ranges.  Every frame belongs to exactly one rank, so summing the ranks' statistics gives the whole-set
  statistics."""
  rank, world_size = int(rank), int(world_size)
  if indices is None:
    n = int(n_frames_total)
    return [((n * rank) // world_size, (n * (rank + 1)) // world_size)]
  items = list(indices.items()) if hasattr(indices, "items") else list(indices)
  spans = sorted(((int(s), int(e)) for _, (s, e) in items), key=lambda x: x[0])
  mine = shard_utterances([e - s for s, e in spans], world_size)[rank]
  out = []
  for i in mine:   # ascending job order; merge neighbours
    s, e = spans[i]
    if out and out[-1][1] == s:
      out[-1] = (out[-1][0], e)
    elif e > s:
      out.append((s, e))
  return out


def take_ranges(a, ranges):
  """Rows of `a` (numpy array / memmap, or None) in the given frame ranges, concatenated."""
  if a is None:
    return None
  if len(ranges) == 1:
    return a[ranges[0][0]:ranges[0][1]]
  return np.concatenate([a[s:e] for s, e in ranges], axis=0)


def pack_stats(Z, F, S, L, n):
  """Z [1,M] | F [D,M] | S [D,M] | sum-LLK | nframes -> the packed fp64 vector the kernels use."""
  return np.concatenate([np.asarray(Z, np.float64).reshape(-1), np.asarray(F, np.float64).reshape(-1),
                         np.asarray(S, np.float64).reshape(-1), [float(L), float(n)]])


def unpack_stats(packed, D, M):
  packed = np.asarray(packed, dtype=np.float64)
  Z = packed[:M].reshape(1, M)
  F = packed[M:M + D * M].reshape(D, M)
  S = packed[M + D * M:M + 2 * D * M].reshape(D, M)
  return Z, F, S, float(packed[-2]), float(packed[-1])


def file_shard(n_files, rank, world_size):
  """Contiguous [lo, hi) range of the files rank `rank` owns in the T-matrix E-step (Tmatrix._estep_device):
  files are equal-cost there (one tv x tv system each), so an even contiguous split is balanced."""
  n_files, rank, world_size = int(n_files), int(rank), int(world_size)
  return (n_files * rank) // world_size, (n_files * (rank + 1)) // world_size


def pack_tmat_stats(LU, RU, llk, nframes):
  """LU [M, t2] | RU [tv, M*D] | llk | nframes -> the packed fp64 vector of odin_tmat_estep (odin_tmat_acc_size)."""
  return np.concatenate([np.asarray(LU, np.float64).reshape(-1), np.asarray(RU, np.float64).reshape(-1),
                         [float(llk), float(nframes)]])


def unpack_tmat_stats(packed, nmix, tv_dim, feat_dim):
  packed = np.asarray(packed, dtype=np.float64)
  t2 = tv_dim * (tv_dim + 1) // 2
  nLU = nmix * t2
  LU = packed[:nLU].reshape(nmix, t2)
  RU = packed[nLU:nLU + tv_dim * nmix * feat_dim].reshape(tv_dim, nmix * feat_dim)
  return LU, RU, float(packed[-2]), float(packed[-1])


def _dist():
  import torch.distributed as td
  if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
    return td
  return None


def allreduce_stats(stats):
  """In-place sum over ranks of the packed statistics (torch tensor, CUDA -> NCCL,
  CPU -> gloo).  No-op for a single process.  Returns `stats`."""
  td = _dist()
  if td is not None:
    td.all_reduce(stats, op=td.ReduceOp.SUM)
  return stats


def gather_rows(local_rows, local_index, n_total):
  """Per-utterance statistics (gmm_tmat.py:769-913) need no collective on the data path: each
  rank owns the rows of its utterances.  This host-side helper assembles the [n_total, width]
  matrix in JOB order on every rank from each rank's (rows, job indices)."""
  local_rows = np.asarray(local_rows)
  td = _dist()
  if td is None:
    out = np.zeros((n_total, local_rows.shape[1]), dtype=local_rows.dtype)
    out[np.asarray(local_index, dtype=np.int64)] = local_rows
    return out
  parts = [None] * td.get_world_size()
  td.all_gather_object(parts, (np.asarray(local_index, dtype=np.int64), local_rows))
  out = np.zeros((n_total, local_rows.shape[1]), dtype=local_rows.dtype)
  for idx, rows in parts:
    out[idx] = rows
  return out
