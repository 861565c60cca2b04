"""Diagonal GMM-UBM with the surface of ``odin.ml.GMM`` (reference:
odin/ml/gmm_tmat.py:270-1338) over the B200 kernels of libodin_b200.

Same constructor keywords, attributes (``mean [D,M]``, ``sigma [D,M]`` =
VARIANCE, ``w [1,M]``), methods and pickle tuple as the reference, so a recipe
switches by changing the import.  All arithmetic of the E-step / M-step / split
runs in CUDA (``odin_gmm_*`` in include/odin_b200.h); this file is host logic
only: argument handling, the split-and-train schedule (gmm_tmat.py:625-699),
batch selection for ``downsample`` (gmm_tmat.py:135-168), host<->device staging
and the NCCL all-reduce that replaces the reference's multiprocessing sum
(gmm_tmat.py:249-265, 1199-1220).

There is no CPU path: ``device`` is accepted for signature compatibility only.
"""
import os
import pickle
import random
from collections import defaultdict
from collections.abc import Mapping

import numpy as np

from .. import _lib
from .. import sharding

EPS = 1e-6  # gmm_tmat.py:27

_NITER_SCHEDULE = [1, 2, 4, 4, 4, 4, 6, 6, 10, 10, 10, 10, 10, 16, 16]  # gmm_tmat.py:677


def _torch():
  import torch
  return torch


def _dist():
  td = _torch().distributed
  if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
    return td
  return None


def minibatch(n, batch_size):
  """odin.utils.minibatch: contiguous (start, end) ranges."""
  batch_size = int(batch_size)
  return [(s, min(s + batch_size, n)) for s in range(0, n, batch_size)]


class _DeviceFrames(object):
  """Feeds [N, D] float32 frames to the kernels: a CUDA tensor is used in place,
  a host array is streamed through two pinned buffers on a copy stream so the
  H2D copy of chunk i+1 overlaps the E-step of chunk i.

  float16 host data -- what the recipes store (AsType('float16'), examples/fsdd_ivec.py:105,197)
  -- crosses PCIe at its stored width and is widened on the device (odin_feat_convert): half the bytes of the
  host-side up-cast this class used to do, on a path that is PCIe / host-DRAM bound."""

  def __init__(self, X, chunk_frames=1 << 18):
    _lib.require_cuda()
    torch = _torch()
    self.torch = torch
    self.resident = None
    self.host = None
    self.host_pinned = None
    if isinstance(X, torch.Tensor):
      if not X.is_cuda:
        if X.is_pinned() and X.dtype in (torch.float32, torch.float16) and X.is_contiguous() and X.dim() == 2:
          self.host_pinned = X
        X = X.numpy()
      else:
        if X.dtype != torch.float32 or not X.is_contiguous():
          X = X.to(torch.float32).contiguous()
        self.resident = X
    if self.resident is None:
      X = np.asarray(X) if not isinstance(X, np.memmap) else X
      if X.ndim != 2:
        raise ValueError("`X` must be a 2-D matrix [n_samples, feat_dim]")
      self.host = X
    self.n, self.dim = (self.resident.shape if self.resident is not None else self.host.shape)
    self.kdim = int(self.dim)   # row width handed to the kernels (GMM._kdim: zero columns appended), see set_kdim
    self.chunk = int(os.environ.get('ODIN_GMM_CHUNK_FRAMES', chunk_frames))   # (env: A/B runs of the streaming depth)

  shape = property(lambda self: (self.n, self.dim))
  ndim = 2

  def set_kdim(self, k):
    """Rows are delivered `k` >= dim wide, the extra columns zero (GMM._kdim)."""
    k = int(k)
    if k != self.kdim:
      assert k >= self.dim and self._prepared is None
      if self.resident is not None:
        out = self.torch.zeros((self.n, k), dtype=self.torch.float32, device=self.resident.device)
        out[:, :self.dim] = self.resident[:, :self.dim]
        self.resident = out
      self.kdim = k
    return self
  _prepared = None
  _prepared_failed = False
  reuse = False   # set by callers that keep this object across EM iterations (fit, bench)

  def prepared(self, gmm_handle):
    """Tensor-core operand images of the RESIDENT frames (odin_gmm_frames_create), built on first
    use and kept for the following EM iterations; None when they do not apply / do not fit."""
    if self.resident is None or self._prepared_failed or self.n == 0:
      return None
    if self._prepared is None:
      lib = _lib.load()
      h = _lib.C.c_void_p()
      rc = lib.odin_gmm_frames_create(gmm_handle, _lib.ptr(self.resident), self.n, _lib.C.byref(h),
                                      _lib.current_stream())
      if rc < 0:
        self._prepared_failed = True   # ODIN_ENOMEM / unsupported D: fall back to odin_gmm_estep
        return None
      self._prepared = h
    return self._prepared

  def __del__(self):
    try:
      if self._prepared is not None:
        _lib.load().odin_gmm_frames_destroy(self._prepared)
        self._prepared = None
    except Exception:
      pass

  def cache_on_device(self, reserve_bytes=2 << 30):
    """Upload once if it fits (used by fit(): EM re-reads the data every iteration)."""
    if self.resident is not None:
      return True
    torch = self.torch
    need = self.n * self.kdim * 4
    free, _ = torch.cuda.mem_get_info()
    if need + reserve_bytes > free:
      return False
    out = torch.empty((self.n, self.kdim), dtype=torch.float32, device="cuda")
    for dev, s, e in self.chunks():
      out[s:e].copy_(dev)
    self.resident, self.host, self.host_pinned = out, None, None
    return True

  def chunks(self):
    """Yields (device tensor [n_i, D] float32, start, end)."""
    torch = self.torch
    if self.resident is not None:
      yield self.resident, 0, self.n
      return
    n, D, K = self.n, self.dim, self.kdim
    if n == 0:
      return
    ch = min(self.chunk, n)
    direct = self.host_pinned is not None  # pinned float32 / float16 tensor: DMA straight from the caller's buffer
    half = np.dtype(self.host.dtype) == np.float16   # stored width on the wire, widened on the device
    wire = torch.float16 if half else torch.float32
    pinned = None if direct else [torch.empty((ch, D), dtype=wire).pin_memory() for _ in range(2)]
    dev = [torch.zeros((ch, K), dtype=torch.float32, device="cuda") for _ in range(2)]   # pad columns stay zero
    dev_wire = [torch.empty((ch, D), dtype=wire, device="cuda") for _ in range(2)] if (half or K != D) else dev
    lib = _lib.load()
    copy_stream = torch.cuda.Stream()
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    # Chunk size: the first copy and the last E-step cannot overlap anything, so smaller chunks shorten the exposed
    # ends until per-chunk overheads win (6 M frames at 2048 mixtures, tools/gmm_e2e_sweep.py: 1 M-frame chunks
    # 29.7 ms per EM iteration, 512 k 28.5, 256 k 27.7, 128 k 29.6; ramped sizes were no better than fixed ones).
    ranges = minibatch(n, ch)

    def stage(i):
      s, e = ranges[i]
      b = i & 1
      if direct:
        src = self.host_pinned[s:e]
      else:
        copied[b].synchronize()  # the previous DMA out of pinned[b] has finished
        np.copyto(pinned[b][:e - s].numpy(), self.host[s:e], casting="unsafe")
        src = pinned[b][:e - s]
      with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(consumed[b])  # kernels of chunk i-2 are done with dev[b]
        dev_wire[b][:e - s].copy_(src, non_blocking=True)
        copied[b].record(copy_stream)

    for b in range(2):
      consumed[b].record()
      copied[b].record(copy_stream)
    stage(0)
    for i, (s, e) in enumerate(ranges):
      if i + 1 < len(ranges):
        stage(i + 1)
      b = i & 1
      torch.cuda.current_stream().wait_event(copied[b])
      if K != D:     # widen (if float16) and place into the first D columns of the padded rows
        dev[b][:e - s, :D].copy_(dev_wire[b][:e - s])
      elif half:
        _lib.check(lib.odin_feat_convert(_lib.ptr(dev_wire[b]), 0, _lib.ptr(dev[b]), 1, (e - s) * D, _lib.current_stream()))
      yield dev[b][:e - s], s, e
      consumed[b].record()


class GMM(object):
  r"""Gaussian Mixture Model with diagonal covariance (see module docstring).

  Parameters follow gmm_tmat.py:341-346.  Extra keyword ``impl`` selects the
  kernel family: 0 auto, 1 fp32 CUDA cores, 2 3xTF32 tcgen05, 3 3xFP16 tcgen05.

  Multi-GPU (one process per GPU under ``torch.distributed``): by default every rank is handed the SAME global
  ``X`` / ``indices`` / ``sad`` -- exactly what a recipe written for the reference passes -- and takes its own
  share of the utterances (``sharding.rank_frame_ranges``, the device-side form of ``_split_jobs``,
  gmm_tmat.py:102-133); the packed statistics meet in one all-reduce per EM iteration and the M-step is
  replicated.  ``local_shard=True`` says the caller already holds only this rank's frames (the front-end keeps
  features on the GPU that extracted them).
  """

  STANDARD_CPU_BATCH_SIZE = 12 * 1024 * 1024  # gmm_tmat.py:338-339
  STANDARD_GPU_BATCH_SIZE = 25 * 1024 * 1024

  def __init__(self, nmix, nmix_start=1, niter=16, dtype='float32',
               allow_rollback=True, exit_on_error=False,
               batch_size_cpu='auto', batch_size_gpu='auto',
               downsample=1, stochastic_downsample=True,
               device='gpu', ncpu=1, gpu_factor=80,
               seed=1234, path=None, name=None, impl=0, local_shard=False):
    self.local_shard = bool(local_shard)
    self._path = path if isinstance(path, str) else None
    nmix = int(nmix)
    if nmix < 1:
      raise ValueError("Number of Mixture must be greater than 1.")
    self._nmix = nmix
    self._curr_nmix = int(np.clip(int(nmix_start), 1, self._nmix))
    self._feat_dim = None
    self._niter = int(niter)
    self.batch_size_cpu = batch_size_cpu
    self.batch_size_gpu = batch_size_gpu
    self.downsample = int(downsample)
    self.stochastic_downsample = bool(stochastic_downsample)
    self._seed = int(seed)
    self.gpu_factor = int(gpu_factor)
    self.ncpu = int(ncpu) if ncpu is not None else 1
    self.set_device(device)
    self._llk_hist = defaultdict(list)
    self.allow_rollback = bool(allow_rollback)
    self.exit_on_error = bool(exit_on_error)
    self._stop_fitting = False
    self._dtype = np.dtype(dtype)
    self._name = ('GMM_%08x' % random.getrandbits(32)) if name is None else str(name)
    self.impl = int(impl)
    self._init_device_state()

  # ------------------------------------------------------------------ state
  @property
  def _kdim(self):
    """Feature dimension the KERNELS see.  The tensor-core E-step needs D % 4 == 0 (16-byte frame rows); any
    other D up to 60 (e.g. 39 = 13 MFCC + deltas) is carried with zero columns up to the next multiple of four and a
    unit-variance, zero-mean model in those columns: they add exactly zero to every Mahalanobis term and the same
    constant to every mixture's log-likelihood (removed again in `_unpack_stats` / the scoring calls), so posteriors
    and statistics are unchanged -- and the frames run on tcgen05 instead of the 13x slower fp32 kernels."""
    D = self._feat_dim
    if D % 4 == 0 or D > 60 or self.impl not in (0, 3) or os.environ.get("ODIN_GMM_NO_PAD", "0") == "1":
      return D
    return (D + 3) & ~3

  def _pad_frames(self, dev):
    """[n, D] CUDA float32 -> [n, kdim] with zero columns appended (no-op when kdim == D)."""
    k = self._kdim
    if k == self._feat_dim:
      return dev
    out = _torch().zeros((dev.shape[0], k), dtype=dev.dtype, device=dev.device)
    out[:, :self._feat_dim] = dev
    return out

  def _fix_pad_model(self, var_value):
    """Pad rows of the device model: mean 0, variance `var_value` (1 - EPS for the E-step: precision exactly 1 and
    log(var + EPS) = 0; 0 around the mix-up so that a pad column is never the split direction)."""
    D, k, M = self._feat_dim, self._kdim, self._curr_nmix
    if k == D:
      return
    self._d_mean[:k * M].view(k, M)[D:] = 0.0
    self._d_var[:k * M].view(k, M)[D:] = var_value
    _lib.check(_lib.load().odin_gmm_set_params(self._handle, M, _lib.ptr(self._d_mean), _lib.ptr(self._d_var),
                                               _lib.ptr(self._d_w), _lib.current_stream()))

  def _frames(self, X):
    """X (array / tensor / _DeviceFrames) -> _DeviceFrames delivering rows of the kernel width."""
    fr = X if isinstance(X, _DeviceFrames) else _DeviceFrames(X)
    fr.reuse = fr.reuse or isinstance(X, _DeviceFrames)
    return fr.set_kdim(self._kdim)

  def _init_device_state(self):
    self._handle = None
    self._d_mean = self._d_var = self._d_w = None
    self._d_stats = None
    self._d_flag = None
    self._synced = (None, None, None)

  def __getstate__(self):  # gmm_tmat.py:388-400 (same 21-tuple)
    if not self.is_initialized:
      raise RuntimeError("GMM hasn't been initialized, nothing to save")
    return (self.mean, self.sigma, self.w,
            self.allow_rollback, self.exit_on_error,
            self._nmix, self._curr_nmix, self._feat_dim,
            self._niter, self.batch_size_cpu, self.batch_size_gpu,
            self.downsample, self.stochastic_downsample,
            self._seed, self._llk_hist,
            self.ncpu, self._device, self.gpu_factor,
            self._dtype, self._path, self._name)

  def __setstate__(self, states):
    (self.mean, self.sigma, self.w,
     self.allow_rollback, self.exit_on_error,
     self._nmix, self._curr_nmix, self._feat_dim,
     self._niter, self.batch_size_cpu, self.batch_size_gpu,
     self.downsample, self.stochastic_downsample,
     self._seed, self._llk_hist,
     self.ncpu, self._device, self.gpu_factor,
     self._dtype, self._path, self._name) = states
    self._stop_fitting = False
    self.impl = 0
    self.local_shard = False
    self._feat_const = self._feat_dim * np.log(2 * np.pi)
    self._init_device_state()

  def __del__(self):
    try:
      if getattr(self, "_handle", None) is not None:
        _lib.load().odin_gmm_destroy(self._handle)
        self._handle = None
    except Exception:
      pass

  def __str__(self):
    if not self.is_initialized:
      return '<"%s" nmix:%d initialized:False>' % (self.name, self._nmix)
    return '<"%s" nmix:%s ndim:%s mean:%s std:%s w:%s>' % (
        self.name, self._nmix, self._feat_dim, self.mean.shape, self.sigma.shape, self.w.shape)

  # ------------------------------------------------------------- properties
  def set_device(self, device):
    device = str(device).lower()
    if device not in ('cpu', 'gpu', 'mix'):
      raise ValueError("`device` must be one of the following: 'cpu', 'gpu', or 'mix'")
    self._device = device  # kept for pickle compatibility; compute is always CUDA
    return self

  device = property(lambda self: self._device)
  path = property(lambda self: self._path)
  name = property(lambda self: self._name)
  is_initialized = property(lambda self: self._feat_dim is not None)
  is_fitted = property(lambda self: self._curr_nmix == self._nmix)
  nmix = property(lambda self: self._nmix)
  dtype = property(lambda self: self._dtype)
  history = property(lambda self: tuple(self._llk_hist))

  @property
  def feat_dim(self):
    if not self.is_initialized:
      raise RuntimeError("GMM has not been initialized on data.")
    return self._feat_dim

  def get_params(self, deep=True):
    return dict(nmix=self._nmix, nmix_start=self._curr_nmix, niter=self._niter, dtype=self._dtype,
                allow_rollback=self.allow_rollback, exit_on_error=self.exit_on_error,
                batch_size_cpu=self.batch_size_cpu, batch_size_gpu=self.batch_size_gpu,
                downsample=self.downsample, stochastic_downsample=self.stochastic_downsample,
                device=self._device, ncpu=self.ncpu, gpu_factor=self.gpu_factor, seed=self._seed,
                path=self._path, name=self._name)

  # --------------------------------------------------------- device plumbing
  def _ensure_handle(self):
    _lib.require_cuda()
    torch = _torch()
    lib = _lib.load()
    if self._handle is None:
      import ctypes as C
      h = C.c_void_p()
      _lib.check(lib.odin_gmm_create(self._kdim, self._nmix, C.byref(h)))
      self._handle = h
      D, M = self._kdim, self._nmix
      self._d_mean = torch.empty(D * M, dtype=torch.float32, device="cuda")
      self._d_var = torch.empty(D * M, dtype=torch.float32, device="cuda")
      self._d_w = torch.empty(M, dtype=torch.float32, device="cuda")
      self._d_flag = torch.zeros(1, dtype=torch.int32, device="cuda")
      self._synced = (None, None, None)
    return lib

  def _upload_params(self, force=False):
    """Host attributes -> device when they were replaced since the last sync."""
    lib = self._ensure_handle()
    if not force and self._synced[0] is self.mean and self._synced[1] is self.sigma \
        and self._synced[2] is self.w:
      return lib
    torch = _torch()
    D, M = self._feat_dim, self._curr_nmix
    for dst, src, n in ((self._d_mean, self.mean, D * M), (self._d_var, self.sigma, D * M),
                        (self._d_w, self.w, M)):
      a = np.ascontiguousarray(np.asarray(src, dtype=np.float32)).reshape(-1)
      if a.shape[0] != n:
        raise ValueError("parameter has %d elements, expected %d" % (a.shape[0], n))
      dst[:n].copy_(torch.from_numpy(a))
    if self._kdim != D:
      self._fix_pad_model(1.0 - EPS)   # (also calls odin_gmm_set_params)
    else:
      _lib.check(lib.odin_gmm_set_params(self._handle, M, _lib.ptr(self._d_mean), _lib.ptr(self._d_var),
                                         _lib.ptr(self._d_w), _lib.current_stream()))
    self._synced = (self.mean, self.sigma, self.w)
    return lib

  def _download_params(self):
    D, M = self._feat_dim, self._curr_nmix   # (the first D of the kdim rows)
    self.mean = self._d_mean[:D * M].cpu().numpy().reshape(D, M).astype(self._dtype, copy=False)
    self.sigma = self._d_var[:D * M].cpu().numpy().reshape(D, M).astype(self._dtype, copy=False)
    self.w = self._d_w[:M].cpu().numpy().reshape(1, M).astype(self._dtype, copy=False)
    self._synced = (self.mean, self.sigma, self.w)

  def _resfresh_cpu_posterior(self):
    """Reference name (gmm_tmat.py:493-504): call after editing mean/sigma/w in
    place; re-uploads the model and refreshes the cached constants on device."""
    self._upload_params(force=True)

  _refresh_gpu_posterior = _resfresh_cpu_posterior

  def _stats_buffer(self):
    torch = _torch()
    n = (2 * self._kdim + 1) * self._curr_nmix + 2
    if self._d_stats is None or self._d_stats.shape[0] != n:
      self._d_stats = torch.zeros(n, dtype=torch.float64, device="cuda")
    return self._d_stats

  def _llk_pad_const(self):
    """What the pad columns take off every frame's log-likelihood: 1/2 (kdim - D) log 2 pi."""
    return 0.5 * (self._kdim - self._feat_dim) * np.log(2 * np.pi)

  def _unpack_stats(self, stats_host, zero, first, second, llk):
    D, K, M = self._feat_dim, self._kdim, self._curr_nmix
    Z = stats_host[:M].reshape(1, M)
    F = stats_host[M:M + K * M].reshape(K, M)[:D]
    S = stats_host[M + K * M:M + 2 * K * M].reshape(K, M)[:D]
    L, nfr = stats_host[-2] + self._llk_pad_const() * stats_host[-1], stats_host[-1]
    out = []
    if zero:
      out.append(Z)
    if first:
      out.append(F)
    if second:
      out.append(S)
    if llk:
      out.append(np.array(L / nfr if nfr > 0 else 0.0))
    return out[0] if len(out) == 1 else out

  # --------------------------------------------------------- initialization
  def initialize(self, X):  # gmm_tmat.py:562-622
    indices = None
    if isinstance(X, (tuple, list)):
      tmp = [i for i in X if hasattr(i, 'shape')][0]
      rest = [i for i in X if i is not tmp]
      indices = rest[0] if rest else None
      X = tmp
    if isinstance(indices, Mapping):
      indices = list(indices.items())
    if not hasattr(X, 'shape') or len(X.shape) != 2:
      raise ValueError("`X` must be a 2-D array [n_samples, feat_dim]")
    feat_dim = int(X.shape[1])
    if self.is_initialized:
      if feat_dim != self._feat_dim:
        raise RuntimeError("Input must be 2-D matrix with the 1st dimension equal to: %d" % self._feat_dim)
      return X, indices
    self._feat_dim = feat_dim
    self._feat_const = self._feat_dim * np.log(2 * np.pi)
    if isinstance(self.batch_size_cpu, str):
      self.batch_size_cpu = int(GMM.STANDARD_CPU_BATCH_SIZE / (self._feat_dim * self._dtype.itemsize))
    if isinstance(self.batch_size_gpu, str):
      self.batch_size_gpu = int(GMM.STANDARD_GPU_BATCH_SIZE / (self._feat_dim * self._dtype.itemsize))
    self.mean = np.zeros((feat_dim, self._curr_nmix), dtype=self._dtype)
    self.sigma = np.ones((feat_dim, self._curr_nmix), dtype=self._dtype)
    self.w = np.ones((1, self._curr_nmix), dtype=self._dtype)
    return X, indices

  # ---------------------------------------------------------------- E-step
  def _selected_mask(self, n, sad, indices):
    """Frame selection of gmm_tmat.py:135-232 as a uint8 mask (None = all).  With `downsample > 1` the picks are
    those of the reference's default single-job fan-out (ncpu=1; gmm_tmat.py:102-133 splits the frame axis into
    `ncpu` jobs that each re-seed and shuffle their own batches, so its picks depend on `ncpu`): batches of
    int(batch_size_cpu / floor(2 ** (curr_nmix / 1024))) frames -- or whole files with `indices` -- shuffled with
    `random.seed(seed + curr_nmix + curr_niter)`, the first always kept, the others with probability 1 / downsample."""
    mask = None
    if indices is not None:
      mask = np.zeros(n, dtype=np.uint8)
      for _, (s, e) in indices:
        mask[int(s):int(e)] = 1
    if self.downsample > 1:
      curr_niter = len(self._llk_hist[self._curr_nmix])
      random.seed(int(self._seed + self._curr_nmix + curr_niter) if self.stochastic_downsample
                  else int(self._seed))
      if indices is None:
        reduction = np.floor(np.power(2, self._curr_nmix / 1024))
        units = minibatch(n, int(self.batch_size_cpu / reduction))
      else:
        # the reference hands its (single) job the files popped from the END of the list (gmm_tmat.py:1171-1181),
        # i.e. in reverse order, and shuffles that list (gmm_tmat.py:185)
        units = [(int(s), int(e)) for _, (s, e) in indices][::-1]
      random.shuffle(units)
      keep = np.zeros(n, dtype=np.uint8)
      for i, (s, e) in enumerate(units):
        if i == 0 or random.random() <= 1. / self.downsample:
          keep[s:e] = 1
      mask = keep if mask is None else (mask & keep)
    if sad is not None:
      sad = (np.asarray(sad).reshape(-1) != 0).astype(np.uint8)
      mask = sad if mask is None else (mask & sad)
    return mask

  def _shard(self, X, sad, indices):
    """Global (X, sad, indices) -> this rank's (X_local, ranges); `ranges` is None when nothing is cut (single
    process, `local_shard`, or data that already lives on this rank's GPU)."""
    td = _dist()
    if td is None or self.local_shard or isinstance(X, _DeviceFrames):
      return X, None
    torch = _torch()
    if isinstance(X, torch.Tensor) and X.is_cuda:
      return X, None   # a CUDA tensor is this rank's own data by construction
    ranges = sharding.rank_frame_ranges(X.shape[0], indices, td.get_rank(), td.get_world_size())
    return sharding.take_ranges(X, ranges), ranges

  def _local_mask(self, n_global, sad, indices, ranges):
    """Selection mask over this rank's frames.  The selection itself (SAD, `indices`, down-sampling picks) is
    made over the GLOBAL frame axis with the reference's seeded draws, so it does not depend on the number of
    ranks; each rank then keeps the part that falls into its ranges."""
    mask = self._selected_mask(n_global, sad, indices)
    if mask is None or ranges is None:
      return mask
    return np.ascontiguousarray(sharding.take_ranges(mask, ranges))

  def _estep_device(self, frames, mask, second=True):
    """Runs the E-step kernels over `frames` (a _DeviceFrames); returns the packed
    fp64 statistics tensor on device, all-reduced over ranks."""
    torch = _torch()
    lib = self._upload_params()
    stats = self._stats_buffer()
    stats.zero_()
    d_mask = None
    if mask is not None:
      if isinstance(mask, torch.Tensor):
        d_mask = mask.to(device="cuda", dtype=torch.uint8).contiguous()
      else:
        d_mask = torch.from_numpy(np.ascontiguousarray(mask)).cuda()
    frames.set_kdim(self._kdim)
    prep = None
    if self.impl in (0, 3) and self._kdim % 4 == 0 and self._kdim <= 60 \
        and os.environ.get("ODIN_H_NO_PREPARED", "0") != "1":   # (the images depend on the data only: every stage of fit)
      prep = frames.prepared(self._handle) if frames.reuse else None
    if prep is not None:
      _lib.check(lib.odin_gmm_estep_frames(self._handle, prep, _lib.ptr(d_mask), 1 if second else 0,
                                           _lib.ptr(stats), _lib.current_stream()))
    else:
      for dev, s, e in frames.chunks():
        sad_ptr = None if d_mask is None else _lib.C.c_void_p(d_mask.data_ptr() + s)
        _lib.check(lib.odin_gmm_estep(self._handle, _lib.ptr(dev), sad_ptr, e - s, 1 if second else 0,
                                      _lib.ptr(stats), self.impl, _lib.current_stream()))
    sharding.allreduce_stats(stats)  # gmm_tmat.py:249-265 -> one NCCL all-reduce per EM iteration
    return stats

  def _fast_expectation(self, X, zero=True, first=True, second=True, llk=True, on_gpu=True):
    """gmm_tmat.py:997-1041 (L is the SUM of frame log-likelihoods here, as in the reference)."""
    self.initialize(X)
    stats = self._estep_device(self._frames(X), None, second).cpu().numpy()
    out = self._unpack_stats(stats, zero, first, second, False)
    out = [out] if not isinstance(out, list) else out
    if llk:
      out.append(stats[-2])
    return out if len(out) > 1 else out[0]

  def expectation(self, X, sad=None, zero=True, first=True, second=True, llk=True,
                  device=None, print_progress=True):
    """gmm_tmat.py:1043-1231 -> Z [1,M], F [D,M], S [D,M], L (mean log-likelihood)."""
    X, indices = self.initialize(X)
    n_global = X.shape[0]
    if sad is not None:
      assert sad.shape[0] == n_global, \
          "Number of samples for X and sad mismatch X.shape=%s and sad.shape=%s" % (tuple(X.shape), sad.shape)
    Xl, ranges = self._shard(X, sad, indices)
    frames = self._frames(Xl)
    mask = self._local_mask(n_global, sad, indices, ranges)
    stats = self._estep_device(frames, mask, second).cpu().numpy()
    return self._unpack_stats(stats, zero, first, second, llk)

  # ---------------------------------------------------------------- M-step
  def maximization(self, Z, F, S, floor_const=None):
    """gmm_tmat.py:1233-1276 on device (fp64) from host statistics."""
    torch = _torch()
    D, K, M = self._feat_dim, self._kdim, self._curr_nmix
    pad = np.zeros(((K - D), M), dtype=np.float64)
    packed = np.concatenate([np.asarray(Z, dtype=np.float64).reshape(-1),
                             np.asarray(F, dtype=np.float64).reshape(D, M).reshape(-1), pad.reshape(-1),
                             np.asarray(S, dtype=np.float64).reshape(D, M).reshape(-1), pad.reshape(-1), [0.0, 0.0]])
    assert packed.shape[0] == (2 * K + 1) * M + 2
    self._upload_params()
    stats = self._stats_buffer()
    stats.copy_(torch.from_numpy(packed))
    return self._maximization_device(stats, floor_const)

  def _maximization_device(self, stats, floor_const=None):
    lib = self._ensure_handle()
    if floor_const is not None:
      # gmm_tmat.py:1255-1257 floors the new variances at (sigma w^T) * floor_const BEFORE the negative-variance
      # test (:1259-1272), so a positive floor prevents the rollback.  The M-step kernel computes
      # sigma = S / (Z + EPS) - mu^2; flooring is the same as raising S to (floor + mu^2) (Z + EPS), which is done
      # here on the packed statistics (O(D M) host work; not used by `fit`).
      D, M = self._kdim, self._curr_nmix
      st = stats.cpu().numpy().copy()
      Z, F, S = st[:M].reshape(1, M), st[M:M + D * M].reshape(D, M), st[M + D * M:M + 2 * D * M].reshape(D, M)
      iN = 1.0 / (Z + EPS)
      mu = F * iN
      sig = S * iN - mu * mu
      vfloor = sig.dot((Z / Z.sum()).T) * floor_const
      S[...] = np.maximum(sig, vfloor) / iN + mu * mu / iN
      stats.copy_(_torch().from_numpy(st))
    _lib.check(lib.odin_gmm_mstep(self._handle, _lib.ptr(stats), 1 if self.allow_rollback else 0,
                                  _lib.ptr(self._d_mean), _lib.ptr(self._d_var), _lib.ptr(self._d_w),
                                  _lib.ptr(self._d_flag), _lib.current_stream()))
    self._fix_pad_model(1.0 - EPS)
    self._download_params()
    if int(self._d_flag.item()) != 0 and self.exit_on_error:
      self._stop_fitting = True
    return self

  def expectation_maximization(self, X, sad=None, device=None, print_progress=True):
    """gmm_tmat.py:1278-1306."""
    X, indices = self.initialize(X)
    n_global = X.shape[0]
    Xl, ranges = self._shard(X, sad, indices)
    frames = self._frames(Xl)
    curr_nmix = self._curr_nmix
    mask = self._local_mask(n_global, sad, indices, ranges)
    stats = self._estep_device(frames, mask, True)
    tail = stats[-2:].cpu().numpy()
    L = float(tail[0] / tail[1]) if tail[1] > 0 else 0.0
    self._maximization_device(stats)
    self._llk_hist[curr_nmix].append(L)
    if print_progress:
      print("#mix:%.2d #iter:%.2d llk:%.4f" % (curr_nmix, len(self._llk_hist[curr_nmix]), L))
    self._checkpoint()
    return self

  def gmm_mixup(self):
    """gmm_tmat.py:1308-1338 on device."""
    if self._curr_nmix >= self._nmix:
      return
    lib = self._upload_params()
    new_m = min(2 * self._curr_nmix, self._nmix)
    self._fix_pad_model(0.0)           # a pad column must never be the direction of the split
    _lib.check(lib.odin_gmm_mixup(self._handle, new_m, _lib.ptr(self._d_mean), _lib.ptr(self._d_var),
                                  _lib.ptr(self._d_w), _lib.current_stream()))
    self._curr_nmix = new_m
    self._fix_pad_model(1.0 - EPS)
    self._download_params()
    self._checkpoint()
    return self

  def _checkpoint(self):
    if self._path is not None:
      td = _dist()
      if td is None or td.get_rank() == 0:
        with open(self._path, 'wb') as f:
          pickle.dump(self, f)

  # ------------------------------------------------------------------- fit
  def fit(self, X, y=None, print_progress=False):
    """gmm_tmat.py:625-699: split-and-train schedule.  `X` is an array (numpy or
    CUDA tensor) or a tuple (X, sad) / (X, indices) / (X, sad, indices)."""
    if not isinstance(X, (tuple, list)):
      X = (X,)
    sad = None
    indices = None
    if len(X) == 1:
      data = X[0]
    elif len(X) == 2:
      if hasattr(X[1], 'shape') and X[0].shape[0] == X[1].shape[0]:
        data, sad = X
      else:
        data, indices = X
    elif len(X) == 3:
      data, sad, indices = X
    else:
      raise ValueError("No support for `X` in type of list with length: %d" % len(X))
    assert hasattr(data, 'shape') and len(data.shape) == 2, \
        'Input data must be instance of 2-D ndarray but give: %s' % str(type(data))
    if indices is not None:
      if isinstance(indices, Mapping):
        indices = list(indices.items())
      indices = sorted(indices, key=lambda x: x[1][0])
    self.initialize(data)
    n_global = data.shape[0]
    local, ranges = self._shard(data, sad, indices)   # multi-GPU: this rank's utterances (gmm_tmat.py:102-133)
    frames = self._frames(local)
    frames.cache_on_device()
    frames.reuse = True
    niter = list(_NITER_SCHEDULE)
    niter[int(np.log2(self._nmix))] = self._niter
    self._stop_fitting = False
    while True:
      curr_nmix = self._curr_nmix
      curr_niter = niter[int(np.log2(curr_nmix))] - len(self._llk_hist[curr_nmix])
      for _ in range(max(curr_niter, 0)):
        self._em_frames(frames, sad, indices, print_progress, n_global, ranges)
        if self._stop_fitting:
          return self
      if curr_nmix < self._nmix:
        self.gmm_mixup()
      else:
        break
    return self

  def _em_frames(self, frames, sad, indices, print_progress, n_global=None, ranges=None):
    curr_nmix = self._curr_nmix
    mask = self._local_mask(frames.n if n_global is None else n_global, sad, indices, ranges)
    stats = self._estep_device(frames, mask, True)
    tail = stats[-2:].cpu().numpy()
    L = float(tail[0] / tail[1]) if tail[1] > 0 else 0.0
    self._maximization_device(stats)
    self._llk_hist[curr_nmix].append(L)
    if print_progress:
      print("#mix:%.2d #iter:%.2d llk:%.4f" % (curr_nmix, len(self._llk_hist[curr_nmix]), L))
    self._checkpoint()

  # ------------------------------------------------------ scoring utilities
  def _score_device(self, X, want_post, want_logprob):
    torch = _torch()
    self.initialize(X)
    lib = self._upload_params()
    frames = self._frames(X)
    M = self._curr_nmix
    llk = np.empty((frames.n, 1), dtype=np.float32)
    post = np.empty((frames.n, M), dtype=np.float32) if want_post else None
    logp = np.empty((frames.n, M), dtype=np.float32) if want_logprob else None
    for dev, s, e in frames.chunks():
      n = e - s
      d_llk = torch.empty(n, dtype=torch.float32, device="cuda")
      d_post = torch.empty((n, M), dtype=torch.float32, device="cuda") if want_post else None
      d_logp = torch.empty((n, M), dtype=torch.float32, device="cuda") if want_logprob else None
      _lib.check(lib.odin_gmm_score(self._handle, _lib.ptr(dev), n, _lib.ptr(d_llk), _lib.ptr(d_post),
                                    _lib.ptr(d_logp), _lib.current_stream()))
      llk[s:e, 0] = d_llk.cpu().numpy()
      if want_post:
        post[s:e] = d_post.cpu().numpy()
      if want_logprob:
        logp[s:e] = d_logp.cpu().numpy()
    c = np.float32(self._llk_pad_const())
    if c != 0:
      llk += c
      if want_logprob:
        logp += c
    return llk, post, logp

  def logprob(self, X):
    """gmm_tmat.py:916-938 -> [batch, nmix]."""
    return self._score_device(X, False, True)[2]

  def postprob(self, X, gpu='auto'):
    """gmm_tmat.py:940-966 -> [batch, nmix]."""
    return self._score_device(X, True, False)[1]

  def llk(self, X, gpu='auto'):
    """gmm_tmat.py:968-995 -> [batch, 1]."""
    return self._score_device(X, False, False)[0]

  def score(self, X, y=None):
    """gmm_tmat.py:701-706 -> [batch, 1]."""
    return self.llk(X)

  # ------------------------------------------------ per-utterance statistics
  def _utt_stats_device(self, frames_dev, d_sad, offsets):
    """offsets: int64 numpy [n_utt+1] into frames_dev. Returns (Z, Fhat) CUDA tensors."""
    torch = _torch()
    lib = self._upload_params()
    n_utt = len(offsets) - 1
    D, K, M = self._feat_dim, self._kdim, self._curr_nmix
    if frames_dev.shape[1] != K:
      frames_dev = self._pad_frames(frames_dev)
    Z = torch.empty((n_utt, M), dtype=torch.float32, device="cuda")
    Fh = torch.empty((n_utt, M * K), dtype=torch.float32, device="cuda")
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    _lib.check(lib.odin_gmm_utt_stats(self._handle, _lib.ptr(frames_dev), _lib.ptr(d_sad), _lib.as_i64_ptr(off),
                                      n_utt, _lib.ptr(Z), _lib.ptr(Fh), self.impl, _lib.current_stream()))
    if K != D:   # drop the pad columns: index m * D + d (gmm_tmat.py:754-757)
      Fh = Fh.view(n_utt, M, K)[:, :, :D].reshape(n_utt, M * D)
    return Z, Fh

  def transform(self, X, zero=True, first=True, device=None):
    """gmm_tmat.py:708-767 -> (Z [1,M], Fhat [1, D*M]), Fhat index = m*D + d."""
    zero, first = bool(zero), bool(first)
    if not zero and not first:
      raise ValueError("One of `zero` or `first` must be True")
    self.initialize(X)
    assert X.ndim == 2 and X.shape[1] == self.feat_dim, \
        "`X` must be 2-D matrix, with `X.shape[1]=%d`; but given: %s" % (self.feat_dim, str(X.shape))
    frames = self._frames(X)
    frames.cache_on_device(0)
    Z, Fh = self._utt_stats_device(frames.resident, None, np.array([0, frames.n], dtype=np.int64))
    Z, Fh = Z.cpu().numpy(), Fh.cpu().numpy()
    if zero and first:
      return Z, Fh
    return Z if zero else Fh

  def transform_to_disk(self, X, indices, sad=None, pathZ=None, pathF=None, name_path=None,
                        dtype='float32', device='gpu', ncpu=None, override=True, utt_batch=2048):
    """gmm_tmat.py:769-913.  Z [n_utt, M] and Fhat [n_utt, M*D] are written as
    .npy files (the reference's bigarray.MmapArray container is a third-party
    format outside this path); returns the utterance names in processing order
    (sorted by start, like the reference).

    Multi-GPU: the utterances are dealt to the ranks (`sharding.shard_utterances`), every rank writes the rows
    of its own utterances into the shared files (rank 0 creates them) -- no collective on the data path
    (SURVEY 8e); `local_shard=True` keeps everything on the calling rank."""
    if isinstance(indices, Mapping):
      indices = list(indices.items())
    indices = sorted(indices, key=lambda x: x[1][0])
    self.initialize(X)
    torch = _torch()
    n_utt = len(indices)
    td = None if self.local_shard else _dist()
    rank = td.get_rank() if td is not None else 0
    mine = list(range(n_utt)) if td is None else \
        sharding.shard_utterances([int(e) - int(s) for _, (s, e) in indices], td.get_world_size())[rank]
    frames = self._frames(X)
    resident = frames.cache_on_device() if td is None else frames.resident is not None
    D, M = self._feat_dim, self._curr_nmix   # (a partially fitted model emits _curr_nmix columns)
    z_dat = f_dat = None
    if rank == 0:
      for p in (pathZ, pathF):
        if p is not None and os.path.exists(p) and override:
          os.remove(p)
      if pathZ is not None:
        z_dat = np.lib.format.open_memmap(pathZ, mode='w+', dtype=np.dtype(dtype), shape=(n_utt, M))
      if pathF is not None:
        f_dat = np.lib.format.open_memmap(pathF, mode='w+', dtype=np.dtype(dtype), shape=(n_utt, M * D))
    if td is not None:
      td.barrier()
      if rank != 0:
        z_dat = np.load(pathZ, mmap_mode='r+') if pathZ is not None else None
        f_dat = np.load(pathF, mmap_mode='r+') if pathF is not None else None
    d_sad_all = None
    if sad is not None:
      assert sad.shape[0] == frames.n
      sad = (np.asarray(sad).reshape(-1) != 0).astype(np.uint8)
      if resident:
        d_sad_all = torch.from_numpy(sad).cuda()
    Zs, Fs = [], []
    utt_batch = max(1, min(int(utt_batch), (1 << 31) // (4 * M * (self._kdim + 1))))   # <= 2 GiB of statistics per launch
    for b0 in range(0, len(mine), utt_batch):
      rows = mine[b0:b0 + utt_batch]
      spans = [(int(indices[i][1][0]), int(indices[i][1][1])) for i in rows]
      lens = np.array([e - s for s, e in spans], dtype=np.int64)
      off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
      contiguous = all(spans[k][1] == spans[k + 1][0] for k in range(len(spans) - 1))
      lo, hi = spans[0][0], spans[-1][1]
      d_sad = None
      if resident:   # one ragged batch per launch: the utterances of the batch back to back
        dev = frames.resident[lo:hi] if contiguous else torch.cat([frames.resident[s:e] for s, e in spans], 0)
        if sad is not None:
          d_sad = d_sad_all[lo:hi] if contiguous else torch.cat([d_sad_all[s:e] for s, e in spans], 0)
      else:
        host = frames.host[lo:hi] if contiguous else np.concatenate([frames.host[s:e] for s, e in spans], 0)
        dev = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float32)).cuda()
        if sad is not None:
          d_sad = torch.from_numpy(np.ascontiguousarray(
              sad[lo:hi] if contiguous else np.concatenate([sad[s:e] for s, e in spans]))).cuda()
      Z, Fh = self._utt_stats_device(dev, d_sad, off)
      Z, Fh = Z.cpu().numpy(), Fh.cpu().numpy()
      if z_dat is not None:
        z_dat[rows] = Z
      if f_dat is not None:
        f_dat[rows] = Fh
      if z_dat is None and f_dat is None:
        Zs.append(Z)
        Fs.append(Fh)
    for d in (z_dat, f_dat):
      if d is not None:
        d.flush()
    if td is not None:
      td.barrier()
    names = [n for n, _ in indices]
    if isinstance(name_path, str) and rank == 0:
      np.savetxt(fname=name_path, X=names, fmt='%s')
    if z_dat is None and f_dat is None:
      Zl = np.concatenate(Zs, 0) if Zs else np.zeros((0, M), np.float32)
      Fl = np.concatenate(Fs, 0) if Fs else np.zeros((0, M * D), np.float32)
      self.last_utt_stats_ = (Zl, Fl) if td is None else \
          (sharding.gather_rows(Zl, mine, n_utt), sharding.gather_rows(Fl, mine, n_utt))
    return names
