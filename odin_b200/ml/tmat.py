"""``odin.ml.Tmatrix`` (odin/ml/gmm_tmat.py:1343-2090) on the CUDA kernels of
``csrc/tmat_kernels.cu``: total-variability training on Baum-Welch statistics and i-vector
extraction, fp64 like the reference's default.

Same constructor arguments, attributes (``Tm``, ``T_invS``, ``T_invS_Tt``, ``Sigma``) and methods
(``fit``, ``expectation``, ``maximization``, ``expectation_maximization``, ``transform``,
``transform_to_disk``) as the reference; the arguments that only steer the reference's CPU / TF
fan-out (``device``, ``ncpu``, ``gpu_factor``, ``batch_size_*``, ``cache_path``) are accepted and
ignored.  Every arithmetic step goes through the C-ABI (``odin_tmat_*``); there is no CPU path.

One difference is visible: ``maximization(orthogonalize=True)`` replaces ``T`` by
``diag(s) V^T`` of its SVD (gmm_tmat.py:1857-1859); the sign of each row is LAPACK's choice in
the reference and the Jacobi iteration's here, so rows of ``Tm`` (and the matching i-vector
coordinates) can differ by a sign.  Everything built from ``T^T T`` -- the likelihood, cosine and
PLDA scores -- is unaffected.

Multi-GPU: with ``torch.distributed`` initialised the files are split into contiguous rank
shards, each rank accumulates the packed statistics ``LU | RU | llk | nframes`` and ONE
all-reduce per EM iteration combines them (the M-step is replicated, deterministic).
"""
import ctypes
import os
import pickle
import uuid as _uuid
from collections.abc import Mapping

import numpy as np

from .. import _lib
from .gmm import GMM, _dist, _torch


class Tmatrix(object):
  STANDARD_CPU_BATCH_SIZE = 64 * 1024 * 1024
  STANDARD_GPU_BATCH_SIZE = 64 * 1024 * 1024

  def __init__(self, tv_dim, gmm, niter=16, dtype='float64', batch_size_cpu='auto', batch_size_gpu='auto',
               device='gpu', ncpu=1, gpu_factor=3, cache_path='/tmp', seed=1234, path=None, name=None):
    if not (isinstance(gmm, GMM) and gmm.is_initialized and gmm.is_fitted):
      raise ValueError("`gmm` must be instance of odin.ml.gmm.GMM both is_initialized and is_fitted.")
    if np.dtype(dtype) != np.float64:
      raise NotImplementedError("Tmatrix is accelerated in float64 (the reference's default) only")
    self._is_fitted = False
    self.niter = niter
    self._tv_dim = int(tv_dim)
    self._t2_dim = self._tv_dim * (self._tv_dim + 1) // 2
    self._feat_dim = gmm.feat_dim
    self._nmix = gmm.nmix
    self._gmm = gmm
    self._path = path if isinstance(path, str) else None
    self._seed = seed
    self._llk_hist = []
    self._name = 'Tmatrix_%s' % _uuid.uuid4().hex[:8] if name is None else str(name)
    if not os.path.isdir(cache_path):
      raise ValueError('`cache_path` must be a directory.')
    self.cache_path = cache_path
    self._dtype = np.dtype(dtype)
    self._device = device
    self.ncpu = int(ncpu) if ncpu is not None else 1
    self.gpu_factor = int(gpu_factor)
    # gmm_tmat.py:1452-1461 (kept for the pickle tuple; the device path sizes its own batches)
    self.batch_size_cpu = int(Tmatrix.STANDARD_CPU_BATCH_SIZE / (self.feat_dim * self.nmix * self._dtype.itemsize)) \
        if isinstance(batch_size_cpu, str) else int(batch_size_cpu)
    self.batch_size_gpu = int(Tmatrix.STANDARD_GPU_BATCH_SIZE / (self.feat_dim * self.nmix * self._dtype.itemsize)) \
        if isinstance(batch_size_gpu, str) else int(batch_size_gpu)
    # gmm_tmat.py:1466-1471 (host: the numpy generator is part of the contract)
    self.Sigma = np.array(np.asarray(gmm.sigma).reshape((1, self.feat_dim * self.nmix), order='F'), dtype=self.dtype)
    np.random.seed(self._seed)
    Tm = (np.random.randn(self.tv_dim, self.feat_dim * self.nmix) * self.Sigma.sum() * 0.001).astype(self.dtype)
    self._h = None
    self._create()
    self._upload(Tm, self.Sigma)

  # ---- handle ---------------------------------------------------------------
  def _create(self):
    _lib.require_cuda()
    lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(lib.odin_tmat_create(self._tv_dim, self._nmix, self._feat_dim, ctypes.byref(h)))
    self._h = h
    self._acc_size = int(lib.odin_tmat_acc_size(h))

  def _upload(self, Tm, Sigma=None):
    torch = _torch()
    lib = _lib.load()
    d_T = torch.as_tensor(np.ascontiguousarray(Tm, dtype=np.float64)).cuda()
    d_S = torch.as_tensor(np.ascontiguousarray(Sigma, dtype=np.float64).ravel()).cuda() if Sigma is not None else None
    _lib.check(lib.odin_tmat_set_model(self._h, _lib.ptr(d_T), _lib.ptr(d_S), _lib.current_stream()))
    torch.cuda.current_stream().synchronize()

  def _download(self, which):
    torch = _torch()
    lib = _lib.load()
    shape = {'Tm': (self.tv_dim, self.feat_dim * self.nmix), 'T_invS': (self.tv_dim, self.feat_dim * self.nmix),
             'T_invS_Tt': (self.nmix, self.t2_dim)}[which]
    out = torch.empty(shape, dtype=torch.float64, device='cuda')
    args = [None, None, None]
    args[['Tm', 'T_invS', 'T_invS_Tt'].index(which)] = _lib.ptr(out)
    _lib.check(lib.odin_tmat_get_model(self._h, args[0], args[1], args[2], _lib.current_stream()))
    return out.cpu().numpy()

  def __del__(self):
    try:
      if self._h is not None:
        _lib.load().odin_tmat_destroy(self._h)
        self._h = None
    except Exception:
      pass

  Im = property(lambda self: np.eye(self.tv_dim, dtype=self.dtype))   # gmm_tmat.py:1463

  def __getstate__(self):
    """The reference's 21-tuple (gmm_tmat.py:1483-1499), so a tmat.pkl moves between the two in either direction
    (given a GMM class importable under the pickled name)."""
    return (self.Im, self.Sigma, self.Tm, self._gmm,
            self._tv_dim, self._t2_dim, self._feat_dim, self._nmix,
            self._seed, self._llk_hist,
            self.batch_size_cpu, self.batch_size_gpu,
            self.niter, self.ncpu, self._device, self.gpu_factor,
            self.cache_path, self._dtype,
            self._is_fitted, self._path, self._name)

  def __setstate__(self, s):
    (_, self.Sigma, Tm, self._gmm,
     self._tv_dim, self._t2_dim, self._feat_dim, self._nmix,
     self._seed, self._llk_hist,
     self.batch_size_cpu, self.batch_size_gpu,
     self.niter, self.ncpu, self._device, self.gpu_factor,
     self.cache_path, self._dtype,
     self._is_fitted, self._path, self._name) = s
    self._h = None
    self._create()
    self._upload(Tm, self.Sigma)

  # ---- properties (gmm_tmat.py:1538-1576) -----------------------------------
  device = property(lambda self: self._device)
  feat_dim = property(lambda self: self._feat_dim)
  tv_dim = property(lambda self: self._tv_dim)
  t2_dim = property(lambda self: self._t2_dim)
  nmix = property(lambda self: self._nmix)
  path = property(lambda self: self._path)
  name = property(lambda self: self._name)
  dtype = property(lambda self: self._dtype)
  gmm = property(lambda self: self._gmm)
  is_fitted = property(lambda self: self._is_fitted)
  Tm = property(lambda self: self._download('Tm'))
  T_invS = property(lambda self: self._download('T_invS'))
  T_invS_Tt = property(lambda self: self._download('T_invS_Tt'))

  def set_device(self, device):
    self._device = device
    return self

  # ---- statistics -----------------------------------------------------------
  def _check_ZF(self, Z, F):
    assert Z.ndim == 2 and Z.shape[1] == self.nmix, \
        "Zero-th order statistics must be 2-D matrix, and `Z.shape=[?, %d]; but given: %s" % (self.nmix, str(Z.shape))
    assert F.ndim == 2 and F.shape[1] == self.nmix * self.feat_dim, \
        "First order statistics must be 2-D matrix, and `F.shape=[?, %d]; but given: %s" % \
        (self.nmix * self.feat_dim, str(F.shape))
    assert Z.shape[0] == F.shape[0]

  def _to_device(self, A):
    torch = _torch()
    if isinstance(A, torch.Tensor):
      return A.to(device='cuda', dtype=torch.float64).contiguous()
    return torch.from_numpy(np.array(A, copy=True)).cuda().to(torch.float64)   # (memmaps may be read-only)

  def _estep_device(self, Z, F):
    """-> packed CUDA statistics LU | RU | llk | nframes, all-reduced over the ranks."""
    torch = _torch()
    lib = _lib.load()
    self._check_ZF(Z, F)
    n = Z.shape[0]
    dist = _dist()
    lo, hi = 0, n
    if dist is not None:
      from ..sharding import file_shard
      lo, hi = file_shard(n, dist.get_rank(), dist.get_world_size())
    acc = torch.zeros(self._acc_size, dtype=torch.float64, device='cuda')
    # stream the shard through the device in slabs (F is the large operand: nmix * feat_dim doubles per file)
    slab = max(1, int((1 << 30) // (8 * self.nmix * self.feat_dim)))
    for s in range(lo, hi, slab):
      e = min(hi, s + slab)
      dZ, dF = self._to_device(Z[s:e]), self._to_device(F[s:e])
      _lib.check(lib.odin_tmat_estep(self._h, _lib.ptr(dZ), _lib.ptr(dF), e - s, _lib.ptr(acc), _lib.current_stream()))
      torch.cuda.current_stream().synchronize()   # dZ / dF are released next
    if dist is not None:
      dist.all_reduce(acc)
    return acc

  def _split_acc(self, acc):
    nLU = self.nmix * self.t2_dim
    nRU = self.tv_dim * self.nmix * self.feat_dim
    LU = acc[:nLU].reshape(self.nmix, self.t2_dim)
    RU = acc[nLU:nLU + nRU].reshape(self.tv_dim, self.nmix * self.feat_dim)
    return LU, RU, acc[nLU + nRU], acc[nLU + nRU + 1]

  def expectation(self, Z, F, device=None, print_progress=True):
    """gmm_tmat.py:1727-1816 -> LU [nmix, t2], RU [tv, nmix*feat_dim], llk, nframes."""
    LU, RU, llk, nframes = self._split_acc(self._estep_device(Z, F))
    return LU.cpu().numpy(), RU.cpu().numpy(), float(llk), float(nframes)

  def _mstep_device(self, acc, min_div_est=True, orthogonalize=True):
    self._is_fitted = True
    _lib.check(_lib.load().odin_tmat_mstep(self._h, _lib.ptr(acc), 1 if min_div_est else 0, 1 if orthogonalize else 0,
                                           _lib.current_stream()))
    return self

  def maximization(self, LU, RU, nframes=None, min_div_est=True, orthogonalize=True):
    """gmm_tmat.py:1818-1865."""
    if min_div_est and nframes is None:
      raise ValueError("`nframes` must be specified if `min_div_est=True`")
    torch = _torch()
    acc = torch.zeros(self._acc_size, dtype=torch.float64, device='cuda')
    aLU, aRU, _, _ = self._split_acc(acc)
    aLU.copy_(self._to_device(LU))
    aRU.copy_(self._to_device(RU))
    acc[-1] = float(nframes) if nframes is not None else 1.0
    return self._mstep_device(acc, min_div_est, orthogonalize)

  def expectation_maximization(self, Z, F, device=None, print_progress=True):
    """gmm_tmat.py:1867-1895."""
    nfiles = Z.shape[0]
    acc = self._estep_device(Z, F)
    llk = float(acc[-2])
    self._mstep_device(acc, True, True)
    self._llk_hist.append(llk / nfiles)
    self._checkpoint()
    return self

  def _checkpoint(self):
    """Rank 0 only, through a temporary file (concurrent writers on a shared file system would corrupt it)."""
    if self.path is not None:
      td = _dist()
      if td is None or td.get_rank() == 0:
        tmp = "%s.tmp%d" % (self.path, os.getpid())
        with open(tmp, 'wb') as f:
          pickle.dump(self, f)
        os.replace(tmp, self.path)

  # ---- sklearn surface ------------------------------------------------------
  def _stats_of(self, X):
    if isinstance(X, (tuple, list)):
      Z, F = X
      self._check_ZF(np.asarray(Z) if not hasattr(Z, 'ndim') else Z, np.asarray(F) if not hasattr(F, 'ndim') else F)
      return Z, F
    return self.gmm.transform(X)

  def transform(self, X):
    """gmm_tmat.py:1898-1942: (Z, F) statistics or frames [n, feat_dim] -> i-vectors [n_files, tv_dim]."""
    torch = _torch()
    Z, F = self._stats_of(X)
    dZ, dF = self._to_device(Z), self._to_device(F)
    out = torch.empty((dZ.shape[0], self.tv_dim), dtype=torch.float64, device='cuda')
    _lib.check(_lib.load().odin_tmat_ivector(self._h, _lib.ptr(dZ), _lib.ptr(dF), dZ.shape[0], _lib.ptr(out),
                                             _lib.current_stream()))
    return out.cpu().numpy()

  def transform_to_disk(self, Z, F, path=None, dtype='float32', device='gpu', ncpu=None, override=True):
    """gmm_tmat.py:1944-2042: i-vectors of every file, written as a .npy matrix [n_files, tv_dim]
    (the reference's MmapArray container is a third-party format outside this path)."""
    self._check_ZF(Z, F)
    n = Z.shape[0]
    if path is not None and os.path.exists(path) and override:
      os.remove(path)
    dat = np.lib.format.open_memmap(path, mode='w+', dtype=np.dtype(dtype), shape=(n, self.tv_dim)) \
        if path is not None else np.empty((n, self.tv_dim), dtype=np.dtype(dtype))
    slab = max(1, int((1 << 30) // (8 * self.nmix * self.feat_dim)))
    for s in range(0, n, slab):
      e = min(n, s + slab)
      dat[s:e] = self.transform((Z[s:e], F[s:e])).astype(dat.dtype)
    if path is not None:
      dat.flush()
    return dat

  def fit(self, X, y=None):
    """gmm_tmat.py:2044-2090: X = (Z, F) statistics, or (frames, indices) to run the UBM's
    per-utterance statistics first."""
    if not isinstance(X, (tuple, list)) or len(X) != 2:
      raise ValueError("`X` must be tuple or list of length 2.")
    a, b = X
    if any(hasattr(i, 'shape') and i.shape[1] == self.feat_dim for i in X) and \
        any(isinstance(i, (tuple, list, Mapping)) for i in X):
      frames = a if hasattr(a, 'shape') else b
      indices = b if hasattr(a, 'shape') else a
      tmpZ = os.path.join(self.cache_path, 'Z_%s.npy' % _uuid.uuid4().hex[:12])
      tmpF = os.path.join(self.cache_path, 'F_%s.npy' % _uuid.uuid4().hex[:12])
      try:
        self.gmm.transform_to_disk(frames, indices, pathZ=tmpZ, pathF=tmpF, dtype='float32', override=True)
        Z, F = np.load(tmpZ), np.load(tmpF)
      finally:
        for p in (tmpZ, tmpF):
          if os.path.exists(p):
            os.remove(p)
    elif any(i.shape[1] == self.nmix for i in X) and any(i.shape[1] == self.feat_dim * self.nmix for i in X):
      Z = [i for i in X if i.shape[1] == self.nmix][0]
      F = [i for i in X if i.shape[1] == self.nmix * self.feat_dim][0]
    else:
      raise ValueError("The input arguments must be tuple of (Z, F) or (X, indices).")
    for _ in range(self.niter):
      self.expectation_maximization(Z, F)
    return self
