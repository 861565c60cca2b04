"""``odin.ml.Ivector`` (odin/ml/ivector.py:83-520): the GMM-UBM + T-matrix pair behind one ``fit`` /
``transform`` with its on-disk layout (``gmm.pkl``, ``tmat.pkl``, ``zstat_*``, ``fstat_*``, ``ivec_*``,
``name_*`` under ``path``).  Host orchestration only: every arithmetic step is a call into
``odin_b200.ml.GMM`` / ``odin_b200.ml.Tmatrix`` (CUDA kernels behind the C-ABI).

Differences from the reference, both forced by formats outside this path: statistics and i-vectors are
``.npy`` files (the reference's ``MmapArray`` container is the third-party ``bigarray`` format), and the
T-matrix is trained in float64 whatever ``dtype`` says (the reference passes ``dtype`` -- float32 by
default -- to ``Tmatrix``, whose own default and recommendation is float64, gmm_tmat.py:1356-1361)."""
import os
import pickle
import uuid as _uuid

import numpy as np

from .gmm import GMM
from .tmat import Tmatrix


class Ivector(object):

  def __init__(self, path, nmix=None, tv_dim=None, nmix_start=1, niter_gmm=16, niter_tmat=16, allow_rollback=True,
               exit_on_error=False, downsample=1, stochastic_downsample=True, device='gpu', ncpu=1,
               gpu_factor_gmm=80, gpu_factor_tmat=3, dtype='float32', seed=1234, name=None):
    for key, val in list(locals().items()):
      if key in ('self', 'path', 'seed'):
        continue
      setattr(self, key, val)
    self._rand = np.random.RandomState(seed=seed)
    path = str(path)
    if not os.path.exists(path):
      os.mkdir(path)
    elif not os.path.isdir(path):
      raise ValueError("Path to '%s' is not a directory" % str(path))
    self._path = path
    self._gmm = None
    self._tmat = None

  # ---- models (ivector.py:123-172) -------------------------------------------
  @property
  def gmm(self):
    if self._gmm is None:
      if os.path.exists(self.gmm_path):
        with open(self.gmm_path, 'rb') as f:
          self._gmm = pickle.load(f)
        assert self._gmm.nmix == self.nmix, \
            "Require GMM with %d components, but found %s, at path: '%s'" % (self.nmix, str(self._gmm), self.gmm_path)
      else:
        self._gmm = GMM(nmix=self.nmix, nmix_start=self.nmix_start, niter=self.niter_gmm, dtype=self.dtype,
                        allow_rollback=self.allow_rollback, exit_on_error=self.exit_on_error,
                        downsample=self.downsample, stochastic_downsample=self.stochastic_downsample,
                        device=self.device, ncpu=self.ncpu, gpu_factor=self.gpu_factor_gmm, seed=1234,
                        path=self.gmm_path,
                        name="IvecGMM_%s" % (self.name if self.name is not None else str(self._rand.randint(10e8))))
    return self._gmm

  @property
  def tmat(self):
    if self._tmat is None:
      if os.path.exists(self.tmat_path):
        with open(self.tmat_path, 'rb') as f:
          self._tmat = pickle.load(f)
        assert self._tmat.tv_dim == self.tv_dim, \
            "Require T-matrix with %d dimensions, but found %s, at path: '%s'" % \
            (self.tv_dim, str(self._tmat), self.tmat_path)
      else:
        self._tmat = Tmatrix(tv_dim=self.tv_dim, gmm=self.gmm, niter=self.niter_tmat, dtype='float64',
                             device=self.device, ncpu=self.ncpu, gpu_factor=self.gpu_factor_tmat, cache_path='/tmp',
                             seed=1234, path=self.tmat_path,
                             name='IvecTmat_%s' % (self.name if self.name is not None else str(self._rand.randint(10e8))))
    return self._tmat

  path = property(lambda self: self._path)
  gmm_path = property(lambda self: os.path.join(self.path, 'gmm.pkl'))
  tmat_path = property(lambda self: os.path.join(self.path, 'tmat.pkl'))
  z_path = property(lambda self: os.path.join(self.path, 'zstat_train.npy'))
  f_path = property(lambda self: os.path.join(self.path, 'fstat_train.npy'))
  ivec_path = property(lambda self: os.path.join(self.path, 'ivec_train.npy'))
  name_path = property(lambda self: os.path.join(self.path, 'name_train'))
  feat_dim = property(lambda self: self.gmm.feat_dim)
  is_gmm_fitted = property(lambda self: self.gmm.is_fitted)
  is_tmat_fitted = property(lambda self: self.is_gmm_fitted and self.tmat.is_fitted)
  is_fitted = property(lambda self: self.is_gmm_fitted and self.is_tmat_fitted)

  def get_z_path(self, name=None):
    return self.z_path if name is None else os.path.join(self.path, 'zstat_%s.npy' % name)

  def get_f_path(self, name=None):
    return self.f_path if name is None else os.path.join(self.path, 'fstat_%s.npy' % name)

  def get_i_path(self, name=None):
    return self.ivec_path if name is None else os.path.join(self.path, 'ivec_%s.npy' % name)

  def get_name_path(self, name=None):
    return self.name_path if name is None else os.path.join(self.path, 'name_%s' % name)

  # ---- statistics (ivector.py:18-78) -----------------------------------------
  def _extract_stats(self, X, sad, indices, z_path, f_path, name_path):
    gmm = self.gmm
    if indices is None:   # every row is one sample
      n = X.shape[0]
      idx = [(str(i), (i, i + 1)) for i in range(n)]
      gmm.transform_to_disk(X, indices=idx, sad=None, pathZ=z_path, pathF=f_path, dtype='float32', override=True)
      if sad is not None:   # rows removed by the SAD keep all-zero statistics (ivector.py:52-64)
        keep = np.asarray(sad).astype(bool).ravel()
        for p in (z_path, f_path):
          a = np.load(p, mmap_mode='r+')
          a[~keep] = 0
          a.flush()
    else:
      gmm.transform_to_disk(X, indices=indices, sad=sad, pathZ=z_path, pathF=f_path, name_path=name_path,
                            dtype='float32', override=True)

  # ---- sklearn surface (ivector.py:260-470) ------------------------------------
  def fit(self, X, indices=None, sad=None, refit_gmm=False, refit_tmat=False, extract_ivecs=False, keep_stats=False):
    new_gmm = (not self.gmm.is_fitted or refit_gmm)
    if new_gmm:
      data = [X]
      if sad is not None:
        data.append(sad)
      if indices is not None:
        data.append(indices)
      self.gmm.fit(data)
      if self.gmm.path is not None and not os.path.exists(self.gmm.path):
        with open(self.gmm.path, 'wb') as f:
          pickle.dump(self.gmm, f)
    new_tmat = (not self.tmat.is_fitted or new_gmm or refit_tmat)
    new_ivec = extract_ivecs and (new_tmat or not os.path.exists(self.ivec_path))
    if not new_gmm and os.path.exists(self.z_path) and os.path.exists(self.f_path):
      new_stats = False
    else:
      new_stats = new_gmm or new_tmat or new_ivec
    if new_stats:
      self._extract_stats(X, sad, indices, self.z_path, self.f_path, self.name_path)
    if new_tmat or new_ivec:
      Z, F = np.load(self.z_path, mmap_mode='r'), np.load(self.f_path, mmap_mode='r')
      if new_tmat:
        self.tmat.fit((Z, F))
      if new_ivec:
        self.tmat.transform_to_disk(path=self.ivec_path, Z=Z, F=F, dtype='float32', override=True)
      del Z, F
    if not keep_stats:
      for p in (self.z_path, self.f_path):
        if os.path.exists(p):
          os.remove(p)
    return self

  def transform(self, X, indices=None, sad=None, save_ivecs=False, keep_stats=False, name=None):
    if not self.is_fitted:
      raise ValueError("Ivector has not been fitted, call Ivector.fit(...) first")
    n_files = X.shape[0] if indices is None else len(indices)
    name = _uuid.uuid4().hex[:8] if name is None else str(name)
    z_path, f_path = self.get_z_path(name), self.get_f_path(name)
    i_path = self.get_i_path(name) if save_ivecs else None
    name_path = self.get_name_path(name)
    if i_path is not None and os.path.exists(i_path):
      ivec = np.load(i_path, mmap_mode='r')
      assert ivec.shape[0] == n_files and ivec.shape[1] == self.tv_dim, \
          "Need i-vectors for %d files, found exists data at path:'%s' with shape:%s" % (n_files, i_path, ivec.shape)
      return ivec
    if not (os.path.exists(z_path) and os.path.exists(f_path)):
      for p in (z_path, f_path, name_path):
        if os.path.exists(p):
          os.remove(p)
      self._extract_stats(X, sad, indices, z_path, f_path, name_path)
    Z, F = np.load(z_path, mmap_mode='r'), np.load(f_path, mmap_mode='r')
    ivec = self.tmat.transform_to_disk(path=i_path, Z=Z, F=F, dtype='float32')
    del Z, F
    if not keep_stats:
      for p in (z_path, f_path):
        if os.path.exists(p):
          os.remove(p)
    return ivec

  def __str__(self):
    return "<Ivector GMM:%s Tmat:%s path:'%s' nmix:%s tv_dim:%s>" % (
        self.is_gmm_fitted, self.is_tmat_fitted, self.path, self.nmix, self.tv_dim)
