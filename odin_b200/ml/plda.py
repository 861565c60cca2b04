"""Probabilistic LDA back-end with the surface of ``odin.ml.PLDA`` (reference: odin/ml/plda.py:26-423).

SURVEY 8f-4.  The model works on a few thousand i-vectors of <= 600 dimensions: the EM iteration is two
[n_phi x n_phi] inverses per distinct class size, one [n_classes, feat_dim] x [feat_dim, n_phi] product and two
linear solves; scoring is two small GEMMs.  That is a fraction of a second of LAPACK on the host and nothing the device
would speed up, so -- unlike everything upstream of the i-vectors -- it runs in numpy / scipy float64, in the
reference's operation order (same `solve` / `inv` / `svd` calls, same random initialisation from
`RandomState(random_state)`), which keeps scores comparable to the reference's to rounding.  Measured at the FSDD
recipe's scale (2 400 x 64-dim i-vectors, n_phi 32, 12 iterations, 8 host cores): 0.25 s for fit + scoring, once per
experiment.
"""
import warnings
from numbers import Number

import numpy as np
from scipy.linalg import cholesky, inv, solve, svd

from .scoring import VectorNormalizer, compute_class_avg, compute_within_cov


def logdet(A):
  """plda.py:21-24."""
  u = cholesky(A)
  return 2 * np.log(np.diag(u)).sum()


def _unique_keep_order(seq):
  seen = set()
  return [x for x in seq if x not in seen and not seen.add(x)]


def _as_array(X):
  return np.asarray(X) if isinstance(X, (tuple, list)) else X


class PLDA(object):
  """plda.py:26-423: simplified Gaussian PLDA (eigenvoice subspace `Phi_` + full residual covariance `Sigma_`).

  Parameters follow plda.py:75-79; attributes `Sigma_, Phi_, Sb_, St_, Lambda_, Uk_, Q_hat_, X_model_`."""

  def __init__(self, n_phi=None, centering=True, wccn=True, unit_length=True, n_iter='auto', improve_threshold=1e-1,
               labels=None, dtype='float64', random_state=None, verbose=0):
    self.n_phi_ = int(n_phi) if n_phi is not None else None
    if isinstance(n_iter, str):
      n_iter = n_iter.lower()
      assert n_iter == 'auto', 'Invalid `n_iter` value: %s' % n_iter
    elif isinstance(n_iter, Number):
      assert n_iter > 0, "`n_iter` must greater than 0, but given: %d" % n_iter
    self.n_iter_ = n_iter
    self.improve_threshold_ = float(improve_threshold)
    self.feat_dim_ = None
    self._labels = labels
    self.verbose_ = int(verbose)
    self._normalizer = VectorNormalizer(centering=centering, wccn=wccn, unit_length=unit_length, lda=False, concat=False)
    self._dtype = np.dtype(dtype)
    if random_state is None:
      self._rand_state = np.random.RandomState(None)
    elif isinstance(random_state, Number):
      self._rand_state = np.random.RandomState(seed=random_state)
    elif isinstance(random_state, np.random.RandomState):
      self._rand_state = random_state
    else:
      raise ValueError("Invalid argument for `random_state`: %s" % str(random_state))
    self.Sigma_ = self.Phi_ = self.Sb_ = self.St_ = None

  dtype = property(lambda self: self._dtype)
  feat_dim = property(lambda self: self.feat_dim_)
  normalizer = property(lambda self: self._normalizer)
  labels = property(lambda self: self._labels)
  num_classes = property(lambda self: len(self._labels))

  @property
  def is_fitted(self):
    return all(hasattr(self, n) for n in ('Lambda_', 'Uk_', 'Q_hat_', 'X_model_'))

  def get_params(self, deep=True):
    return dict(n_phi=self.n_phi_, n_iter=self.n_iter_, improve_threshold=self.improve_threshold_, labels=self._labels,
                dtype=self._dtype, verbose=self.verbose_)

  def __getstate__(self):  # plda.py:147-153 (same 16-tuple)
    if not self.is_fitted:
      raise RuntimeError("The PLDA have not been fitted, nothing to pickle!")
    return (self.n_phi_, self.n_iter_, self.feat_dim_, self._labels, self.verbose_, self._normalizer, self._dtype,
            self._rand_state, self.Sigma_, self.Phi_, self.Sb_, self.St_, self.Lambda_, self.Uk_, self.Q_hat_,
            self.X_model_)

  def __setstate__(self, states):
    (self.n_phi_, self.n_iter_, self.feat_dim_, self._labels, self.verbose_, self._normalizer, self._dtype,
     self._rand_state, self.Sigma_, self.Phi_, self.Sb_, self.St_, self.Lambda_, self.Uk_, self.Q_hat_,
     self.X_model_) = states
    self.improve_threshold_ = 1e-1

  # ---- helpers ----------------------------------------------------------------------------------------------
  def initialize(self, X, labels):
    """plda.py:162-200: random full `Sigma_` and a length-normalised random `Phi_` from `random_state`."""
    feat_dim = X.shape[1]
    if self.feat_dim is None:
      self.feat_dim_ = int(feat_dim)
      if self._labels is None:
        self._labels = labels
      if self.feat_dim <= self.n_phi_:
        raise RuntimeError("`feat_dim=%d` must be greater than `n_phi=%d`" % (self.feat_dim, self.n_phi_))
      self.Sigma_ = (1. / self.feat_dim * np.eye(self.feat_dim) +
                     self._rand_state.randn(self.feat_dim, self.feat_dim)).astype(self.dtype)
      self.Phi_ = self.normalizer.transform(self._rand_state.randn(self.n_phi_, self.feat_dim)).T.astype(self.dtype)
      self.Sb_ = np.zeros((self.feat_dim, self.feat_dim), dtype=self.dtype)
      self.St_ = np.zeros((self.feat_dim, self.feat_dim), dtype=self.dtype)
    if self.feat_dim != feat_dim:
      raise ValueError("Mismatch the input feature dimension, %d != %d" % (self.feat_dim, feat_dim))
    if self.num_classes != len(labels):
      raise ValueError("Mismatch the number of output classes, %d != %d" % (self.num_classes, len(labels)))

  def _update_caches(self):
    """plda.py:203-212: the matrices of the two-covariance scoring rule."""
    iSt = inv(self.St_)
    iS = inv(self.St_ - np.dot(np.dot(self.Sb_, iSt), self.Sb_))
    Q = iSt - iS
    P = np.dot(np.dot(iSt, self.Sb_), iS)
    U, s, _ = svd(P, full_matrices=False)
    self.Lambda_ = np.diag(s[:self.n_phi_])
    self.Uk_ = U[:, :self.n_phi_]
    self.Q_hat_ = np.dot(np.dot(self.Uk_.T, Q), self.Uk_)

  def fit_maximum_likelihood(self, X, y):
    """plda.py:214-235: closed-form two-covariance model."""
    X, y = _as_array(X), _as_array(y)
    X = self.normalizer.fit(X, y).transform(X)
    classes = np.unique(y)
    self.initialize(X, labels=classes)
    Sw = compute_within_cov(X, y, classes)
    self.St_ = np.cov(X.T)
    self.Sb_ = self.St_ - Sw
    self._update_caches()
    self.X_model_ = np.dot(compute_class_avg(X, y, classes=classes), self.Uk_)
    return self

  def fit(self, X, y):
    """plda.py:237-306: EM re-estimation of the eigenvoice subspace."""
    X, y = _as_array(X), _as_array(y)
    assert X.shape[0] == y.shape[0], \
        "Number of samples mismatch in `X` and `y`, %d != %d" % (X.shape[0], y.shape[0])
    y_counts = np.bincount(y)
    classes = np.unique(y)
    X = self.normalizer.fit(X, y).transform(X)
    self.initialize(X, labels=classes)
    F = np.zeros((self.num_classes, self.feat_dim))
    for clz in np.unique(y):
      F[clz, :] = X[y == clz, :].sum(axis=0)
    X_sqr = np.dot(X.T, X)
    it, last_llk = 0, None
    while True:
      Ey, Eyy = self.expectation_plda(F, y_counts)
      self.maximization_plda(X, X_sqr, F, Ey, Eyy)
      llk = None
      if self.verbose_ > 1 or isinstance(self.n_iter_, str):
        llk = self.compute_llk(X)
      if self.verbose_ > 0:
        print('#iter:%-3d \t [llk = %s]' % (it + 1, 'None' if llk is None else '%.2f' % llk))
      it += 1
      if isinstance(self.n_iter_, Number):
        if it >= self.n_iter_:
          break
      elif it > 2 and last_llk is not None:
        if llk - last_llk < self.improve_threshold_:
          break
      last_llk = llk
    self.Sb_ = self.Phi_.dot(self.Phi_.T)
    self.St_ = self.Sb_ + self.Sigma_
    self._update_caches()
    self.X_model_ = np.dot(compute_class_avg(X, y, classes=classes), self.Uk_)
    return self

  def expectation_plda(self, F, cls_counts):
    """plda.py:308-342: posterior mean / second moment of the speaker factors; one inverse per distinct count."""
    num_classes = F.shape[0]
    uniq = _unique_keep_order(cls_counts)
    PhiT_invS = solve(self.Sigma_.T, self.Phi_).T
    PhiT_invS_Phi = np.dot(PhiT_invS, self.Phi_)
    I = np.eye(self.n_phi_)
    inv_terms = {n: inv(I + n * PhiT_invS_Phi) for n in uniq}
    Eyy = np.zeros((self.n_phi_, self.n_phi_))
    Ey = np.zeros((num_classes, self.n_phi_))
    for clz in range(num_classes):
      n = cls_counts[clz]
      Cyy = inv_terms[n]
      Ey[clz, :] = np.dot(Cyy, np.dot(PhiT_invS, F[clz, :]))
      Eyy += n * Cyy
    Eyy += np.dot((Ey * cls_counts[:, None]).T, Ey)
    return Ey, Eyy

  def compute_llk(self, X):
    """plda.py:344-356."""
    n = X.shape[0]
    S = np.dot(self.Phi_, self.Phi_.T) + self.Sigma_
    return -0.5 * (self.feat_dim * n * np.log(2 * np.pi) + n * logdet(S) + np.sum(X * solve(S, X.T).T))

  def maximization_plda(self, X, X_sqr, F, Ey, Eyy):
    """plda.py:358-375."""
    Ey_FT = np.dot(Ey.T, F)
    self.Phi_ = solve(Eyy.T, Ey_FT).T
    self.Sigma_ = 1. / X.shape[0] * (X_sqr - np.dot(self.Phi_, Ey_FT))

  def transform(self, X):
    """plda.py:377-392 -> [num_samples, n_phi]."""
    if not self.is_fitted:
      raise RuntimeError("This model hasn't been fitted!")
    return np.dot(self.normalizer.transform(_as_array(X)), self.Uk_)

  def predict_log_proba(self, X, X_model=None):
    """plda.py:394-423 -> log-likelihood-ratio scores [num_samples, num_classes]."""
    if not self.is_fitted:
      raise RuntimeError("This model hasn't been fitted!")
    if X_model is None:
      X_model = self.X_model_
    else:
      X_model = np.dot(self.normalizer.transform(X_model), self.Uk_)
    if X_model.shape[0] != self.num_classes:
      warnings.warn("The model matrix contains %d classes, but the fitted number of classes is %d" %
                    (X_model.shape[0], self.num_classes))
    X = np.dot(self.normalizer.transform(_as_array(X)), self.Uk_)
    score_h1 = np.sum(np.dot(X_model, self.Q_hat_) * X_model, axis=1, keepdims=True)
    score_h2 = np.sum(np.dot(X, self.Q_hat_) * X, axis=1, keepdims=True)
    score_h1h2 = 2 * np.dot(X, np.dot(X_model, self.Lambda_).T)
    return score_h1h2 + score_h1.T + score_h2

  def predict(self, X):
    return np.asarray(self._labels)[np.argmax(self.predict_log_proba(X), axis=-1)]
