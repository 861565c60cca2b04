"""Vector normalisation and cosine scoring of the i-vector back-end, with the surface of
``odin.ml.scoring`` (reference: odin/ml/scoring.py:15-364, helpers odin/backend/maths.py:110-135).

SURVEY 8f-4: this is <= 600-dimensional dense linear algebra on a few thousand i-vectors (one covariance, one
inverse + Cholesky, one LDA solve, one [n, d] x [d, classes] product) -- microseconds to milliseconds on the
host, with nothing for the device to win; it is evaluated with numpy / scipy in float64 in the reference's
operation order so that a recipe (examples/fsdd_ivec.py:270-330) runs to its scores on this package alone.
The heavy stages feeding it (features, UBM statistics, T-matrix, i-vectors) are the CUDA path.
"""
import numpy as np
from scipy.linalg import cholesky, inv


def length_norm(x, axis=-1, epsilon=1e-12, ord=2):
  """odin/backend/maths.py:110-130: x / max(||x||, sqrt(eps)) (ord 2) or x / max(|x|_1, eps) (ord 1)."""
  ord = int(ord)
  if ord not in (1, 2):
    raise ValueError("only support `ord`: 1 for L1-norm; 2 for Frobenius or Euclidean")
  x = np.asarray(x)
  if ord == 2:
    x_norm = np.sqrt(np.maximum(np.sum(x ** 2, axis=axis, keepdims=True), epsilon))
  else:
    x_norm = np.maximum(np.sum(np.abs(x), axis=axis, keepdims=True), epsilon)
  return x / x_norm


def calc_white_mat(X):
  """odin/backend/maths.py:133-135: lower Cholesky factor of inv(X)."""
  return cholesky(inv(X), lower=True)


def compute_class_avg(X, y, classes, sorting=True):
  """scoring.py:15-40 -> [nb_classes, feat_dim], rows in (sorted) class order."""
  if sorting:
    classes = sorted(classes, reverse=False)
  return np.concatenate([np.mean(X[y == i], axis=0, keepdims=True) for i in classes], axis=0)


def compute_within_cov(X, y, classes=None, class_avg=None):
  """scoring.py:42-69: covariance of the class-centred vectors (np.cov, ddof 1)."""
  if classes is None and class_avg is None:
    raise ValueError("`classes` and `class_avg` cannot be None together")
  if classes is not None:
    class_avg = compute_class_avg(X, y, classes, sorting=True)
  X_mu = X - class_avg[y]
  return np.cov(X_mu.T)


def compute_wccn(X, y, classes=None, class_avg=None):
  """scoring.py:71-93: W with X_norm = X @ W, from the within-class covariance + 1e-6 I."""
  if classes is None and class_avg is None:
    raise ValueError("`classes` and `class_avg` cannot be None together")
  Sw = compute_within_cov(X, y, classes, class_avg)
  Sw = Sw + 1e-6 * np.eye(Sw.shape[0])
  return calc_white_mat(Sw)


class VectorNormalizer(object):
  """scoring.py:95-252: centering -> WCCN whitening -> optional LDA -> length normalisation."""

  def __init__(self, centering=True, wccn=False, unit_length=True, lda=False, concat=False):
    self._centering = bool(centering)
    self._unit_length = bool(unit_length)
    self._wccn = bool(wccn)
    if bool(lda):
      from sklearn.discriminant_analysis import LinearDiscriminantAnalysis
      self._lda = LinearDiscriminantAnalysis()
    else:
      self._lda = None
    self._feat_dim = None
    self._concat = bool(concat)

  feat_dim = property(lambda self: self._feat_dim)
  is_initialized = property(lambda self: self._feat_dim is not None)
  is_fitted = property(lambda self: hasattr(self, '_W'))
  enroll_vecs = property(lambda self: self._enroll_vecs)
  mean = property(lambda self: self._mean)
  vmin = property(lambda self: self._vmin)
  vmax = property(lambda self: self._vmax)
  W = property(lambda self: self._W)
  lda = property(lambda self: self._lda)

  def get_params(self, deep=True):
    return dict(centering=self._centering, wccn=self._wccn, unit_length=self._unit_length,
                lda=self._lda is not None, concat=self._concat)

  def _initialize(self, X, y):
    if not self.is_initialized:
      self._feat_dim = X.shape[1]
    assert self._feat_dim == X.shape[1]
    if isinstance(y, (tuple, list)):
      y = np.asarray(y)
    if y.ndim == 2:
      y = np.argmax(y, axis=-1)
    return y, np.unique(y)

  def normalize(self, X, concat=None):
    if not self.is_fitted:
      raise RuntimeError("VectorNormalizer has not been fitted.")
    if concat is None:
      concat = self._concat
    X_org = (X[:] if not isinstance(X, np.ndarray) else X) if concat else None
    if self._centering:
      X = X - self._mean
    if self._wccn:
      X = np.dot(X, self.W)
    if self._lda is not None:
      X_lda = self._lda.transform(X)
      X = np.concatenate((X_lda, X_org), axis=-1) if concat else X_lda
    if self._unit_length:
      X = length_norm(X, axis=-1, ord=2)
    return X

  def fit(self, X, y):
    y, classes = self._initialize(X, y)
    enroll = compute_class_avg(X, y, classes, sorting=True)
    M = X.mean(axis=0).reshape(1, -1)
    self._mean = M
    if self._centering:
      X = X - M
    # (scoring.py:229-232: the class averages handed to the whitening are those of the UNcentred vectors)
    W = compute_wccn(X, y, classes=None, class_avg=enroll) if self._wccn else 1
    self._W = W
    if self._wccn:
      X = np.dot(X, W)
    if self._unit_length:
      X = length_norm(X, axis=-1)
    if self._lda is not None:
      self._lda.fit(X, y)
    self._enroll_vecs = self.normalize(enroll, concat=False)
    if self._lda is not None:
      X = self._lda.transform(X)
      X = length_norm(X, axis=-1, ord=2)
    self._vmin, self._vmax = X.min(0, keepdims=True), X.max(0, keepdims=True)
    return self

  def transform(self, X):
    return self.normalize(X)


class Scorer(object):
  """scoring.py:254-364: cosine scoring against the class (enrolment) averages, or an RBF SVM on the normalised
  vectors (sklearn, like the reference)."""

  def __init__(self, centering=True, wccn=True, lda=True, concat=False, method='cosine', labels=None):
    self._normalizer = VectorNormalizer(centering=centering, wccn=wccn, lda=lda, concat=concat)
    self._labels = labels
    method = str(method).lower()
    if method not in ('cosine', 'svm'):
      raise ValueError('`method` must be one of the following: cosine, svm; but given: "%s"' % method)
    self._method = method

  method = property(lambda self: self._method)
  feat_dim = property(lambda self: self._normalizer.feat_dim)
  labels = property(lambda self: self._labels)
  nb_classes = property(lambda self: len(self._labels))
  is_initialized = property(lambda self: self._normalizer.is_initialized)
  is_fitted = property(lambda self: self._normalizer.is_fitted)
  normalizer = property(lambda self: self._normalizer)
  lda = property(lambda self: self._normalizer.lda)

  def fit(self, X, y):
    if isinstance(X, (tuple, list)):
      X = np.asarray(X)
    if isinstance(y, (tuple, list)):
      y = np.asarray(y)
    self._normalizer.fit(X, y)
    if self._labels is None:
      if y.ndim >= 2:
        y = np.argmax(y, axis=-1)
      self._labels = np.unique(y)
    if self.method == 'svm':
      from sklearn.svm import SVC
      X = self._normalizer.transform(X)
      X = 2 * (X - self._normalizer.vmin) / (self._normalizer.vmax - self._normalizer.vmin) - 1
      self._svm = SVC(C=1, kernel='rbf', gamma='auto', coef0=1, shrinking=True, random_state=0, probability=True,
                      tol=1e-3, cache_size=1e4, class_weight='balanced')
      self._svm.fit(X, y)
      self.predict_proba = self._predict_proba
    return self

  def _predict_proba(self, X):
    if self.method != 'svm':
      raise RuntimeError("`predict_proba` only for 'svm' method")
    return self._svm.predict_proba(self._normalizer.transform(X))

  def predict_log_proba(self, X):
    return self.transform(X)

  def transform(self, X):
    X = self._normalizer.transform(X)
    if self.method == 'cosine':
      return np.dot(X, self._normalizer.enroll_vecs.T)
    X = 2 * (X - self._normalizer.vmin) / (self._normalizer.vmax - self._normalizer.vmin) - 1
    return self._svm.predict_log_proba(X)

  def predict(self, X):
    """Evaluable.evaluate's decision (odin/ml/base.py): arg-max over the class scores."""
    return np.asarray(self._labels)[np.argmax(self.predict_log_proba(X), axis=-1)]
