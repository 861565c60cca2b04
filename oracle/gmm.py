"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GMM-UBM Baum-Welch path.

numpy restatement of ``odin/ml/gmm_tmat.py`` (reference = trungnt13/odin-ai,
numpy branch; the TF1 "GPU" branch is dead code at HEAD).
Citations are ``gmm_tmat.py:line``.

Pinning: the reference holds no tests / golden vectors for this path, so the
oracle is pinned against the reference itself executed here under
``oracle/ref_shim.py`` (tests/test_oracle_vs_reference.py) and through the
committed fixtures tests/golden/gmm_*.npz (oracle/make_golden.py), plus the
RNG-free self-check values of SURVEY.md Appendix B.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module; the product package never does.

Layout follows the reference: mean [D,M], sigma [D,M] (= VARIANCE despite the
name, :496), w [1,M]; statistics Z [1,M], F [D,M], S [D,M].
"""
import numpy as np

EPS = 1e-6  # gmm_tmat.py:27

# split-and-train schedule, gmm_tmat.py:677
NITER_SCHEDULE = [1, 2, 4, 4, 4, 4, 6, 6, 10, 10, 10, 10, 10, 16, 16]


def posterior_constants(mean, sigma, w):
  """gmm_tmat.py:493-504 -> (precision [D,M], mu_precision [D,M], C [1,M])."""
  precision = 1 / (sigma + EPS)
  C = (np.sum(mean**2 * precision, axis=0, keepdims=True) +
       np.sum(np.log(sigma + EPS), axis=0, keepdims=True) -
       2 * np.log(w + EPS))
  return precision, mean * precision, C


def _lse(a):
  m = np.max(a, axis=1, keepdims=True)
  return m + np.log(np.sum(np.exp(a - m), axis=1, keepdims=True))  # :78-99


def estep(X, mean, sigma, w, second=True, compute_dtype=None):
  """gmm_tmat.py:1012-1041 -> (Z [1,M], F [D,M], S [D,M] | None, L = sum LLK).

  compute_dtype=None keeps numpy-2 behaviour of the reference as it runs today
  (float32 X, float32 params, np.float64 `_feat_const` => promoted to float64
  from the first addition on); compute_dtype=np.float32 reproduces the
  numpy-1.x era (everything float32); np.float64 casts X and the parameters
  first (strictly more accurate than either; the default checker)."""
  if compute_dtype is not None:
    X = X.astype(compute_dtype, copy=False)
    mean, sigma, w = (a.astype(compute_dtype, copy=False) for a in (mean, sigma, w))
  precision, mu_precision, C = posterior_constants(mean, sigma, w)
  D = X.shape[1]
  feat_const = D * np.log(2 * np.pi)  # :600 (np.float64 scalar)
  if compute_dtype is not None and np.dtype(compute_dtype) == np.float32:
    feat_const = float(feat_const)  # weak scalar: stays float32
  X2 = X**2
  dist = np.dot(X2, precision) - 2 * np.dot(X, mu_precision) + feat_const
  logprob = -0.5 * (C + dist)
  llk = _lse(logprob)
  post = np.exp(logprob - llk)
  Z = np.sum(post, axis=0, keepdims=True)
  F = np.dot(X.T, post)
  S = np.dot(X2.T, post) if second else None
  return Z, F, S, np.sum(llk, axis=None)


def minibatch_ranges(n, batch_size):
  """odin.utils.minibatch (utils/__init__.py:191-231): contiguous ranges."""
  batch_size = int(batch_size)
  return [(s, min(s + batch_size, n)) for s in range(0, n, batch_size)]


def default_batch_size(feat_dim, nmix, itemsize=4, budget=12 * 1024 * 1024):
  """gmm_tmat.py:338,602-607,1124-1126: 12 MB of input per batch, divided by
  floor(2^(nmix/1024))."""
  bs = int(budget / (feat_dim * itemsize))
  return int(bs / np.floor(np.power(2, nmix / 1024)))


def expectation(X, mean, sigma, w, sad=None, batch_size=None, second=True,
                compute_dtype=None):
  """gmm_tmat.py:1043-1231 with ncpu=1, downsample=1: sum of per-batch stats,
  L returned as the mean log-likelihood per (selected) frame."""
  n = X.shape[0]
  if batch_size is None:
    batch_size = n
  Z = F = S = 0.0
  L = 0.0
  nfr = 0
  for s, e in minibatch_ranges(n, batch_size):
    xb = X[s:e]
    if sad is not None:
      xb = xb[np.asarray(sad[s:e]).ravel().astype(bool)]
    if xb.shape[0] == 0:
      continue
    z, f, s2, l = estep(xb, mean, sigma, w, second, compute_dtype)
    Z, F, L = Z + z, F + f, L + l
    if second:
      S = S + s2
    nfr += xb.shape[0]
  return Z, F, (S if second else None), (L / nfr if nfr > 0 else 0.0), nfr


def maximization(Z, F, S, prev, allow_rollback=True):
  """gmm_tmat.py:1233-1276 -> (mean, sigma, w, rolled_back)."""
  iN = 1.0 / (Z + EPS)
  w = Z / Z.sum()
  mean = F * iN
  sigma = S * iN - mean**2
  rolled = False
  if np.any(sigma < 0.0):
    if allow_rollback:
      mean, sigma, w = prev
      rolled = True
    else:
      sigma = np.clip(sigma, 0.0, np.inf)
  return mean, sigma, w, rolled


def mixup(mean, sigma, w, nmix_target):
  """gmm_tmat.py:1308-1338: split every component along its max-variance
  dimension by +-0.55*sqrt(var)."""
  D, M = sigma.shape
  arg = sigma.argmax(0)
  eps = np.zeros((D, M), dtype="f")
  eps[arg, np.arange(M)] = np.sqrt(sigma.max(0))
  p = 0.55 * eps
  mean = np.c_[mean - p, mean + p]
  sigma = np.c_[sigma, sigma]
  w = 0.5 * np.c_[w, w]
  if 2 * M > nmix_target:
    mean, sigma, w = mean[:, :nmix_target], sigma[:, :nmix_target], w[:, :nmix_target]
  return mean, sigma, w


def fit(X, nmix, niter=16, nmix_start=1, sad=None, batch_size=None,
        dtype=np.float32, compute_dtype=None, allow_rollback=True):
  """gmm_tmat.py:625-699 + 1278-1306.  Returns (mean, sigma, w, llk_hist)."""
  D = X.shape[1]
  cur = int(np.clip(int(nmix_start), 1, nmix))
  mean = np.zeros((D, cur), dtype=dtype)   # :612-616
  sigma = np.ones((D, cur), dtype=dtype)
  w = np.ones((1, cur), dtype=dtype)
  sched = list(NITER_SCHEDULE)
  sched[int(np.log2(nmix))] = int(niter)
  hist = {}
  while True:
    for _ in range(sched[int(np.log2(cur))]):
      Z, F, S, L, _n = expectation(X, mean, sigma, w, sad, batch_size, True,
                                   compute_dtype)
      mean, sigma, w, _ = maximization(Z, F, S, (mean, sigma, w), allow_rollback)
      hist.setdefault(cur, []).append(float(L))
    if cur < nmix:
      mean, sigma, w = mixup(mean, sigma, w, nmix)
      cur = min(2 * cur, nmix)
    else:
      break
  return mean, sigma, w, hist


def transform(X, mean, sigma, w, compute_dtype=None):
  """gmm_tmat.py:708-767: (Z [1,M], Fhat [1, D*M]) with Fhat = F - mean*Z
  flattened column-major => index m*D + d."""
  Z, F, _, _ = estep(X, mean, sigma, w, second=False, compute_dtype=compute_dtype)
  D, M = mean.shape
  Fhat = np.reshape(F - mean * Z, (1, D * M), order="F")
  return Z, Fhat


def utterance_stats(X, indices, mean, sigma, w, sad=None, compute_dtype=None):
  """gmm_tmat.py:769-913 (transform_to_disk) without the file writer:
  one (Z, Fhat) row per utterance in `indices` = [(name, (start, end)), ...]
  sorted by start."""
  D, M = mean.shape
  names, Zs, Fs = [], [], []
  for name, (s, e) in sorted(indices, key=lambda kv: kv[1][0]):
    x = X[s:e]
    if sad is not None:
      x = x[np.asarray(sad[s:e]).ravel().astype(bool)]
    Z, Fh = transform(x, mean, sigma, w, compute_dtype)
    names.append(name)
    Zs.append(Z)
    Fs.append(Fh)
  return names, np.concatenate(Zs, 0), np.concatenate(Fs, 0)
