"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (trungnt13/odin-ai).

Imports ``odin.preprocessing`` and ``odin/ml/gmm_tmat.py`` from a reference
checkout (default ``/root/reference``) under a compatibility shim, so that the
reference's own numpy/scipy/sklearn code can be executed in this container to
(a) validate the restatement in ``oracle/frontend.py`` / ``oracle/gmm.py`` and
(b) generate the golden vectors committed under ``tests/golden/``
(``oracle/make_golden.py``).

The reference checkout does not exist on the GPU box, so nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this module.  Nothing in
the product package (``odin_b200/``) imports anything under ``oracle/``.

Every patch below works around a bit-rot bug of the reference at HEAD under
Python 3.12 / numpy 2 / sklearn 1.9 (SURVEY.md section 2.3 lists them with
file:line); none changes the arithmetic:

  * ``collections.Mapping`` & friends moved to ``collections.abc``
    (base.py:7, speech.py:46, gmm_tmat.py:13, utils/__init__.py:21)
  * ``np.bool`` / ``np.Inf`` removed from numpy (speech.py:1587, gmm_tmat.py:1268)
  * ``np.random.rand(0, 888888)`` used where ``randint`` was meant (base.py:210-211)
  * ``hz2mel`` returns a 1-element array for scalars which breaks
    ``np.linspace`` in ``mel_filters`` (signal.py:514, 784-786)
  * ``delta_x[idx]`` with ``idx`` a list of slices (signal.py:1062-1064)
  * ``bigarray`` / ``tensorflow`` / ``odin.fuel`` are not importable here
  * ``get_ngpu`` undefined, TF1 graph code dead (gmm_tmat.py:105-106, 506-560)
  * ``random.seed(np.int64)`` rejected by Python >= 3.11 (gmm_tmat.py:147)
  * ``Progbar`` reads a removed tqdm attribute (utils/progbar.py:432)
"""
import collections
import collections.abc
import importlib.util
import inspect
import os
import random as _random
import sys
import types

import numpy as np

DEFAULT_ROOT = os.environ.get("ODIN_REFERENCE_ROOT", "/root/reference")

_STATE = {}


def available(root=DEFAULT_ROOT):
  return os.path.isfile(os.path.join(root, "odin", "preprocessing", "signal.py"))


def _stub(name, **attrs):
  m = types.ModuleType(name)
  m.__dict__.update(attrs)
  sys.modules[name] = m
  parent, _, child = name.rpartition(".")
  if parent in sys.modules:
    setattr(sys.modules[parent], child, m)
  return m


class _Dummy(object):

  def __init__(self, *a, **k):
    pass


def _common(root):
  if _STATE.get("common"):
    return
  for n in ("Mapping", "MutableMapping", "Iterable", "Iterator", "Sequence",
            "Callable", "Set", "MutableSet", "Hashable", "Sized", "Container"):
    if not hasattr(collections, n):
      setattr(collections, n, getattr(collections.abc, n))
  for n, t in (("bool", bool), ("int", int), ("float", float),
               ("object", object), ("str", str), ("Inf", np.inf)):
    if n not in np.__dict__:
      setattr(np, n, t)
  if root not in sys.path:
    sys.path.insert(0, root)
  _STATE["common"] = True


def load_frontend(root=DEFAULT_ROOT):
  """Returns (pp, S): reference ``odin.preprocessing`` package and its
  ``signal`` module, patched as documented in the module docstring."""
  if "pp" in _STATE:
    return _STATE["pp"], _STATE["S"]
  if not available(root):
    raise RuntimeError("reference checkout not found at %s" % root)
  _common(root)
  _stub("bigarray", MmapArray=_Dummy, MmapArrayWriter=_Dummy,
        read_mmaparray_header=lambda *a, **k: None)
  if "tensorflow" not in sys.modules:
    _stub("tensorflow", placeholder=lambda *a, **k: None)
  _stub("odin.fuel", Dataset=_Dummy, MmapDict=_Dummy, MmapData=_Dummy)
  import odin  # noqa: F401
  import odin.preprocessing as pp
  from odin.preprocessing import signal as S
  _rand = np.random.rand
  np.random.rand = lambda *a: 0 if a == (0, 888888) else _rand(*a)
  _hz2mel = S.hz2mel
  S.hz2mel = lambda f: (float(_hz2mel(f)[0]) if np.ndim(f) == 0 else _hz2mel(f))
  src = inspect.getsource(S.delta).replace("delta_x = delta_x[idx]",
                                           "delta_x = delta_x[tuple(idx)]")
  exec(compile(src, S.__file__, "exec"), S.__dict__)
  pp.base.delta = S.delta
  _STATE["pp"], _STATE["S"] = pp, S
  return pp, S


def run_pipeline(steps, X):
  """sklearn >= 1.x refuses Pipeline.transform on an unfitted pipeline; chain
  the extractors by hand exactly as Pipeline.transform would."""
  for e in steps:
    X = e.transform(X)
  return X


class _Rnd(object):
  seed = staticmethod(lambda s=None: _random.seed(None if s is None else int(s)))
  shuffle = staticmethod(_random.shuffle)
  random = staticmethod(_random.random)


class _NoProg(object):

  def __init__(self, *a, **k):
    pass

  def add(self, *a, **k):
    return self

  def __setitem__(self, k, v):
    pass

  def add_notification(self, *a, **k):
    pass


def load_gmm(root=DEFAULT_ROOT):
  """Returns the reference ``odin.ml.gmm_tmat`` module (numpy branch only)."""
  if "G" in _STATE:
    return _STATE["G"]
  if not available(root):
    raise RuntimeError("reference checkout not found at %s" % root)
  _common(root)
  load_frontend(root)  # brings in odin.utils and the bigarray/tf stubs
  from sklearn.base import BaseEstimator, DensityMixin, TransformerMixin
  import odin.utils  # noqa: F401
  _stub("odin.backend", is_tensor=lambda *a, **k: False)
  _stub("odin.ml")
  _stub("odin.ml.base", BaseEstimator=BaseEstimator,
        TransformerMixin=TransformerMixin, DensityMixin=DensityMixin)
  spec = importlib.util.spec_from_file_location(
      "odin.ml.gmm_tmat", os.path.join(root, "odin", "ml", "gmm_tmat.py"))
  G = importlib.util.module_from_spec(spec)
  sys.modules["odin.ml.gmm_tmat"] = G
  spec.loader.exec_module(G)
  G.get_ngpu = lambda: 0
  G.GMM._refresh_gpu_posterior = lambda self: None
  G.random = _Rnd
  G.Progbar = _NoProg
  G.print = lambda *a, **k: None  # silence the per-iteration log
  _STATE["G"] = G
  return G


def make_ref_gmm(nmix, nmix_start=None, niter=16, float32_mode=False, **kw):
  """Reference GMM on the CPU/numpy branch.  ``float32_mode`` reproduces the
  numpy-1.x arithmetic of the reference's era (``_feat_const`` as a Python
  float so that numpy 2 does not promote the E-step to float64)."""
  G = load_gmm()
  g = G.GMM(nmix=nmix, nmix_start=nmix if nmix_start is None else nmix_start,
            niter=niter, device="cpu", ncpu=1, **kw)
  g._float32_mode = bool(float32_mode)
  return g


def ref_gmm_initialize(g, X):
  g.initialize(X)
  if getattr(g, "_float32_mode", False):
    g._feat_const = float(g._feat_const)
  return g


def load_scoring(root=DEFAULT_ROOT):
  """Returns (scoring, plda): the reference's ``odin/ml/scoring.py`` and ``odin/ml/plda.py`` loaded by path.

  Their two helpers from ``odin.backend`` are TensorFlow one-liners (maths.py:110-135) and TensorFlow is not
  installed here; the shim supplies the same expressions in numpy / scipy -- ``cholesky(inv(X))`` (lower) and
  ``x / sqrt(max(sum(x**2), eps))`` -- which is the only arithmetic in this loader that is not the reference's own
  code.  ``Evaluable`` (reporting only) is an empty mixin."""
  if "SC" in _STATE:
    return _STATE["SC"], _STATE["PL"]
  if not available(root):
    raise RuntimeError("reference checkout not found at %s" % root)
  load_gmm(root)   # brings in odin.utils and the odin.ml / odin.backend stubs
  from scipy.linalg import cholesky, inv
  from sklearn.base import BaseEstimator, TransformerMixin
  be = sys.modules["odin.backend"]
  be.calc_white_mat = lambda X: cholesky(inv(X), lower=True)

  def length_norm(x, axis=-1, epsilon=1e-12, ord=2):
    assert int(ord) == 2
    return x / np.sqrt(np.maximum(np.sum(x ** 2, axis=axis, keepdims=True), epsilon))

  be.length_norm = length_norm
  sys.modules["odin.ml.base"].Evaluable = type("Evaluable", (object,), {})
  sys.modules["odin.ml.base"].BaseEstimator = BaseEstimator
  sys.modules["odin.ml.base"].TransformerMixin = TransformerMixin
  mods = []
  for name in ("scoring", "plda"):
    spec = importlib.util.spec_from_file_location("odin.ml." + name, os.path.join(root, "odin", "ml", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["odin.ml." + name] = m
    spec.loader.exec_module(m)
    mods.append(m)
  _STATE["SC"], _STATE["PL"] = mods
  return mods[0], mods[1]
