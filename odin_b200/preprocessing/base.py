"""Extractor contract of ``odin.preprocessing.base`` (reference: base.py:23-391).

Same dict-in / dict-out ``transform`` semantics as the reference (pass-through
of ``ExtractorSignal``, ``None`` handling, input-name checks, lower-case feature
names, merge-with-old-keys) plus the generic dictionary extractors recipes put
around the speech chain.  ``make_pipeline`` returns a ``Pipeline`` whose plan
fuses the speech steps into one batched CUDA launch sequence
(``speech.FusedSpeechFrontEnd``).
"""
import datetime
from collections.abc import Mapping

import numpy as np


def as_tuple(x, t=None):
  if x is None:
    return ()
  if isinstance(x, str) or not hasattr(x, '__iter__'):
    x = (x,)
  x = tuple(x)
  if t is not None:
    if t is int:
      x = tuple(int(i) for i in x)
    elif not all(isinstance(i, t) for i in x):
      raise ValueError("expected elements of type %s" % str(t))
  return x


class ExtractorSignal(object):
  """base.py:23-90: carries a message + action ('ignore' | 'warn' | 'error')
  down the rest of the pipeline instead of a feature dict."""

  def __init__(self):
    self._timestamp = datetime.datetime.now().strftime('%d/%m/%Y %H:%M:%S')
    self._extractor = None
    self._msg = ''
    self._action = 'ignore'
    self._last_input = {}

  message = property(lambda self: self._msg)
  action = property(lambda self: self._action)

  def set_message(self, extractor, msg, last_input):
    if self._extractor is not None:
      raise RuntimeError("This signal has stored message, cannot set message twice.")
    assert isinstance(extractor, Extractor)
    self._extractor = extractor
    self._msg = str(msg)
    self._last_input = last_input
    return self

  def set_action(self, action):
    action = str(action).lower()
    assert action in ('warn', 'error', 'ignore')
    self._action = action
    return self

  def __str__(self):
    if self._extractor is None:
      raise RuntimeError("The Signal has not been configured by the Extractor")
    s = '[%s]%s\n' % (self._timestamp, self._extractor.__class__.__name__)
    s += 'Error message: "%s"\nAction: "%s"\n' % (self._msg, self._action)
    if isinstance(self._last_input, Mapping):
      s += 'Last input keys: %s\n' % sorted(self._last_input.keys())
    else:
      s += 'Last input type: %s\n' % str(type(self._last_input))
    return s


class Extractor(object):
  """base.py:175-391.  Subclasses override ``_transform``."""

  _ID = [0]

  def __init__(self, input_name=None, output_name=None, is_input_layer=False,
               robust_level='ignore', name=None):
    if name is None:
      Extractor._ID[0] += 1
      self._name = "%s%d" % (self.__class__.__name__, Extractor._ID[0])
    else:
      self._name = str(name)
    self._debug = False
    self._is_input_layer = bool(is_input_layer)
    robust_level = str(robust_level).lower()
    assert robust_level in ('ignore', 'warn', 'error')
    self._robust_level = robust_level
    if input_name is not None and not isinstance(input_name, str):
      if not hasattr(input_name, '__iter__'):
        raise ValueError("No support for `input_name` type: %s" % str(type(input_name)))
      input_name = tuple(str(i).lower() for i in input_name)
    self._input_name = input_name
    if output_name is None:
      output_name = self.__class__.__name__.lower() if input_name is None else input_name
    elif not isinstance(output_name, str):
      if not hasattr(output_name, '__iter__'):
        raise ValueError("No support for `output_name` type: %s" % str(type(output_name)))
      output_name = tuple(str(i).lower() for i in output_name)
    self._output_name = output_name

  name = property(lambda self: self._name)
  input_name = property(lambda self: self._input_name)
  output_name = property(lambda self: self._output_name)
  is_input_layer = property(lambda self: self._is_input_layer)
  robust_level = property(lambda self: self._robust_level)

  def get_params(self, deep=True):
    return {k: v for k, v in self.__dict__.items() if not k.startswith('_')}

  def set_debug(self, debug):
    self._debug = bool(debug)
    return self

  def fit(self, X, y=None):
    return self

  def __call__(self, X):
    return self.transform(X)

  def _transform(self, X):
    raise NotImplementedError

  # pre / post halves of base.py:291-357, shared with the batched fused path
  def _check_input(self, X):
    if isinstance(X, ExtractorSignal):
      return X
    if X is None:
      return ExtractorSignal().set_message(self, "`None` value is returned by extractor",
                                           X).set_action(self.robust_level)
    if not self.is_input_layer and not isinstance(X, Mapping):
      return ExtractorSignal().set_message(
          self, "the input to `Extractor.transform` must be instance of dictionary, "
          "but given type: %s" % str(type(X)), X).set_action(self.robust_level)
    if self.input_name is not None and isinstance(X, Mapping):
      for name in as_tuple(self.input_name, t=str):
        if name not in X:
          return ExtractorSignal().set_message(self, "Cannot find features with name: %s" % name,
                                               X).set_action('error')
    return None

  def _merge_output(self, X, y):
    if isinstance(y, ExtractorSignal):
      return y
    if y is None:
      return ExtractorSignal().set_message(
          self, "`None` value is returned by the extractor: %s" % self.__class__.__name__,
          X).set_action(self.robust_level)
    if not isinstance(y, Mapping):
      if isinstance(y, (tuple, list)):
        y = {i: j for i, j in zip(as_tuple(self.output_name, t=str), y)}
      else:
        y = {self.output_name: y}
    tmp = {}
    for name, feat in y.items():
      if any(c.isupper() for c in name):
        return ExtractorSignal().set_message(self, "Name for features cannot contain upper case",
                                             X).set_action('error')
      if feat is None:
        continue
      tmp[name] = feat
    y = tmp
    if isinstance(X, Mapping):
      for name, feat in X.items():
        if any(c.isupper() for c in name):
          return ExtractorSignal().set_message(self, "Name for features cannot contain upper case",
                                               X).set_action('error')
        if name not in y:
          y[name] = str(feat) if isinstance(feat, np.str_) else feat
    return y

  def transform(self, X):
    sig = self._check_input(X)
    if sig is not None:
      return sig
    return self._merge_output(X, self._transform(X))


# ---------------------------------------------------------------------------
# generic dictionary extractors (no arithmetic)
# ---------------------------------------------------------------------------
class Converter(Extractor):
  """base.py:396-430."""

  def __init__(self, converter, input_name='name', output_name='name'):
    super(Converter, self).__init__(input_name=as_tuple(input_name, t=str), output_name=str(output_name))
    if not hasattr(converter, '__call__') and not isinstance(converter, Mapping):
      raise ValueError("`converter` must be call-able.")
    self.converter = converter

  def _transform(self, feat):
    X = [feat[name] for name in self.input_name]
    if hasattr(self.converter, '__call__'):
      name = self.converter(*X)
    else:
      name = self.converter[X[0] if len(X) == 1 else X]
    return {self.output_name: name}


class DeltaExtractor(Extractor):
  """base.py:433-484.  Directly behind MFCCsExtractor the deltas are computed inside the fused front-end;
  anywhere else (any feature, any position) `_transform` runs signal.delta on the device (odin_sig_delta)."""

  def __init__(self, input_name, output_name=None, width=9, order=(0, 1), axis=0):
    super(DeltaExtractor, self).__init__(input_name=as_tuple(input_name, t=str), output_name=output_name)
    width = int(width)
    if width % 2 == 0 or width < 3:
      raise ValueError("`width` must be odd integer >= 3, give value: %d" % width)
    self.width = width
    self.order = as_tuple(order, t=int)
    self.axis = axis

  def _calc_deltas(self, X):   # base.py:470-481
    from . import signal
    import numpy as np
    X = np.asarray(X)
    top = max(self.order)
    deltas = signal.delta(data=X, width=self.width, order=top, axis=self.axis) if top >= 1 else ()
    deltas = (X,) + (tuple(deltas) if isinstance(deltas, (tuple, list)) else (deltas,))
    return np.concatenate([d for i, d in enumerate(deltas) if i in self.order], axis=-1)

  def _transform(self, feat):
    return [self._calc_deltas(feat[name]) for name in self.input_name]


class AsType(Extractor):
  """base.py:616-665: cast features; dtype is a type or {name: type}."""

  def __init__(self, dtype, input_name=None, exclude_pattern=r".+\_sad"):
    super(AsType, self).__init__(input_name=input_name)
    self.dtype = dtype
    self.exclude_pattern = exclude_pattern

  def _transform(self, feat):
    import re
    pat = re.compile(self.exclude_pattern) if isinstance(self.exclude_pattern, str) else None
    out = {}
    names = as_tuple(self.input_name, t=str) if self.input_name is not None else tuple(feat.keys())
    for name in names:
      X = feat[name]
      if not isinstance(X, np.ndarray):
        continue
      if pat is not None and pat.search(name):
        continue
      dt = self.dtype.get(name, None) if isinstance(self.dtype, Mapping) else self.dtype
      if dt is not None:
        out[name] = X.astype(dt)
    return out


class DuplicateFeatures(Extractor):
  """base.py:668-687."""

  def __init__(self, input_name, output_name):
    super(DuplicateFeatures, self).__init__(input_name=as_tuple(input_name, t=str),
                                            output_name=as_tuple(output_name, t=str))

  def _transform(self, feat):
    return {out: feat[inp] for inp, out in zip(self.input_name, self.output_name)}


class RenameFeatures(Extractor):
  """base.py:675-700: renames only the features that exist; anything that is not a Mapping
  (an ExtractorSignal, None) passes through untouched."""

  def __init__(self, input_name, output_name):
    super(RenameFeatures, self).__init__(input_name=as_tuple(input_name, t=str),
                                         output_name=as_tuple(output_name, t=str))

  def transform(self, X):
    if not isinstance(X, Mapping):
      return X
    X = dict(X)
    for inp, out in zip(self.input_name, self.output_name):
      if inp in X:
        X[out] = X.pop(inp)
    return X


class DeleteFeatures(Extractor):
  """base.py:703-720: removes the named features that exist."""

  def __init__(self, input_name):
    super(DeleteFeatures, self).__init__(input_name=as_tuple(input_name, t=str))

  def transform(self, X):
    if not isinstance(X, Mapping):
      return X
    return {k: v for k, v in X.items() if k not in self.input_name}


# ---------------------------------------------------------------------------
# pipeline
# ---------------------------------------------------------------------------
class StackFeatures(Extractor):
  """base.py:724-771: splice `n_context` frames of left and right context into every row
  (signal.stack_frames(frame_length=2c+1, step_length=1, keep_length=True), zeros outside the utterance)."""

  def __init__(self, n_context, input_name=None):
    super(StackFeatures, self).__init__(input_name=as_tuple(input_name, t=str) if input_name is not None else None)
    self.n_context = int(n_context)
    assert self.n_context > 0

  def transform_batch(self, Xs):
    from .. import _lib
    from .speech import _ragged_feature_call
    import torch
    _lib.require_cuda()
    lib = _lib.load()
    Xs = list(Xs)
    checked = [self._check_input(x) for x in Xs]

    def fn(d_x, dim, off):
      y = torch.empty((d_x.shape[0], (2 * self.n_context + 1) * dim), dtype=torch.float32, device='cuda')
      _lib.check(lib.odin_feat_stack(_lib.ptr(d_x), _lib.ptr(y), dim, _lib.as_i64_ptr(off), len(off) - 1,
                                     self.n_context, _lib.current_stream()))
      return y

    res = list(checked)
    live = [i for i, c in enumerate(checked) if c is None]
    if self.input_name is None:   # every 2-D array of the dictionary, per input
      for i in live:
        names = [k for k, v in Xs[i].items() if isinstance(v, np.ndarray) and v.ndim == 2]
        res[i] = _ragged_feature_call([Xs[i]], names, fn)[0]
    else:
      outs = _ragged_feature_call([Xs[i] for i in live], self.input_name, fn)
      for i, o in zip(live, outs):
        res[i] = o
    return res

  def transform(self, X):
    return self.transform_batch([X])[0]


class Pipeline(object):
  """What ``make_pipeline`` returns (base.py:96-136 returns an sklearn Pipeline;
  sklearn >= 1.x refuses ``transform`` on an unfitted pipeline, so the chain is
  run directly).  ``steps`` keeps the (name, extractor) list the user supplied;
  execution follows ``plan`` where the speech steps are fused."""

  def __init__(self, steps):
    self.steps = list(steps)
    from .speech import plan_fusion
    self.plan = plan_fusion([e for _, e in self.steps])

  named_steps = property(lambda self: dict(self.steps))

  def fit(self, X=None, y=None):
    return self

  def transform(self, X):
    return self.transform_batch([X])[0]

  __call__ = transform

  def transform_batch(self, Xs):
    """Runs a list of inputs through the pipeline; the fused speech step
    processes the whole list in one ragged batch on the GPU."""
    Xs = list(Xs)
    for step in self.plan:
      if hasattr(step, 'transform_batch'):
        Xs = step.transform_batch(Xs)
      else:
        Xs = [step.transform(x) for x in Xs]
    return Xs


def make_pipeline(steps, debug=False):
  """base.py:96-136: flattens, drops non-Extractors, names steps ClassName<i>."""
  ID = [0]

  def item2step(x):
    if isinstance(x, (tuple, list)):
      if len(x) == 1 and isinstance(x[0], Extractor):
        ID[0] += 1
        return (x[0].__class__.__name__ + str(ID[0]), x[0])
      elif len(x) == 2:
        if isinstance(x[0], str) and isinstance(x[1], Extractor):
          return tuple(x)
        elif isinstance(x[1], str) and isinstance(x[0], Extractor):
          return (x[1], x[0])
    elif isinstance(x, Extractor):
      ID[0] += 1
      return (x.__class__.__name__ + str(ID[0]), x)
    return None

  if isinstance(steps, Mapping):
    steps = steps.items()
  elif not isinstance(steps, (tuple, list)):
    steps = [steps]
  steps = [s for s in (item2step(i) for i in steps) if s is not None]
  if len(steps) == 0:
    raise ValueError("No instance of odin.preprocessing.base.Extractor found in `steps`.")
  for _, e in steps:
    e.set_debug(debug)
  return Pipeline(steps)
