"""Speech extractors with the constructor surface of ``odin.preprocessing.speech``
(reference: speech.py:345-1756) executed as ONE fused, batched CUDA path.

Every class keeps the reference's keyword arguments and feature-name wiring so
recipe code builds the same pipeline; ``plan_fusion`` (called by
``make_pipeline``) replaces the run

  [AudioReader] [PreEmphasis] STFTExtractor PowerSpecExtractor MelsSpecExtractor
  [MFCCsExtractor [DeltaExtractor]] [SADgmm | SADthreshold] [ApplyingSAD]

by a ``FusedSpeechFrontEnd`` step which launches ``odin_fe_run`` /
``odin_fe_compact`` over a ragged batch of utterances.  Intermediate spectra
('stft', 'spec') are never materialised (they stay in shared memory), and 'raw'
is passed through untouched; everything downstream ('stft_energy', 'mspec',
'mfcc', 'mfcc_energy', 'sad', 'sad_threshold') carries the reference's names.
Outputs are float32 (the reference's chained extractors return float64 for
mspec/mfcc; recipes cast to float16 right after).

A stand-alone speech extractor has no CPU implementation here and raises.
"""
import ctypes
import os
import wave
from collections.abc import Mapping
from numbers import Number

import numpy as np

from .. import _lib
from .base import (DeltaExtractor, Extractor, ExtractorSignal, as_tuple)


def _extract_frame_step_length(sr, frame_length, step_length):
  """speech.py:207-220."""
  if frame_length < 1.:
    frame_length = int(sr * frame_length)
  else:
    frame_length = int(frame_length)
  if step_length is None:
    step_length = frame_length // 4
  elif step_length < 1.:
    step_length = int(sr * step_length)
  else:
    step_length = int(step_length)
  return frame_length, step_length


def _no_cpu(cls):
  raise NotImplementedError(
      "%s has no stand-alone implementation: build the chain with make_pipeline so it runs "
      "fused on the GPU (odin_b200 has no CPU fallback)" % cls.__class__.__name__)


def read_wav(path):
  """16-bit PCM wav -> (float32 array [n] or [n, ch] in [-1, 1), sr).

  The reference decodes files through ``soundfile.read`` (speech.py:127-170), which returns float64 samples
  normalised by 32768, and then casts to float32 (speech.py:453); int16 / 32768 is exact in float32, so the
  values here are bit-identical to that route.  Only array / dict inputs stay unscaled.
  Only plain PCM is handled here (no sox / sphere fall-backs)."""
  with wave.open(path, 'rb') as f:
    if f.getsampwidth() != 2:
      raise ValueError("only 16-bit PCM wav is supported: %s" % path)
    sr, ch, n = f.getframerate(), f.getnchannels(), f.getnframes()
    raw = np.frombuffer(f.readframes(n), dtype='<i2')
  raw = raw.astype(np.float32) * np.float32(1.0 / 32768.0)
  return (raw.reshape(-1, ch) if ch > 1 else raw), sr


class AudioReader(Extractor):
  """speech.py:345-491 -> {'raw', 'sr', 'duration', 'path'[, 'name']}.
  DC removal (``remove_dc``) is applied inside the fused kernels."""

  def __init__(self, sr=None, sr_new=None, best_resample=True, remove_dc=True, dataset=None):
    super(AudioReader, self).__init__(is_input_layer=True)
    self.sr = sr
    self.sr_new = sr_new
    self.best_resample = best_resample
    self.remove_dc = bool(remove_dc)
    self.dataset = dataset
    if sr_new is not None:
      raise NotImplementedError("resampling (sr_new) is outside the accelerated path")

  def _load(self, path_or_array):
    raw = sr = name = path = None
    channel = None
    if isinstance(path_or_array, Mapping):
      if 'sr' in path_or_array:
        sr = int(path_or_array['sr'])
      if 'channel' in path_or_array:
        channel = int(path_or_array['channel'])
      if 'name' in path_or_array:
        name = path_or_array['name']
      if 'raw' in path_or_array:
        raw = path_or_array['raw']
      elif 'path' in path_or_array:
        path = str(path_or_array['path'])
        raw, sr = read_wav(path)
      else:
        raise ValueError('`path_or_array` can be a dictionary, contains following key: sr, raw, path.')
    elif isinstance(path_or_array, str):
      path = path_or_array
      if not os.path.isfile(path):
        raise ValueError("Cannot locate file at path: %s" % path_or_array)
      raw, sr = read_wav(path)
    elif isinstance(path_or_array, (tuple, list)) and len(path_or_array) == 2:
      raw, sr = path_or_array
      if isinstance(raw, str):
        path = raw
        raw, sr2 = read_wav(raw)
        sr = sr2 if sr is None else sr
    elif isinstance(path_or_array, np.ndarray):
      raw = path_or_array
    else:
      raise ValueError("`path_or_array` can be: list, tuple, Mapping, string. But given: %s" %
                       str(type(path_or_array)))
    raw = np.asarray(raw)
    if raw.ndim == 2:
      if raw.shape[0] == 2:
        raw = raw.T
      if channel is not None:
        raw = raw[:, channel]
    if raw.ndim != 1:
      raise ValueError("No support for %d-D signal from file: %s (select a `channel`)" % (raw.ndim, path))
    if sr is None and self.sr is not None:
      sr = int(self.sr)
    ret = {'raw': raw, 'sr': sr, 'duration': (max(raw.shape) / sr) if sr is not None else None,
           'path': os.path.abspath(path) if path is not None else None}
    if name is not None:
      ret['name'] = name
    return ret

  def _transform(self, X):
    return self._load(X)


class PreEmphasis(Extractor):
  """speech.py:540-563."""

  def __init__(self, coeff=0.97, input_name='raw', output_name='raw'):
    super(PreEmphasis, self).__init__(input_name=str(input_name), output_name=str(output_name))
    assert 0. < coeff < 1.
    self.coeff = float(coeff)

  def _transform(self, X):
    """Stand-alone use (outside a fusable run): signal.pre_emphasis on the device (speech.py:555-561)."""
    from . import signal
    raw = np.asarray(X[self.input_name])
    if not 0 < raw.ndim <= 2:
      raise ValueError("Only supper 1 or 2 channel audio but given shape: %s" % str(raw.shape))
    return {self.output_name: signal.pre_emphasis(raw, coeff=self.coeff)}


class STFTExtractor(Extractor):
  """speech.py:655-745."""

  def __init__(self, frame_length=None, step_length=None, n_fft=512, window='hamm', padding=False,
               energy=True, scale=None, input_name=('raw', 'sr'), output_name='stft'):
    if isinstance(input_name, str):
      input_name = (input_name, 'sr')
    assert isinstance(output_name, str), "`output_name` must be string"
    super(STFTExtractor, self).__init__(input_name=input_name, output_name=output_name)
    self.frame_length = frame_length
    self.step_length = step_length
    self.n_fft = n_fft
    self.window = window
    self.padding = bool(padding)
    self.energy = bool(energy)
    assert isinstance(scale, (str, Number, type(None)))
    self.scale = scale

  def _transform(self, X):
    """Stand-alone use: signal.stft on the device (speech.py:715-745) -> complex64 `stft` (+ `stft_energy`)."""
    from . import signal
    y, sr = [X[name] for name in self.input_name]
    scale = X[self.scale] if isinstance(self.scale, str) else self.scale
    if self.frame_length is None:
      raise NotImplementedError("framed input (frame_length=None) is not accelerated: run Framing inside the chain")
    frame_length, step_length = _extract_frame_step_length(sr, self.frame_length, self.step_length)
    res = signal.stft(y, frame_length=frame_length, step_length=step_length, n_fft=self.n_fft, window=self.window,
                      scale=scale, padding=self.padding, energy=self.energy)
    if self.energy:
      return {self.output_name: res[0], '%s_energy' % self.output_name: res[1]}
    return {self.output_name: res}


class PowerSpecExtractor(Extractor):
  """speech.py:748-763."""

  def __init__(self, power=2.0, input_name='stft', output_name='spec'):
    super(PowerSpecExtractor, self).__init__(input_name=input_name, output_name=output_name)
    self.power = float(power)

  def _transform(self, X):
    from . import signal
    return signal.power_spectrogram(S=X[self.input_name], power=self.power)   # speech.py:762-763


class MelsSpecExtractor(Extractor):
  """speech.py:766-802."""

  def __init__(self, n_mels, fmin=64, fmax=None, top_db=80.0, input_name=('spec', 'sr'),
               output_name='mspec'):
    if isinstance(input_name, str):
      input_name = (input_name, 'sr')
    super(MelsSpecExtractor, self).__init__(input_name=input_name, output_name=output_name)
    self.n_mels = int(n_mels)
    self.fmin = fmin
    self.fmax = fmax
    self.top_db = top_db

  def _transform(self, X):
    from . import signal
    return signal.mels_spectrogram(spec=X[self.input_name[0]], sr=X[self.input_name[1]], n_mels=self.n_mels,
                                   fmin=self.fmin, fmax=self.fmax, top_db=self.top_db)   # speech.py:796-802


class MFCCsExtractor(Extractor):
  """speech.py:805-831."""

  def __init__(self, n_ceps, remove_first_coef=True, first_coef_energy=False, input_name='mspec',
               output_name='mfcc'):
    super(MFCCsExtractor, self).__init__(input_name=input_name, output_name=output_name)
    self.n_ceps = int(n_ceps)
    self.remove_first_coef = bool(remove_first_coef)
    self.first_coef_energy = bool(first_coef_energy)

  def _transform(self, X):
    from . import signal
    n_ceps = self.n_ceps + (1 if self.remove_first_coef else 0)   # speech.py:821-831
    mfcc = signal.ceps_spectrogram(mspec=X[self.input_name], n_ceps=n_ceps, remove_first_coef=False)
    ret = {self.output_name: mfcc[:, 1:] if self.remove_first_coef else mfcc}
    if self.first_coef_energy:
      ret['%s_energy' % self.output_name] = mfcc[:, 0]
    return ret


class SADthreshold(Extractor):
  """speech.py:1335-1436."""

  def __init__(self, energy_threshold=0.55, energy_mean_scale=0.5, frame_context=2,
               proportion_threshold=0.12, smooth_window=5, input_name='energy', output_name='sad'):
    super(SADthreshold, self).__init__(input_name=str(input_name), output_name=str(output_name))
    self.energy_threshold = float(energy_threshold)
    self.energy_mean_scale = float(energy_mean_scale)
    self.proportion_threshold = float(proportion_threshold)
    self.frame_context = int(frame_context)
    self.smooth_window = int(smooth_window)
    assert self.energy_mean_scale > 0, 'energy_mean_scale > 0, given: %.2f' % self.energy_mean_scale
    assert self.frame_context >= 0, 'frame_context >= 0, given: %d' % self.frame_context
    assert 0. < self.proportion_threshold < 1., \
        '0 < proportion_threshold < 1, given: %.2f' % self.proportion_threshold

  def _transform(self, X):
    return _standalone_sad(self, [X])[0]

  def transform_batch(self, Xs):
    return _standalone_sad_batch(self, Xs)


class SADgmm(Extractor):
  """speech.py:1439-1477."""

  def __init__(self, nb_mixture=3, nb_train_it=24 + 1, smooth_window=3, input_name='energy',
               output_name='sad'):
    super(SADgmm, self).__init__(input_name=input_name, output_name=output_name)
    self.nb_mixture = int(nb_mixture)
    self.nb_train_it = int(nb_train_it)
    self.smooth_window = int(smooth_window)

  def _transform(self, X):
    return _standalone_sad(self, [X])[0]

  def transform_batch(self, Xs):
    return _standalone_sad_batch(self, Xs)


def _standalone_sad(ex, Xs):
  """SADgmm / SADthreshold on an energy feature that is already in the dictionary (not fused behind an STFT):
  one ragged batch through odin_vad_gmm / odin_vad_threshold.  Returns the output dicts (not merged)."""
  import torch
  _lib.require_cuda()
  lib = _lib.load()
  es = [np.ascontiguousarray(np.asarray(X[ex.input_name], dtype=np.float32).reshape(-1)) for X in Xs]
  off = np.zeros(len(es) + 1, dtype=np.int64)
  np.cumsum([len(e) for e in es], out=off[1:])
  if off[-1] == 0:
    return [{ex.output_name: np.zeros(0, np.uint8), '%s_threshold' % ex.output_name: 0.0} for _ in es]
  d_e = torch.from_numpy(np.concatenate(es)).cuda()
  d_sad = torch.zeros(int(off[-1]), dtype=torch.uint8, device='cuda')
  d_thr = torch.zeros(len(es), dtype=torch.float64, device='cuda')
  if isinstance(ex, SADgmm):
    _lib.check(lib.odin_vad_gmm(_lib.ptr(d_e), _lib.as_i64_ptr(off), len(es), ex.nb_mixture, ex.nb_train_it,
                                ex.smooth_window, 2.0, _lib.ptr(d_sad), _lib.ptr(d_thr), _lib.current_stream()))
  else:
    _lib.check(lib.odin_vad_threshold(_lib.ptr(d_e), _lib.as_i64_ptr(off), len(es), ex.energy_threshold,
                                      ex.energy_mean_scale, ex.frame_context, ex.proportion_threshold,
                                      ex.smooth_window, _lib.ptr(d_sad), _lib.ptr(d_thr), _lib.current_stream()))
  sad, thr = d_sad.cpu().numpy(), d_thr.cpu().numpy()
  outs = []
  for j in range(len(es)):
    m = sad[off[j]:off[j + 1]]
    outs.append({ex.output_name: m.copy() if isinstance(ex, SADgmm) else m.astype(bool),
                 '%s_threshold' % ex.output_name: float(thr[j])})
  return outs


def _standalone_sad_batch(ex, Xs):
  Xs = list(Xs)
  checked = [ex._check_input(x) for x in Xs]
  live = [i for i, c in enumerate(checked) if c is None]
  outs = _standalone_sad(ex, [Xs[i] for i in live]) if live else []
  res = list(checked)
  for i, o in zip(live, outs):
    res[i] = ex._merge_output(Xs[i], o)
  return res


class ApplyingSAD(Extractor):
  """speech.py:1691-1756 (``threshold`` / ``smooth_window`` re-processing of a
  continuous SAD is outside the accelerated path)."""

  def __init__(self, input_name, output_name=None, sad_name='sad', threshold=None, smooth_window=None,
               keep_unvoiced=False):
    super(ApplyingSAD, self).__init__(input_name=as_tuple(input_name, t=str), output_name=output_name)
    self.sad_name = str(sad_name)
    if threshold is not None or smooth_window is not None:
      raise NotImplementedError("ApplyingSAD(threshold=..., smooth_window=...) is not accelerated")
    self.threshold = None
    self.smooth_window = None
    self.keep_unvoiced = bool(keep_unvoiced)

  def _transform(self, X):
    """Stand-alone use (speech.py:1732-1756): rows of every input feature where `sad` is set, compacted on the
    device (odin_fe_compact); None -- the file is dropped -- when no frame is voiced and not `keep_unvoiced`."""
    import torch
    from . import signal
    sad = np.asarray(X[self.sad_name]).reshape(-1)
    if not np.any(sad) and not self.keep_unvoiced:
      return None
    lib = _lib.load()
    h = signal._fe_handle(16000, 400, 160, 512)
    d_sad = torch.from_numpy(np.ascontiguousarray(sad != 0, dtype=np.uint8)).cuda()
    fo = np.array([0, sad.shape[0]], dtype=np.int64)
    out = []
    for name in self.input_name:
      x = np.asarray(X[name])
      assert len(sad) == max(x.shape), \
          "Feature with name: %s, length of sad labels is: %d, but number of sample is: %s" % (name, len(sad), max(x.shape))
      x2 = x.reshape(x.shape[0], -1)
      d_x = torch.from_numpy(np.ascontiguousarray(x2, dtype=np.float32)).cuda()
      d_y = torch.empty_like(d_x)
      d_off = torch.empty(2, dtype=torch.int64, device='cuda')
      _lib.check(lib.odin_fe_compact(h, _lib.ptr(d_sad), _lib.as_i64_ptr(fo), 1, _lib.ptr(d_x), int(d_x.shape[1]),
                                     1 if self.keep_unvoiced else 0, _lib.ptr(d_y), _lib.ptr(d_off), _lib.current_stream()))
      n = int(d_off[1].item())
      out.append(d_y[:n].cpu().numpy().reshape((n,) + x.shape[1:]).astype(x.dtype, copy=False))
    return out


# ---------------------------------------------------------------------------
# fused execution
# ---------------------------------------------------------------------------
_WINDOWS = {'hann': 0, 'hanning': 0, 'hamm': 1, 'hamming': 1}


class AcousticNorm(Extractor):
  """speech.py:1536-1610: mean-variance normalisation (`signal.mvn`) followed by windowed mean
  normalisation (`signal.wmvn(varnorm=False)`), optionally on SAD-selected statistics.  Runs on the
  GPU (`odin_fe_cmvn`): `transform_batch` normalises a whole list of utterances with one launch per
  feature.  Outputs are float32 (the reference keeps the input dtype, float64 in its chains)."""

  def __init__(self, input_name, output_name=None, mean_var_norm=True, windowed_mean_var_norm=False,
               win_length=301, var_norm=True, sad_name=None, ignore_sad_error=True):
    self.sad_name = str(sad_name) if isinstance(sad_name, str) else None
    self.ignore_sad_error = bool(ignore_sad_error)
    super(AcousticNorm, self).__init__(input_name=as_tuple(input_name, t=str), output_name=output_name)
    self.mean_var_norm = bool(mean_var_norm)
    self.windowed_mean_var_norm = bool(windowed_mean_var_norm)
    self.var_norm = bool(var_norm)
    win_length = int(win_length)
    if win_length % 2 == 0:
      raise ValueError("win_length must be odd number")
    if win_length < 3:
      raise ValueError("win_length must >= 3")
    self.win_length = win_length

  def _transform(self, feat):
    raise NotImplementedError  # transform / transform_batch below do the work

  def transform(self, X):
    return self.transform_batch([X])[0]

  def transform_batch(self, Xs):
    _lib.require_cuda()
    import torch
    lib = _lib.load()
    results = list(Xs)
    live = []
    for i, X in enumerate(Xs):
      sig = self._check_input(X)
      if sig is not None:
        results[i] = sig
      else:
        live.append(i)
    if not live:
      return results
    outs = {i: {} for i in live}
    for name in self.input_name:
      mats, sads, use = [], [], []
      for i in live:
        x = np.asarray(Xs[i][name])
        if x.ndim != 2:
          raise ValueError("AcousticNorm expects [time, feature] matrices, %r has shape %s" % (name, x.shape))
        sad = None
        if self.sad_name is not None:
          sad = np.asarray(Xs[i][self.sad_name]).astype(bool).reshape(-1)
          if len(sad) != len(x):
            if not self.ignore_sad_error:
              raise RuntimeError("Features with name: '%s' have length %d, but given SAD has length %d" %
                                 (name, len(x), len(sad)))
            sad = None
        mats.append(np.ascontiguousarray(x, dtype=np.float32))
        sads.append(sad)
        use.append(i)
      # one launch per (has a usable SAD mask, feature width) group
      groups = {}
      for k in range(len(use)):
        groups.setdefault((sads[k] is not None, mats[k].shape[1]), []).append(k)
      for (with_sad, dim), idx in sorted(groups.items()):
        off = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum([mats[k].shape[0] for k in idx], out=off[1:])
        d_x = torch.from_numpy(np.concatenate([mats[k] for k in idx], 0)).cuda()
        d_y = torch.empty_like(d_x)
        d_sad = torch.from_numpy(np.concatenate([sads[k] for k in idx]).astype(np.uint8)).cuda() if with_sad else None
        _lib.check(lib.odin_fe_cmvn(_lib.ptr(d_x), _lib.ptr(d_y), dim, _lib.as_i64_ptr(off), len(idx),
                                    _lib.ptr(d_sad), 1 if self.mean_var_norm else 0, 1 if self.var_norm else 0,
                                    1 if self.windowed_mean_var_norm else 0, self.win_length,
                                    _lib.current_stream()))
        y = d_y.cpu().numpy()
        for j, k in enumerate(idx):
          outs[use[k]][name] = y[off[j]:off[j + 1]].copy()
    out_names = as_tuple(self.output_name, t=str) if self.output_name is not None else self.input_name
    for i in live:
      y = {on: outs[i][n] for n, on in zip(self.input_name, out_names)}
      results[i] = self._merge_output(Xs[i], y)
    return results


class FusedSpeechFrontEnd(Extractor):
  """The CUDA step standing in for a run of speech extractors (see module doc)."""

  def __init__(self, reader, preemph, stft, power, mels, mfcc, delta, sad, apply_sad, alias=None):
    super(FusedSpeechFrontEnd, self).__init__(is_input_layer=reader is not None)
    self.alias = dict(alias or {})   # feature renames deferred past this step (plan_fusion)
    self.reader, self.preemph, self.stft, self.power = reader, preemph, stft, power
    self.mels, self.mfcc, self.delta, self.sad, self.apply_sad = mels, mfcc, delta, sad, apply_sad
    self._handles = {}
    self._want = ('mspec', 'feat', 'energy', 'c0', 'sad')   # SpectraExtractor adds 'spec'
    self._spec_log = True
    self._validate()

  def _validate(self):
    st, pw, ml, mf, dl, sd, ap = (self.stft, self.power, self.mels, self.mfcc, self.delta, self.sad,
                                  self.apply_sad)
    if st.window not in _WINDOWS:
      raise NotImplementedError("window %r is not accelerated (hann / hamm only)" % (st.window,))
    if st.scale is not None or st.frame_length is None:
      raise NotImplementedError("STFTExtractor(scale / pre-framed input) is not accelerated")
    if int(pw.power) != 2:
      raise NotImplementedError("PowerSpecExtractor(power != 2) is not accelerated")
    if pw.input_name != st.output_name or ml.input_name[0] != pw.output_name:
      raise ValueError("feature names of STFT -> PowerSpec -> MelsSpec do not chain")
    if self.preemph is not None and (self.preemph.input_name != st.input_name[0] or
                                     self.preemph.output_name != st.input_name[0]):
      raise ValueError("PreEmphasis must read and write the STFT input feature")
    if mf is not None:
      if mf.input_name != ml.output_name:
        raise ValueError("MFCCsExtractor must read the MelsSpecExtractor output")
      if not mf.remove_first_coef:
        raise NotImplementedError("MFCCsExtractor(remove_first_coef=False) is not accelerated")
    if dl is not None:
      if mf is None or dl.input_name != (mf.output_name,) or dl.axis != 0:
        raise NotImplementedError("DeltaExtractor is accelerated on the MFCC feature (axis=0) only")
      if dl.order not in ((0,), (0, 1), (0, 1, 2)):
        raise NotImplementedError("DeltaExtractor order must be (0,), (0,1) or (0,1,2)")
    if sd is not None:
      sd_in = self.alias.get(sd.input_name, sd.input_name)
      if isinstance(sd, SADgmm):
        if sd_in != '%s_energy' % st.output_name or not st.energy:
          raise NotImplementedError("SADgmm is accelerated on the STFT frame energy ('%s_energy')" %
                                    st.output_name)
      else:
        if mf is None or not mf.first_coef_energy or sd_in != '%s_energy' % mf.output_name:
          raise NotImplementedError("SADthreshold is accelerated on the first cepstral coefficient "
                                    "('<mfcc>_energy', MFCCsExtractor(first_coef_energy=True))")
    if ap is not None and (sd is None or ap.sad_name != sd.output_name):
      raise ValueError("ApplyingSAD must use the SAD computed by the preceding SAD extractor")

  # -- configuration ---------------------------------------------------------
  def _config(self, sr):
    st, ml, mf, dl, sd = self.stft, self.mels, self.mfcc, self.delta, self.sad
    L, hop = _extract_frame_step_length(sr, st.frame_length, st.step_length)
    fmax = sr // 2 if ml.fmax is None else int(ml.fmax)  # signal.py:1672-1681
    fmin = int(ml.fmin)
    if fmin >= fmax:
      raise ValueError("fmin must < fmax, but fmin=%d and fmax=%d" % (fmin, fmax))
    c = _lib.FeConfig()
    c.sr, c.frame_len, c.hop, c.n_fft = int(sr), L, hop, int(st.n_fft)
    c.window = _WINDOWS[st.window]
    c.remove_dc = 1 if (self.reader is not None and self.reader.remove_dc) else 0
    c.preemph = float(self.preemph.coeff) if self.preemph is not None else 0.0
    c.n_mels, c.fmin, c.fmax = ml.n_mels, float(fmin), float(fmax)
    c.top_db = -1.0 if ml.top_db is None else float(ml.top_db)
    c.n_ceps = mf.n_ceps if mf is not None else 0
    c.delta_width = dl.width if dl is not None else 9
    c.delta_order = max(dl.order) if dl is not None else 0
    c.vad_kind = 0 if sd is None else (1 if isinstance(sd, SADgmm) else 2)
    c.vad_nmix = sd.nb_mixture if isinstance(sd, SADgmm) else 3
    c.vad_iters = sd.nb_train_it if isinstance(sd, SADgmm) else 25
    c.vad_smooth = sd.smooth_window if sd is not None else 0
    c.vad_mode = 2.0  # signal.VAD_MODE_STANDARD
    if isinstance(sd, SADthreshold):
      c.thr_energy, c.thr_mean_scale = sd.energy_threshold, sd.energy_mean_scale
      c.thr_proportion, c.thr_context = sd.proportion_threshold, sd.frame_context
    c.padding = 1 if st.padding else 0
    return c

  def _handle(self, sr):
    if sr not in self._handles:
      _lib.require_cuda()
      lib = _lib.load()
      cfg = self._config(sr)
      h = ctypes.c_void_p()
      _lib.check(lib.odin_fe_create(ctypes.byref(cfg), ctypes.byref(h)))
      self._handles[sr] = (h, cfg)
    return self._handles[sr]

  def __del__(self):
    try:
      lib = _lib.load()
      for h, _ in self._handles.values():
        lib.odin_fe_destroy(h)
      self._handles = {}
    except Exception:
      pass

  # -- device run over packed PCM ---------------------------------------------
  def run_packed(self, d_pcm, sample_offsets, sr, want=('mspec', 'feat', 'energy', 'c0', 'sad'), spec_log=True):
    """d_pcm: CUDA tensor int16 / float32 [sum n]; sample_offsets: int64 numpy
    [n_utt+1].  Returns dict of CUDA tensors + 'frame_offsets' (numpy).  'spec' in `want` adds the
    power spectrum [T, n_fft/2+1] (dB + top_db clip when spec_log; SpectraExtractor)."""
    import torch
    lib = _lib.load()
    h, cfg = self._handle(int(sr))
    n_utt = len(sample_offsets) - 1
    so = np.ascontiguousarray(sample_offsets, dtype=np.int64)
    fo = np.zeros(n_utt + 1, dtype=np.int64)
    _lib.check(lib.odin_fe_frame_offsets(h, _lib.as_i64_ptr(so), n_utt, _lib.as_i64_ptr(fo)))
    T = int(fo[-1])
    fd = cfg.n_ceps * (1 + cfg.delta_order)
    dev = d_pcm.device
    out = {'frame_offsets': fo}
    out['mspec'] = torch.empty((T, cfg.n_mels), dtype=torch.float32, device=dev)
    out['feat'] = torch.empty((T, fd), dtype=torch.float32, device=dev) if fd > 0 and 'feat' in want else None
    need_energy = ('energy' in want) or cfg.vad_kind == 1
    out['energy'] = torch.empty(T, dtype=torch.float32, device=dev) if need_energy else None
    need_c0 = cfg.n_ceps > 0 and (('c0' in want) or cfg.vad_kind == 2)
    out['c0'] = torch.empty(T, dtype=torch.float32, device=dev) if need_c0 else None
    has_sad = cfg.vad_kind != 0 and 'sad' in want
    out['sad'] = torch.empty(T, dtype=torch.uint8, device=dev) if has_sad else None
    out['sad_thr'] = torch.empty(n_utt, dtype=torch.float64, device=dev) if has_sad else None
    if d_pcm.dtype == torch.int16:
      pcm_dtype = 0
    elif d_pcm.dtype == torch.float32:
      pcm_dtype = 1
    else:
      raise ValueError("PCM must be int16 or float32")
    out['spec'] = torch.empty((T, cfg.n_fft // 2 + 1), dtype=torch.float32, device=dev) if 'spec' in want else None
    _lib.check(lib.odin_fe_run_spectra(h, _lib.ptr(d_pcm), pcm_dtype, _lib.as_i64_ptr(so), n_utt,
                                       _lib.ptr(out['mspec']), _lib.ptr(out['feat']), _lib.ptr(out['energy']),
                                       _lib.ptr(out['c0']), _lib.ptr(out['sad']), _lib.ptr(out['sad_thr']),
                                       _lib.ptr(out['spec']), 1 if spec_log else 0, _lib.current_stream()))
    return out

  def run_host_packed(self, pcm_pinned, sample_offsets, sr, want=("feat", "sad"), n_chunks=4, out=None,
                      store_dtype=None):
    """End-to-end variant of `run_packed` for HOST buffers: `pcm_pinned` is a pinned CPU tensor (int16 /
    float32) of concatenated utterances; the batch is cut into `n_chunks` groups of whole utterances and
    pipelined over three streams -- H2D copy of chunk i+1, kernels of chunk i and D2H copy of chunk i-1 run
    concurrently (PCIe is full duplex) -- into pinned host outputs.  Returns a dict of pinned CPU tensors
    for the names in `want` (from 'mspec', 'feat', 'energy', 'c0', 'sad') + 'frame_offsets'; valid after
    the call returns (it synchronises).  Measured on the config-3 shard of bench.py (223 MB in, 168 MB out):
    1 chunk 9.5 ms, 2 -> 7.2, 4 -> 6.1, 8 -> 6.7, 16 -> 8.9 (every chunk pays the SADgmm critical path of its
    longest utterance once; tools/fe_e2e_sweep.py).

    `store_dtype='float16'` narrows the float32 features ON THE DEVICE (odin_feat_convert, round to nearest even like
    the `AsType('float16')` tail of the recipes, examples/fsdd_ivec.py:105) before they are copied out: half the
    device-to-host bytes of a path that is PCIe / host-DRAM bound."""
    import torch
    lib = _lib.load()
    h, cfg = self._handle(int(sr))
    half = store_dtype is not None and np.dtype(store_dtype) == np.float16
    if store_dtype is not None and not half and np.dtype(store_dtype) != np.float32:
      raise ValueError("store_dtype must be float32 or float16")
    fdt = torch.float16 if half else torch.float32
    so = np.ascontiguousarray(sample_offsets, dtype=np.int64)
    n_utt = len(so) - 1
    fo = np.zeros(n_utt + 1, dtype=np.int64)
    _lib.check(lib.odin_fe_frame_offsets(h, _lib.as_i64_ptr(so), n_utt, _lib.as_i64_ptr(fo)))
    T = int(fo[-1])
    fd = cfg.n_ceps * (1 + cfg.delta_order)
    widths = {'mspec': (cfg.n_mels, fdt), 'feat': (fd, fdt), 'energy': (0, torch.float32),
              'c0': (0, torch.float32), 'sad': (0, torch.uint8)}
    want = tuple(w for w in want if w in widths)
    if out is None:
      out = {}
    for name in want:
      wd, dt = widths[name]
      if name not in out:
        out[name] = torch.empty((T, wd) if wd else (T,), dtype=dt, pin_memory=True)
    out['frame_offsets'] = fo
    # chunk boundaries: whole utterances, about equal numbers of samples
    if isinstance(n_chunks, (tuple, list)):   # explicit fractions of the samples per chunk (they are normalised)
      fr = np.cumsum(np.asarray(n_chunks, dtype=np.float64))
      targets = so[0] + (so[-1] - so[0]) * fr[:-1] / fr[-1]
    else:
      n_chunks = max(1, min(int(n_chunks), n_utt))
      targets = so[0] + (so[-1] - so[0]) * np.arange(1, n_chunks) / n_chunks
    cuts = [0] + sorted(set(int(c) for c in np.searchsorted(so, targets) if 0 < c < n_utt)) + [n_utt]
    chunks = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    max_s = max(int(so[b] - so[a]) for a, b in chunks)
    dev = torch.device('cuda', torch.cuda.current_device())
    main = torch.cuda.current_stream()
    # the streams and the two PCM buffers are kept: the caching allocator hands blocks back per stream, so
    # fresh streams on every call would turn each chunk's output tensors into cudaMalloc calls
    pipe = getattr(self, '_pipe', None)
    if pipe is None or pipe['dev'] != dev:
      pipe = self._pipe = {'dev': dev, 'streams': [torch.cuda.Stream() for _ in range(3)], 'pcm': None}
    s_in, s_run, s_out = pipe['streams']
    for st in (s_in, s_run, s_out):
      st.wait_stream(main)
    if pipe['pcm'] is None or pipe['pcm'][0].numel() < max_s or pipe['pcm'][0].dtype != pcm_pinned.dtype:
      pipe['pcm'] = [torch.empty(max_s, dtype=pcm_pinned.dtype, device=dev) for _ in range(2)]
    d_pcm = pipe['pcm']
    ev_in = [torch.cuda.Event() for _ in chunks]
    ev_run = [torch.cuda.Event() for _ in chunks]
    ev_out = [torch.cuda.Event() for _ in chunks]
    dev_want = tuple(w for w in ('mspec', 'feat', 'energy', 'c0', 'sad') if w in want)

    def copy_in(i):
      a, b = chunks[i]
      with torch.cuda.stream(s_in):
        if i >= 2:
          s_in.wait_event(ev_run[i - 2])          # the kernels of chunk i-2 are done with this PCM buffer
        d_pcm[i % 2][:int(so[b] - so[a])].copy_(pcm_pinned[int(so[a]):int(so[b])], non_blocking=True)
        ev_in[i].record(s_in)

    copy_in(0)
    keep = []
    for i, (a, b) in enumerate(chunks):
      if i + 1 < len(chunks):
        copy_in(i + 1)
      with torch.cuda.stream(s_run):
        s_run.wait_event(ev_in[i])
        o = self.run_packed(d_pcm[i % 2][:int(so[b] - so[a])], so[a:b + 1] - so[a], sr, want=dev_want)
        if half:
          for name in ('mspec', 'feat'):
            if name in want:
              h16 = torch.empty(o[name].shape, dtype=torch.float16, device=dev)
              _lib.check(lib.odin_feat_convert(_lib.ptr(o[name]), 1, _lib.ptr(h16), 0, o[name].numel(),
                                               _lib.current_stream()))
              o[name] = h16
        ev_run[i].record(s_run)
      with torch.cuda.stream(s_out):
        s_out.wait_event(ev_run[i])
        f0, f1 = int(fo[a]), int(fo[b])
        for name in want:
          out[name][f0:f1].copy_(o[name], non_blocking=True)
        ev_out[i].record(s_out)
      keep.append(o)   # device outputs stay referenced until their D2H copy has finished ...
      if i >= 2:       # ... and no longer: the caching allocator then hands the same blocks to the next chunk (holding
        ev_out[i - 2].synchronize()   # every chunk's outputs made each chunk a fresh cudaMalloc, which serialises
        keep[i - 2] = None            # the three streams: 100 h of audio ran at 34 GB/s H2D instead of PCIe rate)
    for st in (s_in, s_run, s_out):
      main.wait_stream(st)
    main.synchronize()
    del keep
    return out

  def compact(self, sr, d_sad, frame_offsets, d_feat, keep_unvoiced=False):
    """ApplyingSAD on device -> (rows CUDA tensor [T, dim] (first n valid), offsets CUDA int64)."""
    import torch
    lib = _lib.load()
    h, _ = self._handle(int(sr))
    n_utt = len(frame_offsets) - 1
    fo = np.ascontiguousarray(frame_offsets, dtype=np.int64)
    out = torch.empty_like(d_feat)
    off = torch.empty(n_utt + 1, dtype=torch.int64, device=d_feat.device)
    _lib.check(lib.odin_fe_compact(h, _lib.ptr(d_sad), _lib.as_i64_ptr(fo), n_utt, _lib.ptr(d_feat),
                                   int(d_feat.shape[1]), 1 if keep_unvoiced else 0, _lib.ptr(out),
                                   _lib.ptr(off), _lib.current_stream()))
    return out, off

  # -- Extractor interface ---------------------------------------------------
  def transform(self, X):
    return self.transform_batch([X])[0]

  def transform_batch(self, Xs):
    _lib.require_cuda()
    import torch
    results = [None] * len(Xs)
    loaded = {}
    for i, X in enumerate(Xs):
      sig = self._check_input(X)
      if sig is not None:
        results[i] = sig
        continue
      try:
        d = self.reader._load(X) if self.reader is not None else X
        raw_name, sr_name = self.stft.input_name
        if raw_name not in d or sr_name not in d or d[sr_name] is None:
          results[i] = ExtractorSignal().set_message(
              self, "Cannot find features with name: %s / %s" % (raw_name, sr_name), X).set_action('error')
          continue
        raw = np.asarray(d[raw_name])
        if raw.ndim != 1:
          raise ValueError("Only 1-D signals are accelerated, given shape: %s" % str(raw.shape))
        L, _ = _extract_frame_step_length(int(d[sr_name]), self.stft.frame_length, self.stft.step_length)
        if raw.shape[0] + (2 * (L // 2) if self.stft.padding else 0) < L:
          raise ValueError("signal (%d samples) shorter than one frame (%d)" % (raw.shape[0], L))
        loaded[i] = (d, raw, int(d[sr_name]))
      except Exception as e:  # the reference's workers turn exceptions into signals (processor.py:656-671)
        results[i] = ExtractorSignal().set_message(self, "%s: %s" % (type(e).__name__, e),
                                                   X).set_action(self.robust_level)
    # one ragged batch per (sample rate, dtype)
    groups = {}
    for i, (d, raw, sr) in loaded.items():
      kind = 'i2' if raw.dtype == np.int16 else 'f4'
      groups.setdefault((sr, kind), []).append(i)
    for (sr, kind), idxs in groups.items():
      raws = [loaded[i][1] if kind == 'i2' else loaded[i][1].astype(np.float32) for i in idxs]
      off = np.zeros(len(raws) + 1, dtype=np.int64)
      np.cumsum([len(r) for r in raws], out=off[1:])
      pcm = torch.from_numpy(np.concatenate(raws)).pin_memory().cuda(non_blocking=True)
      out = self.run_packed(pcm, off, sr, want=self._want, spec_log=self._spec_log)
      fo = out['frame_offsets']
      comp = None
      if self.apply_sad is not None:
        comp = {}
        for name in self.apply_sad.input_name:
          src = self._device_feature(out, name)
          if src is None:
            raise ValueError("ApplyingSAD: feature %r is not produced by the fused front-end" % name)
          rows, coff = self.compact(sr, out['sad'], fo, src, self.apply_sad.keep_unvoiced)
          comp[name] = (rows.cpu().numpy(), coff.cpu().numpy())
      host = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
      for j, i in enumerate(idxs):
        results[i] = self._emit(loaded[i][0], Xs[i], host, comp, j, int(fo[j]), int(fo[j + 1]))
    return results

  def _device_feature(self, out, name):
    if self.mfcc is not None and name == self._mfcc_out_name():
      return out['feat']
    if name == self.mels.output_name:
      return out['mspec']
    return None

  def _mfcc_out_name(self):
    if self.delta is not None:
      on = self.delta.output_name
      return on[0] if isinstance(on, tuple) else on
    return self.mfcc.output_name

  def _emit(self, base, X, host, comp, j, s, e):
    y = dict(base) if self.reader is not None else {}
    st, ml, mf, sd, ap = self.stft, self.mels, self.mfcc, self.sad, self.apply_sad
    if st.energy and host.get('energy') is not None:
      y['%s_energy' % st.output_name] = host['energy'][s:e, None].copy()
    y[ml.output_name] = host['mspec'][s:e].copy()
    if host.get('spec') is not None:
      y[self.power.output_name] = host['spec'][s:e].copy()
    if mf is not None:
      feat = host['feat'][s:e]
      if self.delta is None:
        y[mf.output_name] = feat.copy()
      else:
        y[mf.output_name] = feat[:, :mf.n_ceps].copy()
        y[self._mfcc_out_name()] = feat.copy()
      if mf.first_coef_energy:
        y['%s_energy' % mf.output_name] = host['c0'][s:e].copy()
    if sd is not None:
      sad = host['sad'][s:e]
      y[sd.output_name] = sad.copy() if isinstance(sd, SADgmm) else sad.astype(bool)
      y['%s_threshold' % sd.output_name] = float(host['sad_thr'][j])
    if ap is not None:
      names = as_tuple(ap.output_name, t=str)
      for name, oname in zip(ap.input_name, names):
        rows, coff = comp[name]
        a, b = int(coff[j]), int(coff[j + 1])
        if b - a == 0:  # speech.py:1739-1741: file dropped
          return ExtractorSignal().set_message(ap, "`None` value is returned by the extractor: ApplyingSAD",
                                               X).set_action(ap.robust_level)
        y[oname] = rows[a:b].copy()
    return self._merge_output(X if isinstance(X, Mapping) else None, y)


def plan_fusion(extractors):
  """Groups a flat list of extractors into execution steps, fusing the speech run.

  Bookkeeping extractors that recipes interleave with the speech steps (Converter, RenameFeatures,
  DuplicateFeatures -- e.g. examples/fsdd_ivec.py:83-98 renames `mfcc_energy` to `energy` right
  before SADthreshold) do not break the run: they are deferred to just after the fused step, in
  their original order, and a rename is taken into account when the SAD input name is resolved."""
  from .base import Converter, RenameFeatures, DuplicateFeatures
  plan, i, n = [], 0, len(extractors)
  speech_types = (PreEmphasis, STFTExtractor, PowerSpecExtractor, MelsSpecExtractor, MFCCsExtractor,
                  SADgmm, SADthreshold, ApplyingSAD)
  transparent = (Converter, RenameFeatures, DuplicateFeatures)
  def scan(j, alias):
    """Pure look-ahead over a run of transparent steps from position j: (position after the run, the steps,
    the alias map with their renames applied).  Nothing is committed: the caller adopts the result only when
    the extractor after the run joins the fused step, so every user step is emitted exactly once."""
    col, al = [], dict(alias)
    while j < n and isinstance(extractors[j], transparent):
      t = extractors[j]
      if isinstance(t, (RenameFeatures, DuplicateFeatures)):
        for x, y in zip(t.input_name, t.output_name):
          al[y] = al.get(x, x)
      col.append(t)
      j += 1
    return j, col, al

  while i < n:
    e = extractors[i]
    j = i
    reader = preemph = None
    deferred, alias = [], {}
    if isinstance(extractors[j], AudioReader):
      k, col, al = scan(j + 1, alias)
      if k < n and isinstance(extractors[k], (PreEmphasis, STFTExtractor)):
        reader = extractors[j]
        deferred, alias, j = deferred + col, al, k
    if j < n and isinstance(extractors[j], PreEmphasis):
      k, col, al = scan(j + 1, alias)
      if k < n and isinstance(extractors[k], STFTExtractor):
        preemph = extractors[j]
        deferred, alias, j = deferred + col, al, k
    if j + 2 < n and isinstance(extractors[j], STFTExtractor) and \
        isinstance(extractors[j + 1], PowerSpecExtractor) and isinstance(extractors[j + 2], MelsSpecExtractor):
      stft, power, mels = extractors[j:j + 3]
      j += 3
      mfcc = delta = sad = apply_sad = None
      k, col, al = scan(j, alias)
      if k < n and isinstance(extractors[k], MFCCsExtractor):
        mfcc = extractors[k]
        deferred, alias, j = deferred + col, al, k + 1
      # Delta and SAD may come in either order
      for _ in range(2):
        k, col, al = scan(j, alias)
        if k < n and delta is None and mfcc is not None and isinstance(extractors[k], DeltaExtractor):
          delta = extractors[k]
          deferred, alias, j = deferred + col, al, k + 1
        elif k < n and sad is None and isinstance(extractors[k], (SADgmm, SADthreshold)):
          sad = extractors[k]
          deferred, alias, j = deferred + col, al, k + 1
      k, col, al = scan(j, alias)
      if k < n and sad is not None and isinstance(extractors[k], ApplyingSAD):
        apply_sad = extractors[k]
        deferred, alias, j = deferred + col, al, k + 1
      # transparent steps past the last fused extractor were only looked at: the main loop emits them in place
      plan.append(FusedSpeechFrontEnd(reader, preemph, stft, power, mels, mfcc, delta, sad, apply_sad,
                                      alias=dict(alias)))
      plan.extend(deferred)
      i = j
      continue
    # [AudioReader] [PreEmphasis] SpectraExtractor | Framing: the extractor takes the reader's DC removal
    # and the pre-emphasis into its own fused run
    if isinstance(e, (AudioReader, PreEmphasis, SpectraExtractor, Framing)):
      j, rd, pe, deferred = i, None, None, []
      if isinstance(extractors[j], AudioReader):
        k, col, _ = scan(j + 1, {})
        if k < n and isinstance(extractors[k], (PreEmphasis, SpectraExtractor, Framing)):
          rd, deferred, j = extractors[j], deferred + col, k
      if j < n and isinstance(extractors[j], PreEmphasis):
        k, col, _ = scan(j + 1, {})
        if k < n and isinstance(extractors[k], (SpectraExtractor, Framing)):
          pe, deferred, j = extractors[j], deferred + col, k
      if j < n and isinstance(extractors[j], (SpectraExtractor, Framing)):
        extractors[j].bind(rd, pe)
        plan.append(extractors[j])
        plan.extend(deferred)
        i = j + 1
        continue
      if isinstance(e, AudioReader) and e.remove_dc:
        raise NotImplementedError(
            "AudioReader(remove_dc=True) at position %d is not followed by a fusable speech step: DC removal "
            "runs inside the fused kernels; odin_b200 has no CPU fallback" % i)
    # a speech extractor outside a fusable run executes on its own: one stage of the chain per call
    # (signal.pre_emphasis / stft / power_spectrogram / mels_spectrogram / ceps_spectrogram / delta on the device)
    plan.append(e)
    i += 1
  return plan


class SpectraExtractor(Extractor):
  """speech.py:849-929 -> signal.spectra (signal.py:1718-1832): STFT, power spectrum, mel
  spectrogram and MFCC in one step, outputs `spec`, `energy`, `mspec`, `mfcc` cast to float32.

  As in the reference, only `(spec, sr, n_mels)` reach `mels_spectrogram` (signal.py:1818), so the
  mel bank always spans 64 Hz .. sr/2 with an 80 dB clip whatever `fmin` / `fmax` say (they are
  validated, then ignored -- SURVEY 8.1-Q4), and `n_ceps` without `n_mels` uses 24 bands
  (signal.py:1684).  Runs on the same kernels as the chained extractors (odin_fe_run_spectra)."""

  def __init__(self, frame_length, step_length=None, n_fft=512, window='hann', n_mels=None, n_ceps=None,
               fmin=64, fmax=None, power=2.0, log=True, padding=False, input_name=('raw', 'sr')):
    super(SpectraExtractor, self).__init__(input_name=input_name)
    self.frame_length = frame_length
    self.step_length = step_length
    self.n_fft = n_fft
    self.window = window
    self.n_mels = n_mels
    self.n_ceps = n_ceps
    self.fmin = fmin
    self.fmax = fmax
    self.power = float(power)
    self.log = bool(log)
    self.padding = bool(padding)
    self._fused = None
    self._reader = self._preemph = None

  def bind(self, reader, preemph):
    """plan_fusion: the AudioReader / PreEmphasis steps preceding this extractor run inside its kernels."""
    if preemph is not None and (preemph.input_name != self.input_name[0] or preemph.output_name != self.input_name[0]):
      raise ValueError("PreEmphasis must read and write the SpectraExtractor input feature")
    self._reader, self._preemph, self._fused = reader, preemph, None
    self._is_input_layer = reader is not None

  def _front_end(self):
    if self._fused is None:
      if int(self.power) != 2:
        raise NotImplementedError("SpectraExtractor(power != 2) is not accelerated")
      stft = STFTExtractor(self.frame_length, self.step_length, n_fft=self.n_fft, window=self.window,
                           padding=self.padding, energy=True, input_name=self.input_name, output_name='stft')
      power = PowerSpecExtractor(2.0, input_name='stft', output_name='spec')
      need_mel = self.n_mels is not None or self.n_ceps is not None
      mels = MelsSpecExtractor(24 if self.n_mels is None else int(self.n_mels), fmin=64, fmax=None, top_db=80.0,
                               input_name=('spec', self.input_name[1]), output_name='mspec')
      mfcc = MFCCsExtractor(int(self.n_ceps), remove_first_coef=True, first_coef_energy=False,
                            input_name='mspec', output_name='mfcc') if self.n_ceps is not None else None
      fe = FusedSpeechFrontEnd(self._reader, self._preemph, stft, power, mels, mfcc, None, None, None)
      fe._want = ('mspec', 'feat', 'energy', 'spec')
      fe._spec_log = self.log
      self._fused = (fe, need_mel)
    return self._fused

  def transform_batch(self, Xs):
    fe, need_mel = self._front_end()
    results = [None] * len(Xs)
    live = []
    for i, X in enumerate(Xs):
      if isinstance(X, ExtractorSignal) or X is None or self._reader is None:
        sig = self._check_input(X)
        if sig is not None:
          results[i] = sig
          continue
      sr = X.get(self.input_name[1]) if isinstance(X, Mapping) else None
      if sr is None and self._reader is not None:
        sr = self._reader.sr
      fmax = (4000 if sr is None else int(sr) // 2) if self.fmax is None else int(self.fmax)   # signal.py:1798-1806
      if int(self.fmin) >= fmax:
        raise ValueError("fmin must < fmax, but fmin=%d and fmax=%d" % (int(self.fmin), fmax))
      live.append(i)
    outs = fe.transform_batch([Xs[i] for i in live])
    for i, o in zip(live, outs):
      if isinstance(o, ExtractorSignal):
        results[i] = o
        continue
      y = {'spec': o['spec'], 'energy': o['stft_energy'],
           'mspec': o['mspec'] if need_mel else None,
           'mfcc': o['mfcc'] if self.n_ceps is not None else None}
      base = {k: v for k, v in o.items() if k not in ('stft_energy', 'spec', 'mspec', 'mfcc')}
      results[i] = self._merge_output(base, y)
    return results

  def transform(self, X):
    return self.transform_batch([X])[0]

  def _transform(self, X):
    _no_cpu(self)


class Framing(Extractor):
  """speech.py:569-620: windowed frames `[T, frame_length]` (float32 here, float64 in the reference) and
  the STFT `scale`.  An AudioReader / PreEmphasis in front of it is taken into its kernel by `plan_fusion`
  (DC removal and pre-emphasis run in the staging copy, like the fused front-end)."""

  def __init__(self, frame_length, step_length=None, window='hamm', padding=False, input_name=('raw', 'sr'),
               output_name='frames'):
    if isinstance(input_name, str):
      input_name = (input_name, 'sr')
    assert isinstance(output_name, str), "`output_name` must be string"
    super(Framing, self).__init__(input_name=input_name, output_name=output_name)
    if step_length is None:
      step_length = frame_length // 4
    self.frame_length = frame_length
    self.step_length = step_length
    self.window = window
    self.padding = bool(padding)
    self._reader = self._preemph = None
    self._handles = {}
    if window not in _WINDOWS:
      raise NotImplementedError("window %r is not accelerated (hann / hamm only)" % (window,))

  def bind(self, reader, preemph):
    if preemph is not None and (preemph.input_name != self.input_name[0] or preemph.output_name != self.input_name[0]):
      raise ValueError("PreEmphasis must read and write the Framing input feature")
    self._reader, self._preemph, self._handles = reader, preemph, {}
    self._is_input_layer = reader is not None

  def _handle(self, sr):
    if sr not in self._handles:
      _lib.require_cuda()
      L, hop = _extract_frame_step_length(sr, self.frame_length, self.step_length)
      c = _lib.FeConfig()
      c.sr, c.frame_len, c.hop = int(sr), L, hop
      c.n_fft = max(256, 1 << int(np.ceil(np.log2(L))))
      c.window = _WINDOWS[self.window]
      c.remove_dc = 1 if (self._reader is not None and self._reader.remove_dc) else 0
      c.preemph = float(self._preemph.coeff) if self._preemph is not None else 0.0
      c.n_mels, c.fmin, c.fmax, c.top_db = 8, 0.0, float(sr // 2), 80.0
      c.n_ceps, c.delta_width, c.delta_order, c.vad_kind = 0, 9, 0, 0
      c.vad_nmix, c.vad_iters, c.vad_smooth, c.vad_mode = 3, 25, 0, 2.0
      c.padding = 1 if self.padding else 0
      h = ctypes.c_void_p()
      _lib.check(_lib.load().odin_fe_create(ctypes.byref(c), ctypes.byref(h)))
      self._handles[sr] = (h, c)
    return self._handles[sr]

  def __del__(self):
    try:
      for h, _ in self._handles.values():
        _lib.load().odin_fe_destroy(h)
      self._handles = {}
    except Exception:
      pass

  def transform(self, X):
    import torch
    if isinstance(X, ExtractorSignal) or X is None or self._reader is None:
      sig = self._check_input(X)
      if sig is not None:
        return sig
    d = self._reader._load(X) if self._reader is not None else X
    raw, sr = np.asarray(d[self.input_name[0]]), int(d[self.input_name[1]])
    lib = _lib.load()
    h, cfg = self._handle(sr)
    pcm = torch.from_numpy(np.ascontiguousarray(raw if raw.dtype == np.int16 else raw.astype(np.float32))).cuda()
    so = np.array([0, raw.shape[0]], dtype=np.int64)
    fo = np.zeros(2, dtype=np.int64)
    _lib.check(lib.odin_fe_frame_offsets(h, _lib.as_i64_ptr(so), 1, _lib.as_i64_ptr(fo)))
    frames = torch.empty((int(fo[1]), cfg.frame_len), dtype=torch.float32, device='cuda')
    _lib.check(lib.odin_fe_frames(h, _lib.ptr(pcm), 0 if raw.dtype == np.int16 else 1, _lib.as_i64_ptr(so), 1,
                                  _lib.ptr(frames), None, _lib.current_stream()))
    win = (ctypes.c_double * cfg.frame_len)()
    _lib.check(lib.odin_fe_get_table(h, 0, win, cfg.frame_len))
    scale = float(np.sqrt(1.0 / np.sum(np.frombuffer(win, dtype=np.float64))**2))
    return self._merge_output(d if isinstance(d, Mapping) else None, {self.output_name: frames.cpu().numpy(), 'scale': scale})

  def _transform(self, X):
    _no_cpu(self)


class CalculateEnergy(Extractor):
  """speech.py:623-649: `[T, 1]` float32 (log) energy of explicit frames (signal.get_energy, signal.py:1421-1440)."""

  def __init__(self, log=True, input_name='frames', output_name='energy'):
    super(CalculateEnergy, self).__init__(input_name=str(input_name), output_name=str(output_name))
    self.log = bool(log)

  def _transform(self, X):
    import torch
    _lib.require_cuda()
    frames = torch.from_numpy(np.ascontiguousarray(X[self.input_name], dtype=np.float32)).cuda()
    e = torch.empty(frames.shape[0], dtype=torch.float32, device='cuda')
    _lib.check(_lib.load().odin_feat_energy(_lib.ptr(frames), _lib.ptr(e), frames.shape[0], frames.shape[1],
                                            1 if self.log else 0, _lib.current_stream()))
    return {self.output_name: e.cpu().numpy()[:, None]}


def _ragged_feature_call(Xs, names, fn):
  """Packs feature `name` of every live input into one ragged [T, dim] batch per (name, dim), runs
  `fn(d_x, dim, frame_offsets) -> CUDA tensor [T, out_dim]` and scatters the rows back."""
  import torch
  results = [dict(x) if isinstance(x, Mapping) else x for x in Xs]
  live = [i for i, x in enumerate(Xs) if isinstance(x, Mapping)]
  for name in names:
    groups = {}
    for i in live:
      groups.setdefault(np.asarray(Xs[i][name]).shape[1], []).append(i)
    for dim, idxs in groups.items():
      mats = [np.ascontiguousarray(Xs[i][name], dtype=np.float32) for i in idxs]
      off = np.zeros(len(mats) + 1, dtype=np.int64)
      np.cumsum([m.shape[0] for m in mats], out=off[1:])
      y = fn(torch.from_numpy(np.concatenate(mats, 0)).cuda(), dim, off).cpu().numpy()
      for j, i in enumerate(idxs):
        results[i][name] = y[off[j]:off[j + 1]].copy()
  return results


class RASTAfilter(Extractor):
  """speech.py:1483-1533: RASTA band-pass along time (signal.rastafilt) and, when `sdc >= 1`, shifted delta
  coefficients appended (signal.shifted_deltas with N = k = n_ceps, P = 3) -> float32."""

  def __init__(self, rasta=True, sdc=1, input_name='mfcc', output_name=None):
    super(RASTAfilter, self).__init__(input_name=as_tuple(input_name, t=str), output_name=output_name)
    self.rasta = bool(rasta)
    self.sdc = int(sdc)

  def transform_batch(self, Xs):
    _lib.require_cuda()
    import torch
    lib = _lib.load()
    Xs = list(Xs)
    checked = [self._check_input(x) for x in Xs]

    def fn(d_x, dim, off):
      w = dim + dim * dim if self.sdc >= 1 else dim
      y = torch.empty((d_x.shape[0], w), dtype=torch.float32, device='cuda')
      _lib.check(lib.odin_feat_rasta_sdc(_lib.ptr(d_x), _lib.ptr(y), dim, _lib.as_i64_ptr(off), len(off) - 1,
                                         1 if self.rasta else 0, max(self.sdc, 0), _lib.current_stream()))
      return y

    live = [x if c is None else None for x, c in zip(Xs, checked)]
    outs = _ragged_feature_call(live, self.input_name, fn)
    names = as_tuple(self.output_name, t=str) if self.output_name is not None else self.input_name
    res = []
    for x, c, o in zip(Xs, checked, outs):
      if c is not None:
        res.append(c)
        continue
      res.append(self._merge_output(x, {on: o[n] for n, on in zip(self.input_name, names)}))
    return res

  def transform(self, X):
    return self.transform_batch([X])[0]

  def _transform(self, X):
    _no_cpu(self)
