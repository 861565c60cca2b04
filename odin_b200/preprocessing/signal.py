"""The free functions of ``odin.preprocessing.signal`` that recipes call directly (the extractor
classes live in ``speech.py`` / ``base.py``): same names, arguments and return shapes as the reference
(odin/preprocessing/signal.py), arithmetic on the CUDA kernels behind the C-ABI.  Array-valued results
are float32 / complex64 (the reference computes the chain in float64 / complex128; SURVEY 8.1-Q6 gives
the tolerance).  Only configuration maths (mel / Hz conversion, table look-ups) and index bookkeeping
stay on the host; there is no CPU fallback for the signal maths."""
import ctypes

import numpy as np

from .. import _lib

_WINDOWS = {'hann': 0, 'hanning': 0, 'hamm': 1, 'hamming': 1}
_HANDLES = {}


def _torch():
  import torch
  return torch


def _fe_handle(sr, frame_len, hop, n_fft, window='hann', padding=False, n_mels=8, fmin=0.0, fmax=None, top_db=80.0,
               n_ceps=0):
  """A cached front-end handle (tables on the device) for one stage configuration."""
  if window not in _WINDOWS:
    raise ValueError("window must be one of %s (the kernels build hann / hamming tables)" % sorted(_WINDOWS))
  fmax = float(sr // 2 if fmax is None else fmax)
  key = (int(sr), int(frame_len), int(hop), int(n_fft), _WINDOWS[window], bool(padding), int(n_mels), float(fmin), fmax,
         -1.0 if top_db is None else float(top_db), int(n_ceps))
  if key not in _HANDLES:
    _lib.require_cuda()
    c = _lib.FeConfig()
    (c.sr, c.frame_len, c.hop, c.n_fft, c.window, pad, c.n_mels, c.fmin, c.fmax, c.top_db, c.n_ceps) = key
    c.padding = 1 if pad else 0
    c.remove_dc, c.preemph = 0, 0.0
    c.delta_width, c.delta_order, c.vad_kind = 9, 0, 0
    c.vad_nmix, c.vad_iters, c.vad_smooth, c.vad_mode = 3, 25, 0, 2.0
    h = ctypes.c_void_p()
    _lib.check(_lib.load().odin_fe_create(ctypes.byref(c), ctypes.byref(h)))
    _HANDLES[key] = h
  return _HANDLES[key]


def _dev(x, dtype=np.float32):
  _lib.require_cuda()   # no CPU fallback: fail loudly without a device / the library
  torch = _torch()
  if isinstance(x, torch.Tensor):
    return x.to(device='cuda', dtype=getattr(torch, np.dtype(dtype).name)).contiguous()
  return torch.from_numpy(np.ascontiguousarray(x, dtype=dtype)).cuda()


# ------------------------------------------------------------------------- configuration maths (host)
def hz2mel(frequencies):
  """signal.py:489-527 (Slaney / librosa scale: linear below 1 kHz, logarithmic above)."""
  f = np.atleast_1d(np.asarray(frequencies, dtype=np.float64))
  f_sp, min_log_hz = 200.0 / 3, 1000.0
  min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
  mels = f / f_sp
  log_t = f >= min_log_hz
  mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
  return mels


def mel2hz(mels):
  """signal.py:529-568."""
  m = np.atleast_1d(np.asarray(mels, dtype=np.float64))
  f_sp, min_log_hz = 200.0 / 3, 1000.0
  min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
  freqs = f_sp * m
  log_t = m >= min_log_mel
  freqs[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
  return freqs


def _table(h, which, n):
  buf = np.zeros(int(n), dtype=np.float64)
  got = _lib.load().odin_fe_get_table(h, which, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(n))
  _lib.check(got)
  return buf


def get_window(window, frame_length, periodic=True):
  """signal.py:813-833: the periodic hann / hamming window, float64, as the kernels use it."""
  if not periodic:
    raise NotImplementedError("only the periodic (fftbins=True) windows of the speech chain are built")
  n_fft = max(256, 1 << int(np.ceil(np.log2(frame_length))))
  return _table(_fe_handle(16000, frame_length, max(1, frame_length // 4), n_fft, window), 0, frame_length)


def mel_filters(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
  """signal.py:736-810 -> [n_mels, 1 + n_fft // 2] float64 (the table the kernels project with)."""
  h = _fe_handle(sr, min(n_fft, 400), 160, n_fft, 'hann', n_mels=n_mels, fmin=fmin, fmax=fmax)
  nb = 1 + n_fft // 2
  return _table(h, 1, n_mels * nb).reshape(n_mels, nb)


def dct_filters(n_filters, n_input):
  """signal.py:683-733 -> [n_filters, n_input] float64 orthonormal DCT-II rows."""
  h = _fe_handle(16000, 400, 160, 512, 'hann', n_mels=n_input, n_ceps=n_filters - 1 if n_filters > 1 else 1)
  return _table(h, 2, max(n_filters, 2) * n_input).reshape(-1, n_input)[:n_filters]


# ------------------------------------------------------------------------- the chain, one stage at a time
def pre_emphasis(s, coeff=0.97):
  """signal.py:955-967 (odin_sig_preemph): 1-D -> append(s[0], s[1:] - coeff * s[:-1]); 2-D -> along the last axis
  with the reference's first column s[:, 0] * (1 - coeff)."""
  torch = _torch()
  s = np.asarray(s)
  if s.ndim not in (1, 2):
    raise ValueError("Only supper 1 or 2 channel audio but given shape: %s" % str(s.shape))
  x = _dev(s.reshape(-1))
  y = torch.empty_like(x)
  rows = 1 if s.ndim == 1 else s.shape[0]
  off = (np.arange(rows + 1, dtype=np.int64) * s.shape[-1])
  _lib.check(_lib.load().odin_sig_preemph(_lib.ptr(x), _lib.ptr(y), _lib.as_i64_ptr(off), rows, float(coeff),
                                          0 if s.ndim == 1 else 1, _lib.current_stream()))
  return y.cpu().numpy().reshape(s.shape)


def get_energy(frames, log=True):
  """signal.py:1421-1440 (odin_feat_energy): [n_frames, frame_length] -> [n_frames, 1] (log) sum of squares."""
  torch = _torch()
  f = _dev(frames)
  if f.dim() != 2:
    raise ValueError("`frames` must be 2-D [n_frames, frame_length]")
  e = torch.empty(f.shape[0], dtype=torch.float32, device='cuda')
  _lib.check(_lib.load().odin_feat_energy(_lib.ptr(f), _lib.ptr(e), f.shape[0], f.shape[1], 1 if log else 0,
                                          _lib.current_stream()))
  return e.cpu().numpy()[:, None]


def stft(y, frame_length=None, step_length=None, n_fft=None, window='hann', scale=None, padding=False, energy=False):
  """signal.py:1442-1562 (odin_fe_stft): 1-D signal -> complex64 [T, 1 + n_fft // 2] scaled by 1 / sum(window)
  (or `scale`), T = 1 + (n - frame_length) // step_length; with `energy` also the float32 log frame energy [T, 1]."""
  torch = _torch()
  y = np.asarray(y)
  if y.ndim != 1:
    raise NotImplementedError("only 1-D signals are accelerated (framed input: use Framing + the fused chain)")
  if frame_length is None:
    raise ValueError("`frame_length` must be given for a 1-D signal")
  frame_length = int(frame_length)
  step_length = frame_length // 4 if step_length is None else int(step_length)
  if n_fft is None:
    n_fft = int(2 ** np.ceil(np.log(frame_length) / np.log(2.0)))
  elif n_fft < frame_length:
    raise ValueError('n_fft must be greater than or equal to `frame_length`.')
  if window is None:
    raise NotImplementedError("window=None (rectangular) is not built by the kernels")
  h = _fe_handle(16000, frame_length, step_length, int(n_fft), window, padding=padding)
  pcm = _dev(y, np.int16 if y.dtype == np.int16 else np.float32)
  so = np.array([0, y.shape[0]], dtype=np.int64)
  fo = np.zeros(2, dtype=np.int64)
  lib = _lib.load()
  _lib.check(lib.odin_fe_frame_offsets(h, _lib.as_i64_ptr(so), 1, _lib.as_i64_ptr(fo)))
  T = int(fo[1])
  S = torch.empty((T, n_fft // 2 + 1, 2), dtype=torch.float32, device='cuda')
  e = torch.empty(T, dtype=torch.float32, device='cuda') if energy else None
  _lib.check(lib.odin_fe_stft(h, _lib.ptr(pcm), 0 if y.dtype == np.int16 else 1, _lib.as_i64_ptr(so), 1, _lib.ptr(S),
                              _lib.ptr(e), _lib.current_stream()))
  S = torch.view_as_complex(S).cpu().numpy()
  if scale is not None:   # the kernels scale by 1 / sum(window) (signal.py:1547)
    S = S * np.complex64(float(scale) * float(np.sum(get_window(window, frame_length))))
  return (S, e.cpu().numpy()[:, None]) if energy else S


def power_spectrogram(S, power=2.0):
  """signal.py:1623-1648 (odin_sig_power): |S| ** int(power) for complex input, S ** int(power) for real input."""
  torch = _torch()
  power = int(power)
  S = np.asarray(S) if not isinstance(S, torch.Tensor) else S
  cplx = 'complex' in str(S.dtype)
  if not cplx and power <= 1:
    return np.asarray(S)
  if cplx:
    d = torch.view_as_real(_dev(S, np.complex64)) if not isinstance(S, torch.Tensor) else torch.view_as_real(S.cuda().to(torch.complex64)).contiguous()
    n = d.numel() // 2
  else:
    d = _dev(S)
    n = d.numel()
  out = torch.empty(tuple(S.shape), dtype=torch.float32, device='cuda')
  _lib.check(_lib.load().odin_sig_power(_lib.ptr(d), 1 if cplx else 0, max(power, 1), _lib.ptr(out), n, _lib.current_stream()))
  return out.cpu().numpy()


def mels_spectrogram(spec, sr, n_mels, fmin=64, fmax=None, top_db=80.0):
  """signal.py:1650-1691 (odin_fe_mels): power spectrum [T, 1 + n_fft // 2] -> dB mel spectrogram [T, n_mels],
  clipped at (max of the matrix - top_db)."""
  torch = _torch()
  spec = np.asarray(spec)
  n_fft = int(2 * (spec.shape[1] - 1))
  if sr is None and fmax is None:
    fmax = 4000
  else:
    fmax = sr // 2 if fmax is None else int(fmax)
  fmin = int(fmin)
  if fmin >= fmax:
    raise ValueError("fmin must < fmax, but fmin=%d and fmax=%d" % (fmin, fmax))
  n_mels = 24 if n_mels is None else int(n_mels)
  h = _fe_handle(int(sr if sr is not None else 2 * fmax), min(n_fft, 400), 160, n_fft, 'hann', n_mels=n_mels, fmin=fmin,
                 fmax=fmax, top_db=top_db)
  d = _dev(spec)
  out = torch.empty((spec.shape[0], n_mels), dtype=torch.float32, device='cuda')
  off = np.array([0, spec.shape[0]], dtype=np.int64)
  _lib.check(_lib.load().odin_fe_mels(h, _lib.ptr(d), _lib.as_i64_ptr(off), 1, _lib.ptr(out), 1, _lib.current_stream()))
  return out.cpu().numpy()


def ceps_spectrogram(mspec, n_ceps, remove_first_coef=True):
  """signal.py:1693-1716 (odin_fe_ceps): orthonormal DCT-II of the mel bands -> [T, n_ceps]."""
  torch = _torch()
  mspec = np.asarray(mspec)
  n_ceps, first = int(n_ceps), (1 if remove_first_coef else 0)
  h = _fe_handle(16000, 400, 160, 512, 'hann', n_mels=mspec.shape[1], n_ceps=max(n_ceps + first - 1, 1))
  d = _dev(mspec)
  out = torch.empty((mspec.shape[0], n_ceps), dtype=torch.float32, device='cuda')
  _lib.check(_lib.load().odin_fe_ceps(h, _lib.ptr(d), mspec.shape[0], first, n_ceps, _lib.ptr(out), _lib.current_stream()))
  return out.cpu().numpy()


def delta(data, width=9, order=1, axis=0):
  """signal.py:1002-1066 (odin_sig_delta): float32 delta (order 1) or [delta, delta-delta] (order 2) along `axis`
  of a 1-D / 2-D array, with the reference's lfilter delay on the second pass."""
  torch = _torch()
  data = np.atleast_1d(np.asarray(data))
  if width < 3 or np.mod(width, 2) != 1:
    raise ValueError('width must be an odd integer >= 3')
  order = int(order)
  if order <= 0:
    raise ValueError('order must be a positive integer')
  if order > 2 or data.ndim > 2:
    raise NotImplementedError("delta is accelerated for order 1 or 2 on 1-D / 2-D data")
  x = data.reshape(-1, 1) if data.ndim == 1 else (data if axis in (0, -2) else data.T)
  d = _dev(x)
  T, dim = d.shape
  d1 = torch.empty_like(d)
  d2 = torch.empty_like(d) if order == 2 else None
  off = np.array([0, T], dtype=np.int64)
  _lib.check(_lib.load().odin_sig_delta(_lib.ptr(d), dim, _lib.as_i64_ptr(off), 1, int(width), order, _lib.ptr(d1),
                                        _lib.ptr(d2), _lib.current_stream()))

  def back(t):
    a = t.cpu().numpy()
    return a.reshape(data.shape) if data.ndim == 1 else (a if axis in (0, -2) else a.T)

  return back(d1) if order == 1 else [back(d1), back(d2)]


def smooth(x, win=11, window='hanning'):
  """signal.py:969-1000 for the `window='flat'` moving average the SAD extractors use (odin_feat_smooth), on a 0/1
  vector; other windows are not part of the accelerated path."""
  torch = _torch()
  if window != 'flat':
    raise NotImplementedError("only window='flat' (the SAD smoothing) is accelerated")
  x = np.asarray(x)
  if x.ndim != 1:
    raise ValueError("smooth only accepts 1 dimension arrays.")
  if x.size < win:
    raise ValueError("Input vector needs to be bigger than window size.")
  if win < 3:
    return x
  d = torch.from_numpy(np.ascontiguousarray(x != 0, dtype=np.uint8)).cuda()
  y = torch.empty(x.shape[0], dtype=torch.float64, device='cuda')
  _lib.check(_lib.load().odin_feat_smooth(_lib.ptr(d), _lib.ptr(y), x.shape[0], int(win), _lib.current_stream()))
  return y.cpu().numpy()


def vad_split_audio(s, sr, maximum_duration=30, minimum_duration=None, frame_length=128, nb_mixtures=3,
                    threshold=0.6, return_vad=False, return_voices=False, return_cut=False):
  """signal.py:341-478: split a long recording at low-energy points.

  Non-overlapping frames of `frame_length` samples (zero-padded at the end) -> log energy
  (odin_feat_energy) -> 3-mixture energy VAD with 33 EM iterations (odin_vad_gmm, the SADgmm kernels) ->
  flat smoothing over `frame_length` frames (odin_feat_smooth) -> frames at or above the `threshold`
  percentile are cut candidates -> greedy grouping up to `maximum_duration` seconds, short groups merged
  (host, a few hundred indices).  Returns the list of segments (views of `s`), then the optional outputs in
  the reference's order."""
  import torch
  _lib.require_cuda()
  lib = _lib.load()
  frame_length = int(frame_length)
  max_d = maximum_duration * sr
  if len(s) < max_d:
    if return_cut or return_vad or return_voices:
      raise ValueError("Cannot return `cut` points, `vad` or `voices` since the original audio is shorter than "
                       "`maximum_duration`, hence, no need for splitting.")
    return [s]
  max_d /= frame_length
  if minimum_duration is None:
    min_d = max_d // 2
  else:
    min_d = np.clip(minimum_duration * sr / frame_length, 0., 0.99 * max_d)
  # ---- device part: energies, VAD, smoothing
  nfr = -(-len(s) // frame_length)
  d_s = torch.zeros(nfr * frame_length, dtype=torch.float32, device='cuda')
  d_s[:len(s)] = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32)).cuda()
  d_e = torch.empty(nfr, dtype=torch.float32, device='cuda')
  _lib.check(lib.odin_feat_energy(_lib.ptr(d_s), _lib.ptr(d_e), nfr, frame_length, 1, _lib.current_stream()))
  off = np.array([0, nfr], dtype=np.int64)
  d_v = torch.zeros(nfr, dtype=torch.uint8, device='cuda')
  _lib.check(lib.odin_vad_gmm(_lib.ptr(d_e), _lib.as_i64_ptr(off), 1, int(nb_mixtures), 33, 0, 2.0, _lib.ptr(d_v), None,
                              _lib.current_stream()))
  if frame_length >= 3 and nfr >= frame_length:
    d_y = torch.empty(nfr, dtype=torch.float64, device='cuda')
    _lib.check(lib.odin_feat_smooth(_lib.ptr(d_v), _lib.ptr(d_y), nfr, frame_length, _lib.current_stream()))
    vad = d_y.cpu().numpy()
  else:
    vad = d_v.cpu().numpy().astype(bool)
  # ---- host part: cut candidates and greedy grouping (signal.py:421-468)
  results = []
  if return_vad:
    results.append(vad)
  indices = np.where(vad >= np.percentile(vad, q=threshold * 100))[0].tolist()
  if len(vad) - 1 not in indices:
    indices.append(len(vad) - 1)
  if return_voices:
    tmp = np.zeros(shape=(len(vad),))
    tmp[indices] = 1
    results.append(tmp)
  segments, start, prev_end = [], 0, 0
  for end in indices:
    if end - start > max_d:
      segments.append((start, prev_end))
      start = prev_end
    elif end - start == max_d:
      segments.append((start, end))
      start = end
    prev_end = end
  if len(segments) == 0:
    segments = [(indices[0], indices[-1])]
  if indices[-1] != segments[-1][-1]:
    segments.append((start, indices[-1]))
  found_under_length = True
  while found_under_length and len(segments) > 1:
    merged, found_under_length = [], False
    for (s1, e1), (s2, e2) in zip(segments, segments[1:]):
      if (e1 - s1) < min_d or (e2 - s2) < min_d:
        merged.append((s1, e2))
        found_under_length = True
      else:
        merged.append((s1, e1))
        merged.append((s2, e2))
    segments = merged
  if return_cut:
    tmp = np.zeros(shape=(segments[-1][-1] + 1,))
    for i, j in segments:
      tmp[i] = 1
      tmp[j] = 1
    results.append(tmp)
  bounds = [[i * frame_length, j * frame_length] for i, j in segments]
  bounds[-1][-1] = s.shape[0]
  out = [s[i:j] for i, j in bounds]
  results = [out] + results
  return results[0] if len(results) == 1 else results
