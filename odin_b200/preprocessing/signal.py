"""Signal-level helpers of ``odin.preprocessing.signal`` that recipes call directly (the extractor
classes live in ``speech.py`` / ``base.py``).  Arithmetic runs on the CUDA kernels behind the C-ABI;
only index bookkeeping stays on the host."""
import numpy as np

from .. import _lib


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
