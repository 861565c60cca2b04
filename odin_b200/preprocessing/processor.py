"""Batcher + feature store with the surface of ``odin.preprocessing.FeatureProcessor``
(reference: processor.py:406-816).

The reference forks ``ncpu`` workers, runs the pipeline one file at a time and
appends every feature to an on-disk ``bigarray.MmapArray`` with an
``indices_<feat>`` map name -> (start, end) (processor.py:582-653), flushing its
cache every ``n_cache`` files.  Here:

* jobs are grouped into ragged batches of ``batch_utts`` utterances for the fused
  CUDA front-end (one launch sequence per batch, ``Pipeline.transform_batch``);
* features are appended in JOB order (the reference's ncpu=1 behaviour, SURVEY.md
  8.1-Q8) to a STREAMING store: with ``path`` given every batch is appended to
  ``<feat>.npy`` as it is produced (``_NpyAppender``: the header is written with
  room for the final shape and patched on close), so host memory holds one batch,
  not the corpus; ``indices_<feat>.csv`` maps name -> (start, end).  The bigarray
  container is a third-party format outside this path;
* under ``torch.distributed`` (one process per GPU) the jobs are dealt to the
  ranks longest-first (``sharding.shard_utterances`` on the sample counts when the
  jobs carry ``raw``, else evenly) and every rank writes ITS utterances to
  ``<feat>.rank<k>.npy`` / ``indices_<feat>.rank<k>.csv``: features stay with the
  GPU that extracted them and feed ``GMM(local_shard=True)`` without ever being
  gathered.  ``shard=False`` makes every rank process every job.
"""
import os

import numpy as np

from .. import sharding
from .base import ExtractorSignal, Pipeline, make_pipeline


class _NpyAppender(object):
  """Appends row blocks to a .npy file whose leading dimension is not known in advance."""
  HEADER = 256   # bytes reserved for magic + header (numpy only needs the declared length to be consistent)

  def __init__(self, path, dtype, tail_shape):
    self.path, self.dtype, self.tail = path, np.dtype(dtype), tuple(int(d) for d in tail_shape)
    self.rows = 0
    self.f = open(path, 'wb')
    self.f.write(b'\x00' * self.HEADER)

  def append(self, block):
    block = np.ascontiguousarray(block, dtype=self.dtype)
    assert tuple(block.shape[1:]) == self.tail, "feature width changed inside one store"
    self.f.write(block.tobytes())
    self.rows += block.shape[0]

  def close(self):
    d = "{'descr': %r, 'fortran_order': False, 'shape': %r, }" % (np.lib.format.dtype_to_descr(self.dtype),
                                                                 (self.rows,) + self.tail)
    hlen = self.HEADER - 10   # version 1.0: 6 magic + 2 version + 2 length bytes
    header = d + ' ' * (hlen - len(d) - 1) + '\n'
    self.f.seek(0)
    self.f.write(b'\x93NUMPY\x01\x00' + np.uint16(hlen).tobytes() + header.encode('latin1'))
    self.f.close()


class FeatureProcessor(object):

  def __init__(self, jobs, path=None, extractor=None, n_cache=0.12, ncpu=1, override=True,
               identifier='name', log_path=None, stop_on_failure=False, batch_utts=256, shard=True):
    self.jobs = list(jobs)
    self.path = path
    if extractor is None:
      raise ValueError("`extractor` must be a pipeline or a list of Extractors")
    self.extractor = extractor if isinstance(extractor, Pipeline) else make_pipeline(extractor)
    self.identifier = str(identifier)
    self.stop_on_failure = bool(stop_on_failure)
    self.batch_utts = int(batch_utts)
    self.override = bool(override)
    self.error_log = []
    td = sharding._dist() if shard else None
    self.rank = td.get_rank() if td is not None else 0
    self.world = td.get_world_size() if td is not None else 1
    if self.world > 1:
      size = [len(j['raw']) if isinstance(j, dict) and hasattr(j.get('raw', None), '__len__') else 1 for j in self.jobs]
      self.job_ids = sharding.shard_utterances(size, self.world)[self.rank]
    else:
      self.job_ids = list(range(len(self.jobs)))

  def _file(self, stem, ext):
    return os.path.join(self.path, stem + ('.rank%d' % self.rank if self.world > 1 else '') + ext)

  def run(self):
    """Returns (features, indices): features[name] is the [N, ...] array of this rank's utterances in job order
    (a read-only memmap of the store when `path` is given), indices[name] maps utterance -> (start, end)."""
    if self.path is not None:
      os.makedirs(self.path, exist_ok=True)
    stores, mem, indices, cursor = {}, {}, {}, {}
    for b0 in range(0, len(self.job_ids), self.batch_utts):
      ids = self.job_ids[b0:b0 + self.batch_utts]
      for k, res in zip(ids, self.extractor.transform_batch([self.jobs[i] for i in ids])):
        if isinstance(res, ExtractorSignal):
          self.error_log.append(str(res))
          if res.action == 'error' or self.stop_on_failure:
            raise RuntimeError(str(res))  # processor.py:713-726
          continue
        name = res.get(self.identifier, None)
        if name is None:
          name = res.get('path', None) or ('job%d' % k)
        for feat_name, X in res.items():
          if not isinstance(X, np.ndarray) or X.ndim == 0:
            continue
          if self.path is not None:
            if feat_name not in stores:
              fn = self._file(feat_name, '.npy')
              if os.path.exists(fn) and not self.override:
                raise RuntimeError("feature store exists and override=False: %s" % fn)
              stores[feat_name] = _NpyAppender(fn, X.dtype, X.shape[1:])
            stores[feat_name].append(X)
          else:
            mem.setdefault(feat_name, []).append(X)
          s = cursor.get(feat_name, 0)
          indices.setdefault(feat_name, {})[name] = (s, s + X.shape[0])
          cursor[feat_name] = s + X.shape[0]
    if self.path is not None:
      out = {}
      for k, st in stores.items():
        st.close()
        out[k] = np.load(st.path, mmap_mode='r')
        with open(self._file('indices_%s' % k, '.csv'), 'w') as f:
          for name, (s, e) in indices[k].items():
            f.write('%s,%d,%d\n' % (name, s, e))
    else:
      out = {k: np.concatenate(v, axis=0) for k, v in mem.items()}
    self.features_, self.indices_ = out, indices
    return out, indices
