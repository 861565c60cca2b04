"""Thin batcher standing in for ``odin.preprocessing.FeatureProcessor``
(reference: processor.py:406-816).

The reference forks ``ncpu`` workers, runs the pipeline one file at a time and
appends every feature to an on-disk ``bigarray.MmapArray`` with an
``indices_<feat>`` map name -> (start, end) (processor.py:582-653).  Here jobs
are grouped into ragged batches for the fused CUDA front-end; features are
concatenated in JOB order (the reference's ncpu=1 behaviour)
and, when ``path`` is given, written as ``<feat>.npy`` + ``indices_<feat>.csv``
(the bigarray container is a third-party format outside this path).
"""
import os

import numpy as np

from .base import ExtractorSignal, Pipeline, make_pipeline


class FeatureProcessor(object):

  def __init__(self, jobs, path=None, extractor=None, n_cache=0.12, ncpu=1, override=True,
               identifier='name', log_path=None, stop_on_failure=False, batch_utts=256):
    self.jobs = list(jobs)
    self.path = path
    if extractor is None:
      raise ValueError("`extractor` must be a pipeline or a list of Extractors")
    self.extractor = extractor if isinstance(extractor, Pipeline) else make_pipeline(extractor)
    self.identifier = str(identifier)
    self.stop_on_failure = bool(stop_on_failure)
    self.batch_utts = int(batch_utts)
    self.error_log = []

  def run(self):
    feats, indices = {}, {}
    cursor = {}
    for b0 in range(0, len(self.jobs), self.batch_utts):
      batch = self.jobs[b0:b0 + self.batch_utts]
      for k, res in enumerate(self.extractor.transform_batch(batch)):
        if isinstance(res, ExtractorSignal):
          self.error_log.append(str(res))
          if res.action == 'error' or self.stop_on_failure:
            raise RuntimeError(str(res))  # processor.py:713-726
          continue
        name = res.get(self.identifier, None)
        if name is None:
          name = res.get('path', None) or ('job%d' % (b0 + k))
        for feat_name, X in res.items():
          if not isinstance(X, np.ndarray) or X.ndim == 0:
            continue
          feats.setdefault(feat_name, []).append(X)
          s = cursor.get(feat_name, 0)
          indices.setdefault(feat_name, {})[name] = (s, s + X.shape[0])
          cursor[feat_name] = s + X.shape[0]
    out = {k: np.concatenate(v, axis=0) for k, v in feats.items()}
    if self.path is not None:
      os.makedirs(self.path, exist_ok=True)
      for k, v in out.items():
        np.save(os.path.join(self.path, k + '.npy'), v)
        with open(os.path.join(self.path, 'indices_%s.csv' % k), 'w') as f:
          for name, (s, e) in indices[k].items():
            f.write('%s,%d,%d\n' % (name, s, e))
    self.features_, self.indices_ = out, indices
    return out, indices
