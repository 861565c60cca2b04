"""PLDA / cosine scoring back-end against golden vectors produced by the REAL reference
(odin/ml/plda.py, odin/ml/scoring.py run under oracle/ref_shim.py, oracle/make_golden.py: scoring_fixtures).
Host float64 linear algebra: runs without a GPU."""
import os
import pickle

import numpy as np
import pytest

from odin_b200 import ml

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scoring.npz")


@pytest.fixture(scope="module")
def g():
  return np.load(GOLD)


def close(a, b, tol=1e-9):
  return float(np.max(np.abs(np.asarray(a) - b))) <= tol * max(1.0, float(np.max(np.abs(b))))


def test_plda_em_matches_reference(g):
  p = ml.PLDA(n_phi=8, n_iter=12, centering=True, wccn=True, unit_length=True, random_state=1234)
  p.fit(g["X"], g["y"])
  assert p.is_fitted and p.num_classes == 10 and p.feat_dim == 24
  for mine, ref in ((p.Phi_, "plda_Phi"), (p.Sigma_, "plda_Sigma"), (p.Lambda_, "plda_Lambda"), (p.Q_hat_, "plda_Qhat")):
    assert close(mine, g[ref]), ref
  # Uk / X_model are singular vectors: defined up to a sign per column, which the score is invariant to
  sg = np.sign(np.sum(p.Uk_ * g["plda_Uk"], axis=0))
  assert close(p.Uk_ * sg, g["plda_Uk"]) and close(p.X_model_ * sg, g["plda_Xmodel"])
  assert close(p.transform(g["Xt"]) * sg, g["plda_proj"])
  assert close(p.predict_log_proba(g["Xt"]), g["plda_scores"])
  assert abs(p.compute_llk(p.normalizer.transform(g["X"])) - float(g["plda_llk"])) < 1e-6 * abs(float(g["plda_llk"]))
  # enrolment vectors supplied by the caller (plda.py:404-408)
  enroll = ml.scoring.compute_class_avg(g["Xt"], g["yt"], np.unique(g["yt"]))
  assert close(p.predict_log_proba(g["Xt"], X_model=enroll), g["plda_enroll_scores"])
  # the scores identify the synthetic classes
  assert np.mean(p.predict(g["Xt"]) == g["yt"]) > 0.9
  # pickle round trip: the reference's 16-tuple (plda.py:147-160)
  q = pickle.loads(pickle.dumps(p))
  assert len(p.__getstate__()) == 16 and close(q.predict_log_proba(g["Xt"]), g["plda_scores"])


def test_plda_maximum_likelihood_and_error_behaviour(g):
  pm = ml.PLDA(n_phi=8, random_state=3).fit_maximum_likelihood(g["X"], g["y"])
  assert close(pm.predict_log_proba(g["Xt"]), g["plda_ml_scores"])
  # n_iter='auto' takes the Cholesky of Phi Phi^T + Sigma every iteration; on this data the reference raises
  # LinAlgError (the first M-steps leave Sigma_ indefinite) -- same arithmetic, same failure
  if int(g["plda_auto_raises"]):
    with pytest.raises(np.linalg.LinAlgError):
      ml.PLDA(n_phi=8, n_iter='auto', improve_threshold=1e-1, random_state=7).fit(g["X"], g["y"])
  with pytest.raises(RuntimeError):
    ml.PLDA(n_phi=24, random_state=0).fit(g["X"], g["y"])          # feat_dim must exceed n_phi (plda.py:169-171)
  with pytest.raises(RuntimeError):
    ml.PLDA(n_phi=4).transform(g["Xt"])                             # not fitted


def test_cosine_scorer_and_vector_normalizer(g):
  for lda in (True, False):
    s = ml.Scorer(centering=True, wccn=True, lda=lda, method='cosine').fit(g["X"], g["y"])
    assert close(s.transform(g["Xt"]), g["cos_scores_lda%d" % lda])
    assert close(s.normalizer.enroll_vecs, g["cos_enroll_lda%d" % lda])
    assert np.mean(s.predict(g["Xt"]) == g["yt"]) > 0.9
  vn = ml.VectorNormalizer(centering=True, wccn=True, unit_length=True, lda=True, concat=True).fit(g["X"], g["y"])
  assert close(vn.transform(g["Xt"]), g["vn_concat"]) and close(vn.W, g["vn_W"]) and close(vn.mean, g["vn_mean"])
  assert close(vn.vmin, g["vn_vmin"]) and close(vn.vmax, g["vn_vmax"])
  with pytest.raises(ValueError):
    ml.Scorer(method='gmm')
  with pytest.raises(RuntimeError):
    ml.VectorNormalizer().transform(g["Xt"])
