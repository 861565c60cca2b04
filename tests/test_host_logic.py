"""CPU: the C-ABI library loads and exports every declared symbol; the host
mirrors of the device integer / float32 logic agree with numpy and the oracle;
the Extractor contract behaves like the reference's (base.py:291-357)."""
import ctypes as C
import os
import pickle
import re

import numpy as np
import pytest

from conftest import ROOT
from odin_b200 import _lib
from odin_b200 import preprocessing as pp
from odin_b200.ml import GMM
from oracle import frontend as F


@pytest.fixture(scope="module")
def lib():
  return _lib.load()


def test_library_exports_every_header_symbol(lib):
  header = open(os.path.join(ROOT, "include", "odin_b200.h")).read()
  declared = set(re.findall(r"\b(odin_[a-z0-9_]+)\s*\(", header))
  declared -= {"odin_fe_config"}
  assert len(declared) >= 20
  for name in sorted(declared):
    assert hasattr(lib, name), "missing export %s" % name
    assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
  assert lib.odin_version() >= 1


def test_no_device_fails_loudly(lib):
  import torch
  if torch.cuda.is_available():
    pytest.skip("CUDA present")
  h = C.c_void_p()
  rc = lib.odin_gmm_create(60, 64, C.byref(h))
  assert rc == _lib.ODIN_ENODEVICE
  assert b"no CPU fallback" in lib.odin_last_error()
  with pytest.raises(_lib.OdinError):
    GMM(4).fit(np.zeros((10, 3), dtype=np.float32))
  pipe = pp.make_pipeline([pp.AudioReader(), pp.STFTExtractor(0.025, 0.01), pp.PowerSpecExtractor(),
                           pp.MelsSpecExtractor(24)])
  with pytest.raises(_lib.OdinError):
    pipe.transform({"raw": np.zeros(8000, dtype=np.int16), "sr": 8000})


def test_frame_offsets_integer_exact(lib):
  rng = np.random.RandomState(0)
  for L, hop in ((400, 160), (200, 40), (512, 128), (400, 400)):
    lens = np.concatenate([[L, L + 1, L + hop - 1, L + hop, 10 * L + 7], rng.randint(L, 200000, size=50)])
    so = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=so[1:])
    fo = np.zeros_like(so)
    assert lib.odin_host_frame_offsets(L, hop, _lib.as_i64_ptr(so), len(lens), _lib.as_i64_ptr(fo)) == 0
    want = [F.num_frames(int(n), L, hop) for n in lens]
    assert list(np.diff(fo)) == want == [1 + (int(n) - L) // hop for n in lens]
  so = np.array([0, 399, 1000], dtype=np.int64)
  fo = np.zeros(3, dtype=np.int64)
  assert lib.odin_host_frame_offsets(400, 160, _lib.as_i64_ptr(so), 2, _lib.as_i64_ptr(fo)) == _lib.ODIN_ESHORT
  assert list(fo) == [0, 0, 2]


def _host_smooth(lib, x, win, wrap):
  x = np.ascontiguousarray(x, dtype=np.uint8)
  out = np.zeros_like(x)
  p = C.POINTER(C.c_uint8)
  assert lib.odin_host_smooth(x.ctypes.data_as(p), len(x), win, wrap, out.ctypes.data_as(p)) == 0
  return out


def test_smooth_matches_oracle_exhaustively(lib):
  """signal.py:969-1000: every 0/1 sequence of length 5..11 for win 3 and 5, bool and
  uint8-wrapping routes, plus random long ones and other windows."""
  for n in range(5, 12):
    for code in range(2**n):
      x = np.array([(code >> i) & 1 for i in range(n)], dtype=np.uint8)
      for win in (3, 5):
        thr = 2.0 / win
        assert np.array_equal(_host_smooth(lib, x, win, 0), F.smooth_flat(x.astype(bool), win) >= thr)
        assert np.array_equal(_host_smooth(lib, x, win, 1), F.smooth_flat(x, win) >= thr)
  rng = np.random.RandomState(1)
  for win in (3, 4, 5, 7, 9):
    for _ in range(20):
      x = (rng.rand(rng.randint(win, 300)) < rng.rand()).astype(np.uint8)
      thr = 2.0 / win
      assert np.array_equal(_host_smooth(lib, x, win, 0), F.smooth_flat(x.astype(bool), win) >= thr)
      assert np.array_equal(_host_smooth(lib, x, win, 1), F.smooth_flat(x, win) >= thr)
  g = np.load(os.path.join(ROOT, "tests", "golden", "smooth.npz"))  # the REAL reference smooth()
  for i in range(int(g["n"])):
    assert np.array_equal(_host_smooth(lib, g["x%d" % i], 3, 0), g["bool3_%d" % i])
    assert np.array_equal(_host_smooth(lib, g["x%d" % i], 5, 1), g["u8_5_%d" % i])


def test_mean_std_bit_exact_with_numpy(lib):
  """signal.py:305 standardises with np.mean/np.std on float32: numpy's pairwise sum."""
  rng = np.random.RandomState(2)
  pf = C.POINTER(C.c_float)
  for n in list(range(1, 40)) + [127, 128, 129, 255, 256, 257, 298, 1000, 1023, 4097, 18001, 60000]:
    e = (rng.randn(n) * 3 + 14).astype(np.float32)
    m, s = C.c_float(), C.c_float()
    assert lib.odin_host_mean_std_f32(e.ctypes.data_as(pf), n, C.byref(m), C.byref(s)) == 0
    assert np.float32(m.value) == np.mean(e), n
    assert np.float32(s.value) == np.std(e), n


# ---------------------------------------------------------------------------
# Extractor contract (base.py:291-357)
# ---------------------------------------------------------------------------
class _Double(pp.Extractor):

  def __init__(self, input_name="x", output_name="y"):
    super(_Double, self).__init__(input_name=input_name, output_name=output_name)

  def _transform(self, X):
    return X[self.input_name] * 2


def test_extractor_contract():
  e = _Double()
  out = e.transform({"x": np.arange(3), "keep": 7})
  assert set(out) == {"x", "y", "keep"} and list(out["y"]) == [0, 2, 4]      # merge keeps old keys
  sig = e.transform({"z": 1})                                              # missing input name -> error signal
  assert isinstance(sig, pp.ExtractorSignal) and sig.action == "error"
  assert e.transform(sig) is sig                                             # signals pass through
  assert isinstance(e.transform(None), pp.ExtractorSignal)                   # None -> signal (robust level)
  assert e.transform(None).action == "ignore"
  assert isinstance(e.transform([1, 2]), pp.ExtractorSignal)                 # non-dict on a non-input layer

  class _Upper(pp.Extractor):

    def _transform(self, X):
      return {"Bad": 1}

  assert _Upper().transform({"a": 1}).action == "error"                      # upper-case names rejected

  class _Tup(pp.Extractor):

    def __init__(self):
      super(_Tup, self).__init__(input_name="x", output_name=("a", "b"))

    def _transform(self, X):
      return (1, None)

  out = _Tup().transform({"x": 0})
  assert out["a"] == 1 and "b" not in out                                    # None values dropped
  assert pp.DeleteFeatures(["x"]).transform({"x": 1, "y": 2}) == {"y": 2}
  assert pp.RenameFeatures("x", "z").transform({"x": 1}) == {"z": 1}
  assert pp.DuplicateFeatures("x", "z").transform({"x": 1}) == {"x": 1, "z": 1}
  out = pp.AsType("float16").transform({"mfcc": np.ones(3), "mfcc_sad": np.ones(3), "n": 3})
  assert out["mfcc"].dtype == np.float16 and out["mfcc_sad"].dtype == np.float64
  assert pp.speech._extract_frame_step_length(16000, 0.025, 0.010) == (400, 160)
  assert pp.speech._extract_frame_step_length(8000, 0.025, 0.005) == (200, 40)
  assert pp.speech._extract_frame_step_length(8000, 256, None) == (256, 64)


def test_make_pipeline_and_fusion_plan():
  steps = [pp.AudioReader(), pp.PreEmphasis(0.97), pp.STFTExtractor(0.025, 0.010, n_fft=512),
           pp.PowerSpecExtractor(), pp.MelsSpecExtractor(40, fmin=64, fmax=8000),
           pp.MFCCsExtractor(20, first_coef_energy=True), pp.DeltaExtractor("mfcc", order=(0, 1, 2)),
           pp.SADgmm(input_name="stft_energy"), None, "junk", pp.DeleteFeatures(["raw"]), pp.AsType("float16")]
  pipe = pp.make_pipeline(steps)
  assert [n for n, _ in pipe.steps][:2] == ["AudioReader1", "PreEmphasis2"]      # base.py:112-121 naming
  assert len(pipe.steps) == 10
  assert [type(s).__name__ for s in pipe.plan] == ["FusedSpeechFrontEnd", "DeleteFeatures", "AsType"]
  cfg = pipe.plan[0]._config(16000)
  assert (cfg.frame_len, cfg.hop, cfg.n_fft, cfg.n_mels, cfg.n_ceps) == (400, 160, 512, 40, 20)
  assert (cfg.delta_order, cfg.vad_kind, cfg.vad_smooth, cfg.window, cfg.remove_dc) == (2, 1, 3, 1, 1)
  with pytest.raises(ValueError):
    pp.make_pipeline([1, 2])
  # a lone speech extractor is planned as its own stage (signal.* on the device) ...
  assert [type(s).__name__ for s in pp.make_pipeline([pp.PreEmphasis(), pp.DeltaExtractor("raw")]).plan] == \
      ["PreEmphasis", "DeltaExtractor"]
  # ... and there is no CPU fallback behind it: without a CUDA device it fails loudly
  import torch
  if not torch.cuda.is_available():
    from odin_b200._lib import OdinError
    with pytest.raises(OdinError):
      pp.PreEmphasis().transform({"raw": np.zeros(4, np.float32)})
  # FSDD recipe wiring (examples/fsdd_ivec.py:80-106): SADthreshold on the first cepstral coefficient
  pipe = pp.make_pipeline([pp.AudioReader(), pp.PreEmphasis(), pp.STFTExtractor(0.025, 0.005, n_fft=512),
                           pp.PowerSpecExtractor(), pp.MelsSpecExtractor(24, fmin=64, fmax=4000),
                           pp.MFCCsExtractor(20, first_coef_energy=True),
                           pp.SADthreshold(input_name="mfcc_energy"), pp.DeltaExtractor("mfcc", order=(0, 1, 2))])
  cfg = pipe.plan[0]._config(8000)
  assert (cfg.vad_kind, cfg.frame_len, cfg.hop, cfg.vad_smooth) == (2, 200, 40, 5)


def test_gmm_host_surface():
  g = GMM(nmix=8, nmix_start=1, niter=4, seed=7, name="ubm")
  assert not g.is_initialized and not g.is_fitted and g.nmix == 8 and g.name == "ubm"
  X = np.zeros((100, 60), dtype=np.float32)
  g.initialize(X)
  assert g.feat_dim == 60 and g.mean.shape == (60, 1) and g.sigma.shape == (60, 1) and g.w.shape == (1, 1)
  assert g.batch_size_cpu == 52428 and g.batch_size_gpu == 109226          # gmm_tmat.py:602-607
  g2 = pickle.loads(pickle.dumps(g))                                         # same 21-tuple state
  assert len(g.__getstate__()) == 21
  assert g2.feat_dim == 60 and g2.nmix == 8 and g2.name == "ubm" and np.array_equal(g2.sigma, g.sigma)
  with pytest.raises(RuntimeError):
    g.initialize(np.zeros((5, 3), dtype=np.float32))
  # frame selection mask: indices + sad (gmm_tmat.py:162-164)
  sad = np.ones(100, dtype=np.uint8)
  sad[:10] = 0
  m = g._selected_mask(100, sad, [("a", (5, 20)), ("b", (50, 60))])
  assert int(m.sum()) == 10 + 10 and m[5:10].sum() == 0
  assert g._selected_mask(100, None, None) is None
  g.downsample = 4
  m = g._selected_mask(200000, None, None)
  assert 0 < int(m.sum()) < 200000 and m[:1].dtype == np.uint8
  assert np.array_equal(m, g._selected_mask(200000, None, None))             # seeded => reproducible


def test_downsample_picks_match_the_reference():
  """g4 (gmm_tmat.py:135-232): with downsample > 1 the frames the E-step visits must be the reference's own picks --
  observed from the real reference's batch generators (oracle/make_golden.py: downsample_fixtures), with and without
  `indices`, stochastic (seed + curr_nmix + curr_niter) and deterministic seeding, SAD applied."""
  g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gmm_downsample.npz"))
  N = g["X"].shape[0]
  indices = [("u%02d" % i, (int(s), int(s + n))) for i, (s, n) in enumerate(zip(g["starts"], g["lens"]))]
  for tag, kw in (("ds4", dict(downsample=4, stochastic_downsample=True)),
                  ("ds3det", dict(downsample=3, stochastic_downsample=False))):
    for use_idx in (0, 1):
      m = GMM(nmix=4, nmix_start=4, niter=1, batch_size_cpu=700, seed=77, **kw)
      m.initialize(g["X"])
      m._llk_hist[4] = [0.0, 0.0]     # two iterations done at this mixture count: enters the stochastic seed
      mask = m._selected_mask(N, g["sad"], indices if use_idx else None)
      ref = g["%s_idx%d_mask" % (tag, use_idx)]
      assert 0 < int(ref.sum()) < int(g["sad"].sum())
      assert np.array_equal(mask, ref), (tag, use_idx)


def test_fusion_plan_of_the_fsdd_recipe():
  """examples/fsdd_ivec.py:80-106 interleaves Converter / RenameFeatures with the speech steps: the
  run must still fuse, the bookkeeping steps are deferred behind the fused step in order, and the
  renamed SAD input (`mfcc_energy` -> `energy`) is resolved."""
  pipe = pp.make_pipeline([
      pp.AudioReader(remove_dc=True), pp.PreEmphasis(coeff=0.97),
      pp.Converter(converter=lambda x: os.path.basename(x).split('.')[0], input_name='path', output_name='name'),
      pp.STFTExtractor(frame_length=0.025, step_length=0.005, n_fft=512, window='hamm', energy=False),
      pp.PowerSpecExtractor(power=2.0, output_name='spec'),
      pp.MelsSpecExtractor(n_mels=24, fmin=64, fmax=4000, input_name=('spec', 'sr'), output_name='mspec'),
      pp.MFCCsExtractor(n_ceps=20, remove_first_coef=True, first_coef_energy=True, input_name='mspec',
                        output_name='mfcc'),
      pp.DeltaExtractor(input_name='mfcc', order=(0, 1, 2)),
      pp.RenameFeatures(input_name='mfcc_energy', output_name='energy'),
      pp.SADthreshold(energy_threshold=0.55, smooth_window=5, input_name='energy', output_name='sad'),
      pp.DeleteFeatures(input_name=('stft', 'spec', 'sad_threshold')),
      pp.AcousticNorm(mean_var_norm=True, windowed_mean_var_norm=True, input_name=('mspec', 'mfcc')),
      pp.AsType(dtype='float16')])
  assert [type(e).__name__ for e in pipe.plan] == ['FusedSpeechFrontEnd', 'Converter', 'RenameFeatures',
                                                   'DeleteFeatures', 'AcousticNorm', 'AsType']
  fused = pipe.plan[0]
  assert fused.alias == {'energy': 'mfcc_energy'} and fused.sad is not None and fused.delta is not None
  assert len(pipe.steps) == 13                                   # the user-visible step list is untouched
  # a SAD on a feature the front-end does not produce is still refused
  with pytest.raises(NotImplementedError):
    pp.make_pipeline([pp.AudioReader(), pp.STFTExtractor(0.025, 0.01), pp.PowerSpecExtractor(),
                      pp.MelsSpecExtractor(24), pp.MFCCsExtractor(20, first_coef_energy=True),
                      pp.RenameFeatures('mspec', 'energy'), pp.SADthreshold(input_name='energy')])
  # AcousticNorm argument checks (speech.py:1571-1579)
  with pytest.raises(ValueError):
    pp.AcousticNorm('mfcc', win_length=300)


def test_plan_fusion_emits_every_transparent_step_exactly_once():
  """Transparent steps (Converter / RenameFeatures / DuplicateFeatures) behind MFCC, SAD or Delta are looked at
  while the planner searches for the next fusable extractor; whether or not one follows, each user step must
  appear exactly once in the plan (a non-idempotent Converter or an in-place rename must not run twice)."""
  head = [pp.AudioReader(), pp.STFTExtractor(0.025, 0.01), pp.PowerSpecExtractor(), pp.MelsSpecExtractor(24),
          pp.MFCCsExtractor(20, first_coef_energy=True)]
  names = lambda pipe: [type(e).__name__ for e in pipe.plan]
  # nothing fusable after the rename: it stays in place, once
  p = pp.make_pipeline(head + [pp.RenameFeatures('mfcc', 'cep'), pp.AcousticNorm(input_name='cep')])
  assert names(p) == ['FusedSpeechFrontEnd', 'RenameFeatures', 'AcousticNorm']
  assert p.plan[0].alias == {}                      # the rename was not taken into the fused step
  # rename in front of the SAD (fused), duplicate in front of ApplyingSAD (fused): both deferred, once each
  p = pp.make_pipeline(head + [pp.RenameFeatures('mfcc_energy', 'energy'), pp.SADthreshold(input_name='energy'),
                               pp.DuplicateFeatures('mfcc', 'mfcc2'), pp.ApplyingSAD(input_name='mfcc')])
  assert names(p) == ['FusedSpeechFrontEnd', 'RenameFeatures', 'DuplicateFeatures']
  assert p.plan[0].sad is not None and p.plan[0].apply_sad is not None
  # ... and when no ApplyingSAD follows, the duplicate is emitted in place, once
  p = pp.make_pipeline(head + [pp.RenameFeatures('mfcc_energy', 'energy'), pp.SADthreshold(input_name='energy'),
                               pp.DuplicateFeatures('mfcc', 'mfcc2'), pp.AsType('float16')])
  assert names(p) == ['FusedSpeechFrontEnd', 'RenameFeatures', 'DuplicateFeatures', 'AsType']
  # a converter between Delta and nothing
  calls = []
  conv = pp.Converter(converter=lambda x: calls.append(x) or x, input_name='path', output_name='path2')
  p = pp.make_pipeline(head + [pp.DeltaExtractor('mfcc', order=(0, 1)), conv, pp.DeleteFeatures('stft')])
  assert names(p) == ['FusedSpeechFrontEnd', 'Converter', 'DeleteFeatures']
  for pipe in (p,):
    user = [e for _, e in pipe.steps]
    flat = [id(e) for e in pipe.plan]
    assert len(flat) == len(set(flat)) and all(id(e) in [id(u) for u in user] or type(e).__name__ == 'FusedSpeechFrontEnd'
                                               for e in pipe.plan)


def test_audio_reader_file_input_is_normalised_like_soundfile(tmp_path):
  """speech.py:127-170, 453: files go through soundfile.read (float64 in [-1, 1)) then astype(float32); arrays and
  dicts are NOT rescaled.  int16 / 32768 is exact in float32."""
  import wave
  x = (np.random.RandomState(3).randn(4000) * 8000).astype(np.int16)
  path = str(tmp_path / "a.wav")
  with wave.open(path, "wb") as f:
    f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000)
    f.writeframes(x.astype("<i2").tobytes())
  rd = pp.AudioReader()
  for inp in (path, {"path": path}, (path, None)):
    d = rd.transform(inp)
    assert d["sr"] == 16000 and d["raw"].dtype == np.float32 and d["path"] == os.path.abspath(path)
    assert np.array_equal(d["raw"], (x.astype(np.float64) / 32768.0).astype(np.float32))
  d = rd.transform({"raw": x, "sr": 16000})
  assert d["raw"].dtype == np.int16 and np.array_equal(d["raw"], x)          # arrays stay unscaled


def test_streaming_npy_store(tmp_path):
  """FeatureProcessor's store appends batches to <feat>.npy without knowing the final length (the header is patched on
  close): np.load / mmap must read back exactly the appended rows, for every dtype / trailing shape it is used with."""
  from odin_b200.preprocessing.processor import _NpyAppender
  for dt, tail in ((np.float16, (60,)), (np.uint8, ()), (np.float64, (3, 2)), (np.float32, (80,))):
    path = str(tmp_path / ("a_%s.npy" % np.dtype(dt).name))
    ap = _NpyAppender(path, dt, tail)
    blocks = [np.random.RandomState(i).rand(*((n,) + tail)).astype(dt) for i, n in enumerate((5, 0, 17, 1))]
    for b in blocks:
      ap.append(b)
    ap.close()
    a, m = np.load(path), np.load(path, mmap_mode='r')
    assert a.dtype == dt and a.shape == (23,) + tail
    assert np.array_equal(a, np.concatenate(blocks)) and np.array_equal(m, a)


def test_plan_fusion_variants():
  """Planning is host logic: which extractor runs take the reader's DC removal / the pre-emphasis into their kernels."""
  from odin_b200 import preprocessing as pp
  from odin_b200.preprocessing.speech import FusedSpeechFrontEnd
  p = pp.make_pipeline([pp.AudioReader(), pp.PreEmphasis(0.97), pp.SpectraExtractor(0.025, 0.010, n_mels=40, n_ceps=13)])
  assert [type(x).__name__ for x in p.plan] == ["SpectraExtractor"]
  assert p.plan[0]._reader is not None and p.plan[0]._preemph.coeff == 0.97 and p.plan[0].is_input_layer
  p = pp.make_pipeline([pp.AudioReader(), pp.Framing(0.025, 0.010), pp.CalculateEnergy(), pp.StackFeatures(2, "frames")])
  assert [type(x).__name__ for x in p.plan] == ["Framing", "CalculateEnergy", "StackFeatures"]
  assert p.plan[0]._reader is not None and p.plan[0]._preemph is None
  p = pp.make_pipeline([pp.AudioReader(), pp.PreEmphasis(0.97), pp.STFTExtractor(0.025, 0.010, padding=True),
                        pp.PowerSpecExtractor(), pp.MelsSpecExtractor(24), pp.MFCCsExtractor(13),
                        pp.SADgmm(input_name="stft_energy"), pp.RASTAfilter(True, 1, "mfcc")])
  assert isinstance(p.plan[0], FusedSpeechFrontEnd) and p.plan[0]._config(16000).padding == 1
  assert [type(x).__name__ for x in p.plan[1:]] == ["RASTAfilter"]
  # DC removal only exists inside the kernels: a reader that nothing fuses with must not silently skip it
  with pytest.raises(NotImplementedError):
    pp.make_pipeline([pp.AudioReader(remove_dc=True), pp.StackFeatures(2, "raw")])
  assert len(pp.make_pipeline([pp.AudioReader(remove_dc=False), pp.StackFeatures(2, "x")]).plan) == 2
