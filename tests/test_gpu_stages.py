"""The chain one stage at a time (8b `pp.signal.*`): the free functions of
odin_b200.preprocessing.signal and the stand-alone `transform` of every single extractor against golden vectors from
the reference's own functions (oracle/make_golden.py: stages_fixtures -> tests/golden/stages.npz)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, relmax
from odin_b200 import preprocessing as pp
from odin_b200.preprocessing import signal as S

pytestmark = pytest.mark.gpu
TOL = 1e-4   # north_star: relative to the matrix maximum


@pytest.fixture(scope="module")
def g():
  return np.load(os.path.join(GOLDEN, "stages.npz"))


def test_pre_emphasis_exact(g):
  # float32 with two roundings like numpy (signal.py:965): bit-exact, 1-D and the 2-D form
  assert np.array_equal(S.pre_emphasis(g["pcm"], 0.97), g["pre"])
  x2 = np.stack([g["pcm"][:4000], g["pcm"][4000:8000]])
  assert np.array_equal(S.pre_emphasis(x2, 0.95), g["pre2d"])
  with pytest.raises(ValueError):
    S.pre_emphasis(np.zeros((2, 2, 2), np.float32))


def test_stft_complex_and_energy(g):
  st, en = S.stft(g["pre"], frame_length=400, step_length=160, n_fft=512, window='hamm', energy=True)
  assert st.shape == g["stft"].shape and st.dtype == np.complex64 and en.shape == g["stft_energy"].shape
  assert relmax(st.real, g["stft"].real) < TOL and relmax(st.imag, g["stft"].imag) < TOL
  assert np.max(np.abs(st - g["stft"])) / np.max(np.abs(g["stft"])) < TOL
  assert relmax(en, g["stft_energy"]) < 2e-6
  pad = S.stft(g["pcm"], frame_length=400, step_length=240, n_fft=1024, window='hann', padding=True)
  assert pad.shape == g["stft_pad"].shape                       # T = 1 + (n + 2 (L // 2) - L) // hop
  assert np.max(np.abs(pad - g["stft_pad"])) / np.max(np.abs(g["stft_pad"])) < TOL
  sc = S.stft(g["pcm"], frame_length=200, step_length=80, n_fft=256, window='hann', scale=0.5)
  assert np.max(np.abs(sc - g["stft_scale"])) / np.max(np.abs(g["stft_scale"])) < TOL
  with pytest.raises(ValueError):
    S.stft(g["pcm"], frame_length=400, n_fft=256)


def test_power_mel_ceps_delta(g):
  st = g["stft"]
  assert relmax(S.power_spectrogram(st, 2.0), g["spec"]) < TOL
  assert relmax(S.power_spectrogram(st, 1.0), g["spec_mag"]) < TOL
  assert relmax(S.power_spectrogram(np.abs(st)[:5].astype(np.float32), 3.0), g["spec_real3"]) < TOL
  ms = S.mels_spectrogram(g["spec"], 16000, 40, fmin=64, fmax=8000, top_db=80.0)
  assert ms.shape == (58, 40) and relmax(ms, g["mspec"]) < TOL
  assert relmax(S.mels_spectrogram(g["spec"], 16000, 24, fmin=100, fmax=None, top_db=20.0), g["mspec_top20"]) < TOL
  with pytest.raises(ValueError):
    S.mels_spectrogram(g["spec"], 16000, 24, fmin=9000, fmax=8000)
  assert relmax(S.ceps_spectrogram(g["mspec"], 20, remove_first_coef=True), g["mfcc"]) < TOL
  assert relmax(S.ceps_spectrogram(g["mspec"], 13, remove_first_coef=False), g["mfcc_keep0"]) < TOL
  d1, d2 = S.delta(g["mfcc"], width=9, order=2, axis=0)
  assert relmax(d1, g["d1"]) < 1e-6 and relmax(d2, g["d2"]) < 1e-6       # fp64 on the device, float32 out
  assert relmax(S.delta(g["mfcc"], width=5, order=1, axis=0), g["d1_w5"]) < 1e-6
  assert relmax(S.delta(g["mfcc"][:, 3], width=9, order=1), g["d1_vec"]) < 1e-6
  assert relmax(S.delta(np.ascontiguousarray(g["mfcc"].T), width=9, order=1, axis=1), g["d1"].T) < 1e-6
  with pytest.raises(ValueError):
    S.delta(g["mfcc"], width=4)
  fr = np.lib.stride_tricks.sliding_window_view(g["pre"], 400)[::160]
  w = S.get_window('hamm', 400)
  assert relmax(S.get_energy(fr * w[None, :], log=True), g["stft_energy"]) < 2e-6
  assert relmax(S.get_energy(fr, log=True), g["energy_frames"]) < 2e-6


def test_tables_and_scale_helpers(g):
  assert S.mel_filters(16000, 512, 40, 64, 8000).shape == (40, 257) and S.dct_filters(21, 40).shape == (21, 40)
  d = S.dct_filters(21, 40)
  assert np.allclose(d.dot(d.T), np.eye(21), atol=1e-12)                       # orthonormal rows
  assert abs(float(S.get_window('hamm', 400).sum()) - 216.0) < 1e-9
  assert np.allclose(S.mel2hz(S.hz2mel([64.0, 999.0, 1000.0, 7999.0])), [64.0, 999.0, 1000.0, 7999.0])


def test_single_extractors_run_standalone(g):
  """Each stage extractor on its own `transform` (no fusable run around it): the reference's step-by-step chain."""
  X = {"raw": g["pcm"], "sr": 16000}
  X = pp.PreEmphasis(0.97).transform(X)
  assert np.array_equal(X["raw"], g["pre"])
  X = pp.STFTExtractor(0.025, 0.010, n_fft=512, window='hamm', energy=True).transform(X)
  assert X["stft"].dtype == np.complex64 and relmax(X["stft_energy"], g["x_stft_energy"]) < 2e-6
  X = pp.PowerSpecExtractor(2.0).transform(X)
  assert relmax(X["spec"], g["spec"]) < TOL
  X = pp.MelsSpecExtractor(40, fmin=64, fmax=8000).transform(X)
  assert relmax(X["mspec"], g["mspec"]) < TOL
  X = pp.MFCCsExtractor(20, remove_first_coef=True, first_coef_energy=True).transform(X)
  assert relmax(X["mfcc_energy"], g["x_mfcc_energy"]) < TOL
  X = pp.DeltaExtractor('mfcc', order=(0, 1, 2)).transform(X)
  assert X["mfcc"].shape == (58, 60) and relmax(X["mfcc"], g["x_mfcc"]) < TOL
  # ApplyingSAD on its own: exact row selection; a file without speech is dropped
  sad = (np.arange(58) % 3 != 0).astype(np.uint8)
  # (the reference asserts len(sad) == max(feature.shape), speech.py:1748: features must have more frames than columns)
  Y = pp.ApplyingSAD(input_name=('mspec', 'stft_energy')).transform(dict(X, sad=sad))
  assert np.array_equal(Y["mspec"], X["mspec"][sad.astype(bool)])
  assert np.array_equal(Y["stft_energy"], X["stft_energy"][sad.astype(bool)])
  with pytest.raises(AssertionError):
    pp.ApplyingSAD(input_name='mfcc')._transform(dict(X, sad=sad))          # 58 frames x 60 columns
  Z = pp.ApplyingSAD(input_name='mspec').transform(dict(X, sad=np.zeros(58, np.uint8)))
  assert isinstance(Z, pp.ExtractorSignal)
  # the fused pipeline and the step-by-step chain agree
  fused = pp.make_pipeline([pp.AudioReader(remove_dc=False), pp.PreEmphasis(0.97),
                            pp.STFTExtractor(0.025, 0.010, n_fft=512, window='hamm'), pp.PowerSpecExtractor(),
                            pp.MelsSpecExtractor(40, fmin=64, fmax=8000), pp.MFCCsExtractor(20, first_coef_energy=True),
                            pp.DeltaExtractor('mfcc', order=(0, 1, 2))]).transform({"raw": g["pcm"], "sr": 16000})
  assert relmax(fused["mfcc"], X["mfcc"]) < TOL and relmax(fused["mspec"], X["mspec"]) < TOL
