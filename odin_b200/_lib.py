"""ctypes binding of libodin_b200.so (C-ABI declared in include/odin_b200.h).

The library is built in-tree by ``odin_b200/csrc/build.sh`` (also called from
``__graft_entry__.build()``).  There is no CPU fallback: if the shared object is
missing, or no CUDA device is present, the product path raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ODIN_B200_LIB selects another build of the same library (A/B timing of two builds in tools/)
LIB_PATH = os.environ.get("ODIN_B200_LIB") or os.path.join(_HERE, "lib", "libodin_b200.so")

ODIN_OK, ODIN_EINVAL, ODIN_ENODEVICE, ODIN_ECUDA, ODIN_ENOMEM, ODIN_ESHORT = 0, -1, -2, -3, -4, -5


class OdinError(RuntimeError):

  def __init__(self, code, msg):
    super(OdinError, self).__init__("libodin_b200 error %d: %s" % (code, msg))
    self.code = code


class FeConfig(C.Structure):
  """Mirror of ``odin_fe_config``."""
  _fields_ = [
      ("sr", C.c_int32), ("frame_len", C.c_int32), ("hop", C.c_int32), ("n_fft", C.c_int32),
      ("window", C.c_int32), ("remove_dc", C.c_int32), ("preemph", C.c_float),
      ("n_mels", C.c_int32), ("fmin", C.c_float), ("fmax", C.c_float), ("top_db", C.c_float),
      ("n_ceps", C.c_int32), ("delta_width", C.c_int32), ("delta_order", C.c_int32),
      ("vad_kind", C.c_int32), ("vad_nmix", C.c_int32), ("vad_iters", C.c_int32),
      ("vad_smooth", C.c_int32), ("vad_mode", C.c_float), ("thr_energy", C.c_float),
      ("thr_mean_scale", C.c_float), ("thr_proportion", C.c_float), ("thr_context", C.c_int32),
      ("padding", C.c_int32),
  ]


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol declared in include/odin_b200.h
SIGNATURES = {
    "odin_last_error": (C.c_char_p, []),
    "odin_version": (C.c_int, []),
    "odin_launch_count": (_i64, []),
    "odin_fe_create": (C.c_int, [C.POINTER(FeConfig), C.POINTER(_vp)]),
    "odin_fe_destroy": (None, [_vp]),
    "odin_fe_feat_dim": (C.c_int, [_vp]),
    "odin_fe_frame_offsets": (C.c_int, [_vp, _pi64, _i32, _pi64]),
    "odin_host_smooth": (C.c_int, [C.POINTER(C.c_uint8), _i32, _i32, _i32, C.POINTER(C.c_uint8)]),
    "odin_host_mean_std_f32": (C.c_int, [C.POINTER(C.c_float), _i32, C.POINTER(C.c_float),
                                         C.POINTER(C.c_float)]),
    "odin_host_frame_offsets": (C.c_int, [_i32, _i32, _pi64, _i32, _pi64]),
    "odin_fe_get_table": (C.c_int, [_vp, _i32, C.POINTER(C.c_double), _i64]),
    "odin_fe_run": (C.c_int, [_vp, _vp, _i32, _pi64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "odin_fe_run_spectra": (C.c_int, [_vp, _vp, _i32, _pi64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "odin_tmat_create": (C.c_int, [_i32, _i32, _i32, C.POINTER(_vp)]),
    "odin_tmat_destroy": (None, [_vp]),
    "odin_tmat_acc_size": (_i64, [_vp]),
    "odin_tmat_set_model": (C.c_int, [_vp, _vp, _vp, _vp]),
    "odin_tmat_get_model": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "odin_tmat_estep": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "odin_tmat_mstep": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "odin_tmat_ivector": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "odin_fe_frames": (C.c_int, [_vp, _vp, _i32, _pi64, _i32, _vp, _vp, _vp]),
    "odin_feat_stack": (C.c_int, [_vp, _vp, _i32, _pi64, _i32, _i32, _vp]),
    "odin_feat_rasta_sdc": (C.c_int, [_vp, _vp, _i32, _pi64, _i32, _i32, _i32, _vp]),
    "odin_feat_energy": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "odin_feat_smooth": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "odin_feat_convert": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _vp]),
    "odin_fe_stft": (C.c_int, [_vp, _vp, _i32, _pi64, _i32, _vp, _vp, _vp]),
    "odin_sig_preemph": (C.c_int, [_vp, _vp, _pi64, _i32, C.c_float, _i32, _vp]),
    "odin_sig_power": (C.c_int, [_vp, _i32, _i32, _vp, _i64, _vp]),
    "odin_fe_mels": (C.c_int, [_vp, _vp, _pi64, _i32, _vp, _i32, _vp]),
    "odin_fe_ceps": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "odin_sig_delta": (C.c_int, [_vp, _i32, _pi64, _i32, _i32, _i32, _vp, _vp, _vp]),
    "odin_vad_gmm": (C.c_int, [_vp, _pi64, _i32, _i32, _i32, _i32, C.c_float, _vp, _vp, _vp]),
    "odin_vad_threshold": (C.c_int, [_vp, _pi64, _i32, C.c_float, C.c_float, _i32, C.c_float, _i32, _vp, _vp, _vp]),
    "odin_fe_compact": (C.c_int, [_vp, _vp, _pi64, _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "odin_fe_cmvn": (C.c_int, [_vp, _vp, _i32, _pi64, _i32, _vp, _i32, _i32, _i32, _i32, _vp]),
    "odin_gmm_create": (C.c_int, [_i32, _i32, C.POINTER(_vp)]),
    "odin_gmm_destroy": (None, [_vp]),
    "odin_gmm_stats_size": (_i64, [_vp, _i32]),
    "odin_gmm_set_params": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp]),
    "odin_gmm_estep": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _i32, _vp]),
    "odin_gmm_frames_create": (C.c_int, [_vp, _vp, _i64, C.POINTER(_vp), _vp]),
    "odin_gmm_frames_destroy": (None, [_vp]),
    "odin_gmm_estep_frames": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp]),
    "odin_gmm_allreduce": (C.c_int, [_vp, _vp, _vp, _vp]),
    "odin_gmm_mstep": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "odin_gmm_mixup": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp]),
    "odin_gmm_utt_stats": (C.c_int, [_vp, _vp, _vp, _pi64, _i32, _vp, _vp, _i32, _vp]),
    "odin_gmm_last_estep_ms": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                         C.POINTER(C.c_int32)]),
    "odin_gmm_last_estep_frames": (_i64, [_vp]),
    "odin_fe_last_run_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "odin_gmm_score": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
}

_lib = None


def load():
  """Loads the shared library (once).  Raises if it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.isfile(LIB_PATH):
    raise OdinError(ODIN_ENODEVICE,
                    "%s not found: build it with odin_b200/csrc/build.sh "
                    "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
  lib = C.CDLL(LIB_PATH)
  for name, (res, args) in SIGNATURES.items():
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args
  _lib = lib
  return lib


def check(rc):
  if rc < 0:
    raise OdinError(rc, load().odin_last_error().decode("utf-8", "replace"))
  return rc


def as_i64_ptr(arr):
  return arr.ctypes.data_as(_pi64)


def ptr(t):
  """Device/host pointer of a torch tensor (or None)."""
  return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
  import torch
  return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
  import torch
  if not torch.cuda.is_available():
    raise OdinError(ODIN_ENODEVICE, "no CUDA device: odin_b200 has no CPU fallback")
  load()
