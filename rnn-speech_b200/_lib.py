"""ctypes binding of librnnspeech_b200.so (the C ABI in include/rnnspeech_b200.h).

There is no CPU fallback: importing this module fails loudly when the shared
library has not been built, and every compute entry point needs a CUDA device.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librnnspeech_b200.so")

RS_OK = 0
DELTA_INTERP, DELTA_EDGE = 0, 1
PCM_F32, PCM_S16 = 0, 1
CTC_BETA_SOURCE, CTC_BETA_DEST = 0, 1
FBANK_DIM = 120


class RnnSpeechError(RuntimeError):
    """Raised when a C-ABI call returns a negative status."""

    def __init__(self, code, message):
        super().__init__("rnnspeech_b200 error %d: %s" % (code, message))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "librnnspeech_b200.so is missing (%s). Build it with `python rnn-speech_b200/build.py` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    return ctypes.CDLL(LIB_PATH)


_lib = _load()

# name -> (restype, argtypes); mirrors include/rnnspeech_b200.h one for one
SIGNATURES = {
    "rs_version": (c_int, []),
    "rs_last_error": (c_char_p, []),
    "rs_sm_count": (c_int, []),
    "rs_launch_count": (c_uint64, []),
    "rs_crc32c": (ctypes.c_uint32, [c_void_p, c_size_t, ctypes.c_uint32]),
    "rs_am_set_params_version": (c_int, [c_void_p, c_uint64]),
    "rs_am_enable_timing": (c_int, [c_void_p, c_int]),
    "rs_am_set_debug_timeline": (c_int, [c_void_p, c_void_p, c_void_p]),
    "rs_am_recurrent_ms": (c_int, [c_void_p, c_int, c_int, POINTER(c_float)]),
    "rs_am_recurrent_trace": (c_int, [c_void_p, c_int, c_int, POINTER(c_float), c_int]),
    "rs_fbank_workspace_bytes": (c_size_t, [c_int, c_int64, c_int]),
    "rs_fbank_num_frames": (c_int64, [c_int64, c_int]),
    "rs_fbank_tables_host": (c_int, [c_int, POINTER(c_float), POINTER(c_float), POINTER(c_int), POINTER(c_int)]),
    "rs_fbank_forward": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "rs_mfcc_workspace_bytes": (c_size_t, [c_int, c_int64, c_int]),
    "rs_mfcc_num_frames": (c_int64, [c_int64, c_int]),
    "rs_mfcc_forward": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_int, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_void_p]),
    "rs_resample_workspace_bytes": (c_size_t, [c_int, c_int64]),
    "rs_resample_num_samples": (c_int64, [c_int64, c_int, c_int]),
    "rs_resample_filter_host": (c_int, [POINTER(c_double), POINTER(c_int)]),
    "rs_resample_forward": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int64, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_size_t, c_void_p]),
    "rs_flac_decode_host": (c_int, [c_void_p, c_size_t, c_void_p, c_int64, POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                    POINTER(c_int64), c_void_p]),
    "rs_pcm16_to_f32": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "rs_pcm_f32_to_mono": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "rs_am_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int]),
    "rs_am_destroy": (None, [c_void_p]),
    "rs_am_param_count": (c_int64, [c_void_p]),
    "rs_am_param_offset": (c_int64, [c_void_p, c_int, c_int]),
    "rs_am_uses_tensor_cores": (c_int, [c_void_p]),
    "rs_am_set_normalization": (c_int, [c_void_p, c_int]),
    "rs_am_reserve_bytes": (c_size_t, [c_void_p]),
    "rs_am_workspace_bytes": (c_size_t, [c_void_p]),
    "rs_am_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_float, c_float,
                              c_uint64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rs_am_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_uint64, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rs_ctc_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "rs_ctc_loss_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rs_ctc_greedy_decode": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rs_ctc_beam_workspace_bytes": (c_size_t, [c_int, c_int]),
    "rs_ctc_beam_search": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "rs_edit_distance": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rs_accumulate_mean": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "rs_memset_zero": (c_int, [c_void_p, c_size_t, c_void_p]),
    "rs_memcpy_h2d_async": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "rs_comm_unique_id": (c_int, [c_void_p, c_size_t]),
    "rs_comm_init": (c_int, [POINTER(c_void_p), c_void_p, c_size_t, c_int, c_int]),
    "rs_allreduce_sum": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "rs_comm_destroy": (None, [c_void_p]),
    "rs_sumsq": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "rs_clip_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float,
                                  c_float, c_float, c_float, c_int64, c_void_p]),
}

# Diagnostic hooks (self-tests, micro-benchmarks): only in librnnspeech_b200_diag.so, loaded on demand by the tests.
DIAG_LIB_PATH = os.path.join(_HERE, "librnnspeech_b200_diag.so")
DIAG_SIGNATURES = {
    "rs_tc_selftest": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "rs_tc_mma_bench": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rs_tc_ts_selftest": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rs_gemm_tc_test": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t,
                                c_void_p]),
    "rs_gemm_tc_bench": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_void_p, c_size_t, POINTER(c_float), c_void_p]),
}
_diag = None


def diag():
    """The diagnostic build of the library (the product library + the RS_DIAG hooks)."""
    global _diag
    if _diag is None:
        if not os.path.exists(DIAG_LIB_PATH):
            raise ImportError("librnnspeech_b200_diag.so is missing (python rnn-speech_b200/build.py)")
        lib = ctypes.CDLL(DIAG_LIB_PATH)
        for name, (res, args) in DIAG_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        lib.rs_last_error.restype = c_char_p
        _diag = lib
    return _diag


def diag_call(name, *args):
    """Call a status-returning diagnostic hook and raise on failure."""
    lib = diag()
    code = getattr(lib, name)(*args)
    if code != RS_OK:
        msg = lib.rs_last_error()
        raise RnnSpeechError(code, msg.decode("utf-8", "replace") if msg else "")
    return code


for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(_lib, _name)          # AttributeError here == symbol missing from the .so
    _fn.restype = _res
    _fn.argtypes = _args


def last_error():
    msg = _lib.rs_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(code):
    if code != RS_OK:
        err = last_error()
        if code == -1:
            raise ValueError("rnnspeech_b200: " + err)      # the reference raises ValueError / InvalidArgumentError
        raise RnnSpeechError(code, err)
    return code


def call(name, *args):
    """Call a status-returning entry point and raise on failure."""
    return check(getattr(_lib, name)(*args))


def raw(name):
    return getattr(_lib, name)
