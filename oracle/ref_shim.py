"""Oracle tooling (test infrastructure, NOT product code).

Imports the reference's own ``util/audioprocessor.py`` from /root/reference in
THIS container, with a stand-in ``librosa`` module that provides only
``feature.delta`` (librosa is not installed and cannot be; no network).  The
reference's ``_extract_fbank`` (util/audioprocessor.py:77-161) is pure numpy
apart from that one call, so everything except ``delta`` is the reference's own
code executing.

``delta`` in the stand-in:
  * mode 'interp' -> scipy.signal.savgol_filter, i.e. the exact call librosa
    >= 0.6.1 makes (the third-party body itself, scipy is installed);
  * mode 'edge'   -> restated librosa <= 0.6.0 behaviour (np.pad 'edge' +
    scipy.signal.lfilter), from memory: unpinned.

/root/reference does not exist on the GPU box: this module is only used by
oracle/gen_golden.py (run here, vectors committed under tests/golden/) and by
``-m "not gpu"`` tests that skip when the reference is absent.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.signal

REFERENCE_ROOT = os.environ.get("RNNSPEECH_REFERENCE", "/root/reference")

_delta_mode = "interp"


def set_delta_mode(mode):
    global _delta_mode
    assert mode in ("interp", "edge")
    _delta_mode = mode


def _delta(data, width=9, order=1, axis=-1, **kwargs):
    data = np.atleast_1d(data)
    if _delta_mode == "interp":
        return scipy.signal.savgol_filter(data, width, deriv=order, axis=axis, mode="interp", polyorder=order)
    half_length = 1 + int(width // 2)
    window = np.arange(half_length - 1.0, -half_length, -1.0)
    window /= np.sum(np.abs(window))
    padding = [(0, 0)] * data.ndim
    padding[axis] = (width, width)
    delta_x = np.pad(data, padding, mode="edge")
    for _ in range(order):
        delta_x = scipy.signal.lfilter(window, 1, delta_x, axis=axis)
    idx = [slice(None)] * delta_x.ndim
    idx[axis] = slice(-half_length - data.shape[axis], -half_length)
    return delta_x[tuple(idx)]


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "util", "audioprocessor.py"))


def load_reference_audioprocessor():
    """Returns the reference's util.audioprocessor module (or raises if the
    reference tree is absent)."""
    if not available():
        raise FileNotFoundError("reference tree not found at %s" % REFERENCE_ROOT)
    fake = types.ModuleType("librosa")
    fake.feature = types.ModuleType("librosa.feature")
    fake.feature.delta = _delta

    def _absent(*a, **k):
        raise RuntimeError("librosa is not installed; only feature.delta is shimmed")
    fake.feature.mfcc = _absent
    fake.load = _absent
    saved = {k: sys.modules.get(k) for k in ("librosa", "librosa.feature")}
    sys.modules["librosa"] = fake
    sys.modules["librosa.feature"] = fake.feature
    try:
        spec = importlib.util.spec_from_file_location(
            "_reference_audioprocessor", os.path.join(REFERENCE_ROOT, "util", "audioprocessor.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load_reference_dataprocessor():
    """Returns the reference's util.dataprocessor module (label codec) with a
    stand-in for the absent ``mutagen`` import."""
    if not available():
        raise FileNotFoundError("reference tree not found at %s" % REFERENCE_ROOT)
    saved = {}
    for name in ("mutagen", "mutagen.mp3", "mutagen.flac", "mutagen.wave"):
        saved[name] = sys.modules.get(name)
        sys.modules[name] = types.ModuleType(name)
    try:
        spec = importlib.util.spec_from_file_location(
            "_reference_dataprocessor", os.path.join(REFERENCE_ROOT, "util", "dataprocessor.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
