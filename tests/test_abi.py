"""The C-ABI shared library loads on a CPU-only box and exports every symbol
that include/rnnspeech_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import numpy as np

from conftest import ROOT


def _declared_symbols(diag=False):
    """Symbols the header declares: the product ABI, or (diag=True) the hooks inside its #ifdef RS_DIAG block."""
    text = open(os.path.join(ROOT, "include", "rnnspeech_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    m = re.search(r"#ifdef RS_DIAG(.*?)#endif", text, flags=re.S)
    block = m.group(1) if m else ""
    if diag:
        text = block
    elif m:
        text = text[:m.start()] + text[m.end():]
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "missing export: " + name


def test_ctypes_table_matches_header(pkg):
    assert sorted(pkg._lib.SIGNATURES) == _declared_symbols()


def test_diagnostic_hooks_live_in_their_own_library(pkg):
    """The self-test / micro-benchmark hooks are behind -DRS_DIAG: absent from the product library, present in
    librnnspeech_b200_diag.so (which also carries the whole product ABI)."""
    names = _declared_symbols(diag=True)
    assert sorted(pkg._lib.DIAG_SIGNATURES) == names and len(names) >= 5
    product = ctypes.CDLL(pkg.LIB_PATH)
    diag = pkg._lib.diag()
    for name in names:
        assert not hasattr(product, name), "diagnostic hook exported by the product library: " + name
        assert hasattr(diag, name)
    for name in _declared_symbols():
        assert hasattr(diag, name)


def test_version_and_error_string(pkg):
    assert pkg._lib.raw("rs_version")() >= 100
    assert isinstance(pkg._lib.last_error(), str)


def test_host_only_entry_points(pkg):
    lib = pkg._lib
    for sec, frames in ((1, 98), (10, 998), (20, 1998)):
        assert lib.raw("rs_fbank_num_frames")(16000 * sec, 16000) == frames
    assert lib.raw("rs_fbank_num_frames")(220500, 22050) == 1000
    assert lib.raw("rs_mfcc_num_frames")(16000, 16000) == 101
    assert lib.raw("rs_fbank_workspace_bytes")(32, 160000, 16000) > 32 * 998 * 40 * 4
    assert lib.raw("rs_ctc_workspace_bytes")(998, 32, 80, 120) > 2 * 998 * 32 * 241 * 4


def test_fbank_tables_match_oracle(pkg):
    from oracle import features
    melw = (ctypes.c_float * (40 * 257))()
    win = (ctypes.c_float * 512)()
    fl, fs = ctypes.c_int(), ctypes.c_int()
    for sr in (8000, 16000, 22050, 44100):
        pkg._lib.call("rs_fbank_tables_host", sr, melw, win, ctypes.byref(fl), ctypes.byref(fs))
        assert (fl.value, fs.value) == features.frame_params(sr)
        fb, _ = features.mel_filterbank_htk(sr)
        np.testing.assert_allclose(np.frombuffer(melw, np.float32).reshape(40, 257), fb, atol=1e-7)
        n = min(512, fl.value)
        np.testing.assert_allclose(np.frombuffer(win, np.float32)[:n], np.hamming(fl.value)[:n], atol=1e-7)


def test_invalid_arguments_raise_without_a_gpu(pkg):
    import pytest
    h = ctypes.c_void_p()
    with pytest.raises(ValueError):
        pkg._lib.call("rs_am_create", ctypes.byref(h), 0, 128, 120, 80, 2, 100)
    with pytest.raises(pkg.RnnSpeechError):
        pkg._lib.call("rs_am_create", ctypes.byref(h), 1, 100000, 120, 80, 2, 100)
    pkg._lib.call("rs_am_create", ctypes.byref(h), 3, 768, 120, 80, 32, 1000)
    assert pkg._lib.raw("rs_am_param_count")(h) == 14319440
    assert pkg._lib.raw("rs_am_param_offset")(h, 0, 0) == 0
    assert pkg._lib.raw("rs_am_param_offset")(h, 2, 1) == 120 * 768 + 768 + 2 * 768 * 4 * 768 + 4 * 768
    pkg._lib.raw("rs_am_destroy")(h)
    with pytest.raises(ValueError):
        pkg.AudioProcessor(100, "spectrogram")
