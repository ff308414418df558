"""Oracle pinning for features (CPU): oracle/features.py against (1) the
reference's own _extract_fbank executed here when /root/reference is present,
(2) the committed golden vectors generated from it, (3) scipy's savgol_filter
(the function librosa >= 0.6.1 delegates delta to)."""
import numpy as np
import pytest
import scipy.signal

from conftest import golden
from oracle import features, ref_shim

CASES = ["cfg1_1s_16k", "ragged_16k", "crop_22k", "trunc_16k", "silence_16k"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", ["interp", "edge"])
def test_fbank_oracle_matches_golden(case, mode):
    g = golden("fbank_%s.npz" % case)
    sr, tmax = int(g["sr"]), int(g["tmax"])
    dm = features.DELTA_INTERP if mode == "interp" else features.DELTA_EDGE
    for i in range(len(g["n"])):
        feat, length = features.fbank(g["sig_%d" % i], sr, tmax, dm)
        assert length == int(g["len_%s_%d" % (mode, i)])
        assert length == features.fbank_num_frames(len(g["sig_%d" % i]), sr)
        np.testing.assert_allclose(feat, g["feat_%s_%d" % (mode, i)], rtol=0, atol=1e-11)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
@pytest.mark.parametrize("sr,n", [(16000, 16000), (16000, 48000), (22050, 30000), (8000, 8000)])
def test_fbank_oracle_matches_reference_run_here(sr, n):
    ap = ref_shim.load_reference_audioprocessor()
    rng = np.random.default_rng(n)
    sig = (0.1 * rng.standard_normal(n)).astype(np.float32)
    for mode, dm in (("interp", features.DELTA_INTERP), ("edge", features.DELTA_EDGE)):
        ref_shim.set_delta_mode(mode)
        ref, ref_len = ap.AudioProcessor(3510, "fbank").process_signal(sig, sr)
        got, got_len = features.fbank(sig, sr, 3510, dm)
        assert ref_len == got_len and ref.shape == got.shape
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-11)


def test_delta_interp_is_savgol():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((40, 57))
    want = scipy.signal.savgol_filter(x, 9, deriv=1, axis=-1, mode="interp", polyorder=1)
    np.testing.assert_allclose(features.delta(x, features.DELTA_INTERP), want, rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        features.delta(x[:, :8], features.DELTA_INTERP)


def test_frame_counts():
    # SURVEY appendix A
    for sec, frames in ((1, 98), (2, 198), (5, 498), (10, 998), (20, 1998)):
        assert features.fbank_num_frames(16000 * sec, 16000) == frames
    assert features.fbank_num_frames(220500, 22050) == 1000
    assert features.frame_params(22050) == (551, 220)
    assert features.mfcc_num_frames(16000, 16000) == 101


def test_mfcc_oracle_shape_and_dct():
    rng = np.random.default_rng(0)
    sig = (0.1 * rng.standard_normal(16000)).astype(np.float32)
    out, T = features.mfcc(sig, 16000, 3510)
    assert out.shape == (101, 20) and T == 101
    basis = features.dct_ortho(128, 128)
    np.testing.assert_allclose(basis @ basis.T, np.eye(128), atol=1e-12)
    w = features.mel_filterbank_slaney(16000, 400)
    assert w.shape == (128, 201) and (w >= 0).all()
