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


def test_mfcc_restatement_against_torchaudio():
    """librosa is absent, torchaudio is not: its MFCC transform with librosa's settings (centred reflect-padded
    frames of n_fft = 0.025 sr, periodic Hann, 128 Slaney mel bands with Slaney normalisation, power_to_db with
    top_db 80, orthonormal DCT-II, 20 coefficients) is an independent implementation of what
    util/audioprocessor.py:63-75 calls; oracle/features.py::mfcc agrees with it to float32 resolution."""
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    rng = np.random.default_rng(0)
    for sr, n in ((16000, 16000), (16000, 9000), (22050, 22050)):
        sig = (0.1 * rng.standard_normal(n)).astype(np.float32)
        got = features.mfcc(sig, sr, 10000)
        got = got[0] if isinstance(got, tuple) else got
        tr = torchaudio.transforms.MFCC(
            sample_rate=sr, n_mfcc=20, dct_type=2, norm="ortho", log_mels=False,
            melkwargs=dict(n_fft=int(round(0.025 * sr)), hop_length=int(round(0.01 * sr)), n_mels=128, f_min=0.0,
                           f_max=sr / 2, center=True, pad_mode="reflect", power=2.0, norm="slaney",
                           mel_scale="slaney", window_fn=torch.hann_window))
        want = tr(torch.from_numpy(sig)).numpy().T
        assert want.shape == got.shape
        assert np.abs(got - want).max() < 2e-5 * np.abs(want).max() + 1e-3


def test_delta_edge_mode_is_the_replicate_padded_fir():
    """DELTA_EDGE (librosa <= 0.6.0) is the width-9 regression FIR over edge-replicated frames; torchaudio's
    compute_deltas(mode='replicate') is that filter with the textbook 1/60 normalisation, old librosa divided by
    sum|w| = 20: the two differ by exactly the factor 3."""
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    rng = np.random.default_rng(1)
    x = rng.standard_normal((40, 57))
    got = features.delta(x, mode=features.DELTA_EDGE)
    want = torchaudio.functional.compute_deltas(torch.from_numpy(x), win_length=9, mode="replicate").numpy() * 3.0
    np.testing.assert_allclose(got, want, atol=1e-12)
