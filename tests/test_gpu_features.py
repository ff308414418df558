"""GPU parity: CUDA fbank kernels (through the C ABI) against the golden vectors
produced by the REFERENCE's own _extract_fbank and against the oracle.

Tolerance: the reference computes in float64 after a float32 pre-emphasis and so do
the kernels (fp64 FFT / mel / log10 / deltas); the result is rounded to float32 once,
as the reference does at models/AcousticModel.py:816.  |diff| <= 1e-4 on features
whose static part is in dB (|value| up to ~160, float32 ulp there is 1.5e-5).
"""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import features

pytestmark = pytest.mark.gpu
ATOL = 1e-4
CASES = ["cfg1_1s_16k", "ragged_16k", "crop_22k", "trunc_16k", "silence_16k"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode", ["interp", "edge"])
def test_fbank_matches_reference_golden(pkg, cuda, case, mode):
    g = golden("fbank_%s.npz" % case)
    sr, tmax = int(g["sr"]), int(g["tmax"])
    sigs = [g["sig_%d" % i] for i in range(len(g["n"]))]
    ap = pkg.AudioProcessor(tmax, "fbank", delta_mode=mode, device=cuda)
    for time_major in (False, True):
        feats, nframes = ap.process_batch(sigs, sr, time_major=time_major)
        feats = feats.cpu().numpy()
        if time_major:
            feats = feats.transpose(1, 0, 2)
        nframes = nframes.cpu().numpy()
        for i in range(len(sigs)):
            want = g["feat_%s_%d" % (mode, i)]
            assert nframes[i] == int(g["len_%s_%d" % (mode, i)])       # pre-truncation length
            keep = want.shape[0]
            assert keep == min(nframes[i], tmax)
            np.testing.assert_allclose(feats[i, :keep], want, rtol=0, atol=ATOL)
            assert np.all(feats[i, keep:] == 0)                         # padded_batch zero fill


def test_process_signal_api(pkg, cuda):
    g = golden("fbank_trunc_16k.npz")
    ap = pkg.AudioProcessor(int(g["tmax"]), "fbank", device=cuda)
    feat, length = ap.process_signal(g["sig_0"], int(g["sr"]))
    assert feat.shape == (50, 120) and length == 98 and feat.dtype == np.float32
    np.testing.assert_allclose(feat, g["feat_interp_0"], atol=ATOL)
    assert ap.feature_size == 120
    with pytest.raises(ValueError):
        ap.process_signal(np.zeros(1500, np.float32), 16000)            # 7 frames < delta width 9


def test_fbank_full_size_batch_against_oracle(pkg, cuda):
    """BASELINE config 2 feature shape: 32 utterances of 10 s @16 kHz."""
    rng = np.random.default_rng(0)
    sigs = [(0.1 * rng.standard_normal(160000)).astype(np.float32) for _ in range(32)]
    ap = pkg.AudioProcessor(1000, "fbank", device=cuda)
    feats, nframes = ap.process_batch(sigs, 16000, time_major=True)
    assert tuple(feats.shape) == (1000, 32, 120)
    assert np.all(nframes.cpu().numpy() == 998)
    feats = feats.cpu().numpy()
    worst = 0.0
    for b in (0, 7, 31):
        want, _ = features.fbank(sigs[b], 16000, 1000)
        worst = max(worst, float(np.abs(feats[:998, b] - want).max()))
        np.testing.assert_allclose(feats[:998, b], want, rtol=0, atol=ATOL)
    assert np.all(feats[998:] == 0)
    # size-independent properties: mean-normalised static part, deterministic re-run
    assert np.abs(feats[:998, :, :40].mean(axis=0)).max() < 1e-3
    again, _ = ap.process_batch(sigs, 16000, time_major=True)
    assert torch.equal(again, torch.from_numpy(feats).to(again.device))
    print("fbank cfg-2 max |diff| vs oracle: %.3e" % worst)


def test_fbank_linearity_property(pkg, cuda):
    """Scaling the PCM by a shifts every log-mel by 20 log10(a): after mean
    normalisation all 120 features are unchanged."""
    rng = np.random.default_rng(1)
    sig = (0.1 * rng.standard_normal(48000)).astype(np.float32)
    ap = pkg.AudioProcessor(3510, "fbank", device=cuda)
    a, _ = ap.process_batch([sig], 16000)
    b, _ = ap.process_batch([sig * 4.0], 16000)
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), atol=2e-4)


@pytest.mark.parametrize("sr,lens", [(16000, [16000, 9000]), (22050, [22050])])
def test_mfcc_matches_oracle(pkg, cuda, sr, lens):
    """MFCC-20 (librosa.feature.mfcc semantics restated in oracle/features.py; parity unpinned
    upstream).  fp32 kernels vs the float64 oracle: |diff| <= 2e-3 on coefficients of magnitude
    up to several hundred (dB-scaled log-mel through an orthonormal DCT)."""
    rng = np.random.default_rng(sr)
    sigs = [(0.1 * rng.standard_normal(n)).astype(np.float32) + 0.2 * np.sin(np.arange(n) * 0.05).astype(np.float32)
            for n in lens]
    ap = pkg.AudioProcessor(3510, "mfcc", device=cuda)
    assert ap.feature_size == 20
    feats, nframes = ap.process_batch(sigs, sr, time_major=False)
    feats, nframes = feats.cpu().numpy(), nframes.cpu().numpy()
    for i, s in enumerate(sigs):
        want, T = features.mfcc(s, sr, 3510)
        assert nframes[i] == T == 1 + len(s) // features.frame_params(sr)[1]
        err = np.abs(feats[i, :T] - want).max()
        print("mfcc sr=%d n=%d: max |diff| %.2e (max |coef| %.1f)" % (sr, len(s), err, np.abs(want).max()))
        assert err < 2e-3
        assert np.all(feats[i, T:] == 0)
    one, length = ap.process_signal(sigs[0], sr)
    assert one.shape == (length, 20) and length == nframes[0]
    short = pkg.AudioProcessor(40, "mfcc", device=cuda)
    f40, l40 = short.process_signal(sigs[0], sr)
    assert f40.shape == (40, 20) and l40 == nframes[0]          # truncated features, untruncated length


def test_mfcc_40_coefficients_baseline_config_1(pkg, cuda):
    """BASELINE config 1 as it is written ("40-dim MFCC", batch 2 of 1 s at 16 kHz): the reference's code keeps
    librosa's 20 coefficients (util/audioprocessor.py:21,65-66) while its README says 40; AudioProcessor(n_mfcc=40)
    runs the latter.  The first 20 of the 40 are the 20-coefficient result (a DCT truncation)."""
    rng = np.random.default_rng(40)
    sigs = [(0.1 * rng.standard_normal(16000)).astype(np.float32) for _ in range(2)]
    ap40 = pkg.AudioProcessor(200, "mfcc", device=cuda, n_mfcc=40)
    ap20 = pkg.AudioProcessor(200, "mfcc", device=cuda)
    assert ap40.feature_size == 40
    f40, n40 = ap40.process_batch(sigs, 16000, time_major=True)
    f20, _ = ap20.process_batch(sigs, 16000, time_major=True)
    assert tuple(f40.shape) == (200, 2, 40) and int(n40[0]) == 101
    f40, f20 = f40.cpu().numpy(), f20.cpu().numpy()
    for b, s in enumerate(sigs):
        want, T = features.mfcc(s, 16000, 200, n_mfcc=40)
        assert np.abs(f40[:T, b] - want).max() < 2e-3
    np.testing.assert_allclose(f40[:, :, :20], f20, atol=1e-5)
    with pytest.raises(ValueError):
        pkg.AudioProcessor(200, "mfcc", n_mfcc=0)
