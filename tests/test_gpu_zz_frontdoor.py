"""GPU parity: the audio front door (csrc/resample.cu through the C ABI) against oracle/resample.py -- the
restatement of librosa.load's int16 -> float32 -> mono -> resampy 'kaiser_best' -> fix_length chain
(util/audioprocessor.py:49) -- and the file path end to end (decode on the host, everything else on the device).

Tolerance: the kernel accumulates each output sample in float64 and rounds once; resampy (and the oracle) round the
running sum to float32 after every tap (up to 2 x 64 / min(1, ratio) taps), so the two differ by a few float32 ulps
of the partial sums: |diff| <= 2e-6 for signals of amplitude <= 1.  The int16 -> float32 / mono step is bit-exact.
"""
import io
import wave

import numpy as np
import pytest
import torch

import flac_writer
from oracle import features
from oracle import resample as R

pytestmark = pytest.mark.gpu
ATOL = 2e-6


def _resample_gpu(pkg, dev, sigs, sr_in, sr_out, fmt="f32", channels=1):
    lib = pkg._lib
    frames = [len(s) // channels for s in sigs]
    in_off = np.zeros(len(sigs) + 1, np.int64)
    np.cumsum(frames, out=in_off[1:])
    out_len = [int(lib.raw("rs_resample_num_samples")(n, sr_in, sr_out)) for n in frames]
    out_off = np.zeros(len(sigs) + 1, np.int64)
    np.cumsum(out_len, out=out_off[1:])
    src = torch.from_numpy(np.concatenate(sigs)).to(dev)
    in_d, out_d = torch.from_numpy(in_off).to(dev), torch.from_numpy(out_off).to(dev)
    out = torch.full((int(out_off[-1]) + 8,), 777.0, dtype=torch.float32, device=dev)     # + guard
    ws = torch.empty((int(lib.raw("rs_resample_workspace_bytes")(len(sigs), max(out_len))),), dtype=torch.uint8,
                     device=dev)
    lib.call("rs_resample_forward", src.data_ptr(), lib.PCM_S16 if fmt == "s16" else lib.PCM_F32, channels,
             in_d.data_ptr(), len(sigs), max(out_len), sr_in, sr_out, out.data_ptr(), out_d.data_ptr(),
             ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    assert np.all(out[int(out_off[-1]):] == 777.0)                                       # nothing past the end
    return [out[out_off[i]:out_off[i + 1]] for i in range(len(sigs))]


@pytest.mark.parametrize("sr_in,sr_out,lens", [
    (16000, 22050, [16000, 3001, 1]),          # the LibriSpeech case; ragged batch; a one-sample utterance
    (8000, 22050, [4000, 777]),
    (44100, 22050, [22050, 5000]),             # exact 2:1
    (48000, 22050, [24000, 7001]),             # truncated table stride + time-register rounding (see resample.cu)
    (22050, 16000, [11025, 2500]),
    (16000, 22050, [160000]),                  # 10 s: many blocks per utterance
    (11025, 22050, [3000, 17]),                # phase period 2 (shared-weights kernel, S = 256)
    (16000, 22051, [5000]),                    # coprime rates: period > 1024, per-output kernel without the replay
])
def test_resample_matches_oracle(pkg, cuda, sr_in, sr_out, lens):
    rng = np.random.default_rng(len(lens) + sr_in)
    sigs = []
    for n in lens:
        t = np.arange(n) / float(sr_in)
        sigs.append((0.3 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 2500 * t) +
                     0.1 * rng.standard_normal(n)).astype(np.float32))
    got = _resample_gpu(pkg, cuda, sigs, sr_in, sr_out)
    for x, y in zip(sigs, got):
        want = R.resample_kaiser_best(x, sr_in, sr_out)
        assert y.shape == want.shape
        np.testing.assert_allclose(y, want, rtol=0, atol=ATOL)


@pytest.mark.parametrize("channels", [1, 2, 3])
def test_int16_interleaved_input(pkg, cuda, channels):
    rng = np.random.default_rng(channels)
    pcm = [rng.integers(-20000, 20000, size=n * channels).astype(np.int16) for n in (5000, 1234)]
    got = _resample_gpu(pkg, cuda, pcm, 16000, 22050, fmt="s16", channels=channels)
    for x, y in zip(pcm, got):
        np.testing.assert_allclose(y, R.load(x, channels, 16000), rtol=0, atol=ATOL)
    # no rate change: conversion + down-mix only, bit-exact
    src = torch.from_numpy(pcm[0]).to(cuda)
    out = torch.empty((5000,), dtype=torch.float32, device=cuda)
    pkg._lib.call("rs_pcm16_to_f32", src.data_ptr(), 5000, channels, out.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
    np.testing.assert_array_equal(out.cpu().numpy(), R.pcm16_to_float_mono(pcm[0], channels))


def test_resample_argument_errors(pkg, cuda):
    lib = pkg._lib
    x = torch.zeros(100, device=cuda)
    off = torch.tensor([0, 100], dtype=torch.int64, device=cuda)
    ws = torch.empty((int(lib.raw("rs_resample_workspace_bytes")(1, 138)),), dtype=torch.uint8, device=cuda)
    args = lambda **k: (x.data_ptr(), k.get("fmt", 0), k.get("ch", 1), off.data_ptr(), 1, 138, k.get("a", 16000),
                        k.get("b", 22050), x.data_ptr(), off.data_ptr(), ws.data_ptr(), k.get("ws", ws.numel()), None)
    with pytest.raises(ValueError):
        lib.call("rs_resample_forward", *args(a=16000, b=16000))
    with pytest.raises(ValueError):
        lib.call("rs_resample_forward", *args(fmt=7))
    with pytest.raises(ValueError):
        lib.call("rs_resample_forward", *args(ch=9))                  # 1..8 channels
    with pytest.raises(pkg.RnnSpeechError):
        lib.call("rs_resample_forward", *args(ws=16))


def _write_wav(path, x, sr):
    x = np.asarray(x)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1 if x.ndim == 1 else x.shape[1])
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(x.astype("<i2").tobytes())


def test_process_audio_file_is_librosa_load_then_process_signal(pkg, cuda, tmp_path):
    """util/audioprocessor.py:41-50 on 16 kHz files: load -> 22 050 Hz -> fbank with frame 551 / hop 220 cropped to
    512 (the reference's behaviour at that rate), checked against oracle resample + oracle fbank."""
    rng = np.random.default_rng(5)
    n = 24000
    t = np.arange(n) / 16000.0
    mono = (6000 * np.sin(2 * np.pi * 300 * t) + 3000 * np.sin(2 * np.pi * 1800 * t) +
            800 * rng.standard_normal(n)).astype(np.int16)
    stereo = np.stack([mono, (0.25 * mono).astype(np.int16)], 1)
    _write_wav(tmp_path / "m.wav", mono, 16000)
    (tmp_path / "s.flac").write_bytes(flac_writer.encode(stereo, 16000, blocksize=1152))
    ap = pkg.AudioProcessor(3510, "fbank", delta_mode="interp", device=cuda)
    from rnn_speech_b200 import audiofile
    for name, pcm, ch in (("m.wav", mono, 1), ("s.flac", stereo.reshape(-1), 2)):
        sig, sr = audiofile.load_audio(str(tmp_path / name), device=cuda)          # librosa.load(file, mono=True)
        assert sr == 22050 and sig.dtype == np.float32
        np.testing.assert_allclose(sig, R.load(pcm, ch, 16000), rtol=0, atol=ATOL)
        feat, length = ap.process_audio_file(str(tmp_path / name))
        # features of the device-resampled signal (the bands above the old Nyquist hold only rounding residue, so
        # the oracle is given the same signal rather than its own resampling of it)
        want, want_len = features.fbank(sig, 22050, delta_mode=features.DELTA_INTERP)
        assert length == want_len and feat.shape == want.shape
        np.testing.assert_allclose(feat, want.astype(np.float32), rtol=0, atol=1e-4)
    # the batch entry point keeps everything on the device and honours a target rate equal to the file's
    feats, nframes = ap.process_audio_files([str(tmp_path / "m.wav"), str(tmp_path / "s.flac")], time_major=True, sr=16000)
    assert feats.shape == (3510, 2, 120) and feats.is_cuda
    want, want_len = features.fbank(R.pcm16_to_float_mono(mono, 1), 16000, delta_mode=features.DELTA_INTERP)
    assert int(nframes[0]) == want_len
    np.testing.assert_allclose(feats[:want.shape[0], 0].cpu().numpy(), want.astype(np.float32), rtol=0, atol=1e-4)


def test_dataset_over_audio_files_matches_direct_extraction(pkg, cuda, tmp_path):
    """AcousticModel.build_dataset over file names (models/AcousticModel.py:801-840): each mini-batch is decoded on the
    prefetch thread and converted / resampled / featurised on its side stream; the result must be what
    process_audio_files gives for the same files, the last batch padded with zero features / length 0 (:147-152)."""
    rng = np.random.default_rng(11)
    names, texts = [], ["it'll do", "coffee", "we've"]
    for i, n in enumerate((12000, 9000, 15000)):
        pcm = (4000 * np.sin(np.arange(n) * (0.02 + 0.01 * i)) + 500 * rng.standard_normal(n)).astype(np.int16)
        path = tmp_path / ("u%d.%s" % (i, "flac" if i % 2 else "wav"))
        if i % 2:
            path.write_bytes(flac_writer.encode(pcm, 16000))
        else:
            _write_wav(path, pcm, 16000)
        names.append(str(path))
    Tmax, B = 160, 2
    ds = pkg.AcousticModel.build_dataset([[n, t, None] for n, t in zip(names, texts)], B, Tmax, 600, "fbank",
                                         pkg.ENGLISH_CHAR_MAP, device=cuda)
    ap = pkg.AudioProcessor(Tmax, "fbank", device=cuda)
    want, want_len = ap.process_audio_files(names, time_major=True)
    torch.cuda.synchronize()
    batches = list(ds)
    assert len(batches) == 2
    got = torch.cat([b[0] for b in batches], dim=1)
    lens = torch.cat([b[1] for b in batches])
    assert got.shape == (Tmax, 4, 120) and lens.shape == (4,)
    np.testing.assert_array_equal(lens.cpu().numpy(), list(want_len.cpu().numpy()) + [0])
    np.testing.assert_allclose(got[:, :3].cpu().numpy(), want.cpu().numpy(), rtol=0, atol=1e-4)
    assert bool((got[:, 3] == 0).all())
    assert batches[0][2].shape[0] == 2 and batches[1][2].shape[0] == 1
    assert list(batches[1][2][0]) == pkg.get_str_labels(pkg.ENGLISH_CHAR_MAP, "we've")


def test_float_interleaved_input(pkg, cuda):
    """Sample formats other than 16-bit reach the device as scaled float32, still interleaved: same mono mix."""
    rng = np.random.default_rng(9)
    x = (0.2 * rng.standard_normal(4000 * 2)).astype(np.float32)
    mono = x.reshape(-1, 2).mean(axis=1, dtype=np.float32)
    got = _resample_gpu(pkg, cuda, [x], 16000, 22050, fmt="f32", channels=2)[0]
    np.testing.assert_allclose(got, R.resample_kaiser_best(mono, 16000, 22050), rtol=0, atol=ATOL)
    got = _resample_gpu(pkg, cuda, [x], 44100, 22050, fmt="f32", channels=2)[0]
    np.testing.assert_allclose(got, R.resample_kaiser_best(mono, 44100, 22050), rtol=0, atol=ATOL)
    src = torch.from_numpy(x).to(cuda)
    out = torch.empty((4000,), dtype=torch.float32, device=cuda)
    pkg._lib.call("rs_pcm_f32_to_mono", src.data_ptr(), 4000, 2, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    np.testing.assert_array_equal(out.cpu().numpy(), mono)
