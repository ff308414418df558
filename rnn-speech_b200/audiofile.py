"""Host-side audio front door ('next').

The reference calls librosa.load(file, mono=True) (util/audioprocessor.py:49):
decode, down-mix, resample to 22 050 Hz with resampy 'kaiser_best'.  librosa is
absent; this reads RIFF/WAV with the standard library and resamples with
scipy.signal.resample_poly -- close to, but not bit-identical with, the
reference's resampler (parity unpinned; documented in DESIGN.md).
"""
import wave
from fractions import Fraction

import numpy as np

TARGET_SR = 22050


def read_wav(path):
    with wave.open(path, "rb") as w:
        nch, width, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError("unsupported WAV sample width %d" % width)
    if nch > 1:
        x = x.reshape(-1, nch).mean(axis=1)
    return x, sr


def load_audio(path, sr=TARGET_SR):
    if not str(path).lower().endswith(".wav"):
        raise NotImplementedError("only RIFF/WAV decoding is built in (got %r); FLAC/MP3 need an external decoder" % (path,))
    x, file_sr = read_wav(path)
    if sr is not None and file_sr != sr:
        from scipy.signal import resample_poly
        frac = Fraction(sr, file_sr)
        x = resample_poly(x, frac.numerator, frac.denominator).astype(np.float32)
        file_sr = sr
    return x, file_sr
