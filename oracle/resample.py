"""Oracle (test infrastructure, NOT product code): CPU restatement of the audio
front door the reference reaches through ``librosa.load(file_name, mono=True)``
(/root/reference/util/audioprocessor.py:49).

librosa and resampy are third-party, absent from /root/reference and from this
image, and unpinned (requirements.txt:1-5): everything here is restated from
their published algorithms -- PARITY UNPINNED upstream.  Independent anchor
used by tests/test_oracle_resample.py: torchaudio's Kaiser-windowed sinc
resampler run with resampy's 'kaiser_best' constants (a different formulation
of the same filter, so agreement is to ~1e-3 of full scale, not bit level).

librosa.load(path, sr=22050, mono=True, res_type='kaiser_best') does:
  1. decode to float32: int16 -> x * (1/32768)          (librosa.util.buf_to_float)
  2. mono: np.mean over channels                         (librosa.core.to_mono)
  3. resampy.resample(y, sr_native, 22050, filter='kaiser_best')
  4. librosa.util.fix_length(y_hat, ceil(n * ratio))     (zero-pads the one missing sample)
  5. returns float32
"""
import numpy as np

TARGET_SR = 22050                       # librosa.load default sr

# resampy's shipped 'kaiser_best' filter = sinc_window(num_zeros=64, precision=9,
# window=kaiser(beta), rolloff) with these two optimised constants (beta is also
# torchaudio's default for its Kaiser resampler, functional.py: 14.769656459379492)
KAISER_BEST_BETA = 14.769656459379492
KAISER_BEST_ROLLOFF = 0.9475937167399596
NUM_ZEROS = 64
PRECISION = 9


def kaiser_best_filter():
    """resampy.filters.sinc_window: right half of a Kaiser-windowed sinc,
    num_zeros * 2**precision + 1 taps.  Returns (interp_win f64, num_table)."""
    from scipy.signal.windows import kaiser
    num_bits = 2 ** PRECISION
    n = num_bits * NUM_ZEROS
    sinc_win = KAISER_BEST_ROLLOFF * np.sinc(KAISER_BEST_ROLLOFF * np.linspace(0, NUM_ZEROS, num=n + 1, endpoint=True))
    taper = kaiser(2 * n + 1, KAISER_BEST_BETA)[n:]
    return taper * sinc_win, num_bits


def resampled_length(n, sr_in, sr_out):
    """(samples resampy writes, samples librosa returns after fix_length)."""
    ratio = float(sr_out) / sr_in
    return int(n * ratio), int(np.ceil(n * ratio))


def resample_kaiser_best(x, sr_in, sr_out):
    """resampy.resample(x, sr_in, sr_out, filter='kaiser_best') followed by
    librosa's fix_length, for a 1-D float32 signal.

    resampy.interpn.resample_f, vectorised over the output index: the loop over
    taps keeps resampy's order (left wing i = 0.., then right wing k = 0..) and
    rounds the running sum to float32 after every tap, which is what its
    ``y[t] += weight * x[n - i]`` does with a float32 ``y`` and float64 weights."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if sr_in == sr_out:
        return x.copy()
    sample_ratio = float(sr_out) / sr_in
    n_orig = x.shape[0]
    n_out, n_fix = resampled_length(n_orig, sr_in, sr_out)
    if n_out < 1:
        raise ValueError("Input signal length=%d is too small to resample from %d->%d" % (n_orig, sr_in, sr_out))
    interp_win, num_table = kaiser_best_filter()
    if sample_ratio < 1:
        interp_win = interp_win * sample_ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    nwin = interp_win.shape[0]

    scale = min(1.0, sample_ratio)
    time_increment = 1.0 / sample_ratio
    index_step = int(scale * num_table)
    # time_register starts at 0.0 and is advanced by += time_increment: cumsum adds sequentially, like the loop
    time_register = np.concatenate([[0.0], np.cumsum(np.full(n_out - 1, time_increment))]) if n_out > 1 \
        else np.zeros(1)
    n = time_register.astype(np.int64)
    y = np.zeros(n_out, dtype=np.float32)
    xd = x.astype(np.float64)

    def wing(frac, count, sign):
        nonlocal y
        index_frac = frac * num_table
        offset = index_frac.astype(np.int64)
        eta = index_frac - offset
        for i in range(int(count.max()) if count.size else 0):
            live = i < count
            idx = np.where(live, offset + i * index_step, 0)
            weight = interp_win[idx] + eta * interp_delta[idx]
            src = np.where(live, n - i if sign < 0 else n + i + 1, 0)
            term = np.where(live, weight * xd[src], 0.0)
            y = (y.astype(np.float64) + term).astype(np.float32)
        return offset

    frac = scale * (time_register - n)
    off_l = (frac * num_table).astype(np.int64)
    wing(frac, np.minimum(n + 1, (nwin - off_l) // index_step), -1)
    frac = scale - frac
    off_r = (frac * num_table).astype(np.int64)
    wing(frac, np.minimum(n_orig - n - 1, (nwin - off_r) // index_step), +1)

    out = np.zeros(n_fix, dtype=np.float32)
    out[:n_out] = y
    return out


def pcm16_to_float_mono(pcm, channels):
    """librosa.util.buf_to_float + to_mono: interleaved int16 -> float32 mono."""
    y = (1.0 / 32768.0) * np.asarray(pcm, dtype=np.int16).astype(np.float32)
    if channels > 1:
        y = np.mean(y.reshape(-1, channels).T, axis=0, dtype=np.float32)
    return y.astype(np.float32)


def load(pcm16, channels, sr_native, sr=TARGET_SR):
    """librosa.load on already-decoded int16 samples."""
    return resample_kaiser_best(pcm16_to_float_mono(pcm16, channels), sr_native, sr)
