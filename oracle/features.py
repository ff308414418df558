"""Oracle (test infrastructure, NOT product code): CPU restatement of the
reference's feature extraction.

Follows /root/reference/util/audioprocessor.py:63-161 (cited per function) and
the librosa functions it calls (librosa itself is absent from /root/reference
and from this image; requirements.txt:1-5 pins no version).

All arithmetic is float64 like the reference (its ``np.zeros`` pad promotes the
signal to float64, util/audioprocessor.py:96-97).
"""
import numpy as np

FRAME_STRIDE = 0.01   # util/audioprocessor.py:6
FRAME_SIZE = 0.025    # util/audioprocessor.py:7

DELTA_INTERP = 0      # librosa >= 0.6.1: scipy.signal.savgol_filter(mode='interp')
DELTA_EDGE = 1        # librosa <= 0.6.0: np.pad(mode='edge') + lfilter, window / sum|w|


def frame_params(sr):
    """util/audioprocessor.py:88-91 -- python round() is banker's rounding."""
    return int(round(FRAME_SIZE * sr)), int(round(FRAME_STRIDE * sr))


def fbank_num_frames(n, sr):
    """util/audioprocessor.py:92"""
    frame_length, frame_step = frame_params(sr)
    return int(np.ceil(float(np.abs(n - frame_length)) / frame_step))


def mel_filterbank_htk(sr, nfft=512, nfilt=40):
    """util/audioprocessor.py:107-133: 40 un-normalised triangles on an HTK mel
    scale; bin edges floor((nfft+1)*hz/sr)."""
    high_freq_mel = 2595.0 * np.log10(1.0 + (float(sr) / 2.0) / 700.0)
    mel_points = np.linspace(0.0, high_freq_mel, nfilt + 2)
    hz_points = 700.0 * (10.0 ** (mel_points / 2595.0) - 1.0)
    bins = np.floor((nfft + 1) * hz_points / sr)
    fb = np.zeros((nfilt, nfft // 2 + 1))
    for m in range(1, nfilt + 1):
        lo, ce, hi = int(bins[m - 1]), int(bins[m]), int(bins[m + 1])
        for k in range(lo, ce):
            fb[m - 1, k] = (k - bins[m - 1]) / (bins[m] - bins[m - 1])
        for k in range(ce, hi):
            fb[m - 1, k] = (bins[m + 1] - k) / (bins[m + 1] - bins[m])
    return fb, bins


def delta(x, mode=DELTA_INTERP, width=9):
    """librosa.feature.delta(data) along the last axis, order 1, width 9
    (called at util/audioprocessor.py:148-149).

    DELTA_INTERP: librosa>=0.6.1 -> scipy.signal.savgol_filter(data, 9, deriv=1,
        polyorder=1, axis=-1, mode='interp').  Interior: sum_j j*x[t+j]/60;
        the 4 edge frames on each side take the slope of the straight line
        fitted to the first / last 9 frames.
    DELTA_EDGE: librosa<=0.6.0 -> edge-replicate padding, FIR window
        arange(4,-5,-1)/sum|w| (=20), i.e. sum_j j*x[clamp(t+j)]/20.
    """
    x = np.asarray(x, dtype=np.float64)
    half = width // 2
    T = x.shape[-1]
    if mode == DELTA_INTERP:
        if T < width:
            raise ValueError("delta(mode='interp') needs at least %d frames, got %d" % (width, T))
        j = np.arange(-half, half + 1, dtype=np.float64)
        denom = float(np.sum(j * j))
        out = np.empty_like(x)
        for t in range(half, T - half):
            out[..., t] = np.tensordot(x[..., t - half:t + half + 1], j, axes=([-1], [0])) / denom
        out[..., :half] = out[..., half:half + 1]
        out[..., T - half:] = out[..., T - half - 1:T - half]
        return out
    elif mode == DELTA_EDGE:
        j = np.arange(-half, half + 1, dtype=np.float64)
        denom = float(np.sum(np.abs(j)))
        idx = np.clip(np.arange(T)[:, None] + j[None, :].astype(np.int64), 0, T - 1)
        return np.tensordot(x[..., idx], j, axes=([-1], [0])) / denom
    raise ValueError("unknown delta mode %r" % (mode,))


def fbank(sig, sr, max_input_seq_length=None, delta_mode=DELTA_INTERP):
    """util/audioprocessor.py:77-161 (_extract_fbank).  Returns (feat[T',120]
    float64, T) with T the pre-truncation frame count."""
    sig = np.asarray(sig)
    emphasized = np.append(sig[0], sig[1:] - 0.97 * sig[:-1])                # :87
    frame_length, frame_step = frame_params(sr)                              # :88-91
    n = len(emphasized)
    num_frames = int(np.ceil(float(np.abs(n - frame_length)) / frame_step))  # :92
    pad_len = num_frames * frame_step + frame_length                         # :94
    pad_signal = np.append(emphasized, np.zeros(pad_len - n))                # :95-96
    idx = (np.arange(frame_length)[None, :] +
           (np.arange(num_frames) * frame_step)[:, None])                    # :98-100
    frames = pad_signal[idx] * np.hamming(frame_length)                      # :101-103
    nfft, nfilt = 512, 40
    mag = np.absolute(np.fft.rfft(frames, nfft))                             # :105 (crops if frame_length>512)
    pow_frames = (1.0 / nfft) * mag ** 2                                     # :106
    fb, _ = mel_filterbank_htk(sr, nfft, nfilt)                              # :107-133
    fbanks = pow_frames @ fb.T                                               # :134
    fbanks = np.where(fbanks == 0, np.finfo(float).eps, fbanks)              # :135
    fbanks = 10.0 * np.log10(fbanks)                                         # :143
    fbanks = fbanks - (np.mean(fbanks, axis=0) + 1e-8)                       # :146
    fbT = fbanks.T
    d1 = delta(fbT, delta_mode)                                              # :148
    d2 = delta(d1, delta_mode)                                               # :149
    feat = np.vstack([fbT, d1, d2]).T                                        # :150-152
    T = len(feat)
    if max_input_seq_length is not None and T > max_input_seq_length:        # :157-159
        feat = feat[:max_input_seq_length]
    return feat, T


# ----------------------------------------------------------------------------
# MFCC: librosa.feature.mfcc(sig, sr, hop_length=round(.01 sr), n_fft=round(.025 sr))
# (util/audioprocessor.py:65-66).  librosa is absent -> restated from its
# published algorithm (librosa 0.6-0.9 defaults).  PARITY UNPINNED.
# ----------------------------------------------------------------------------

def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank_slaney(sr, n_fft, n_mels=128):
    """librosa.filters.mel(sr, n_fft, n_mels=128, fmin=0, fmax=sr/2, htk=False, norm=1)."""
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return w * enorm[:, None]


def dct_ortho(n_out, n_in):
    """Orthonormal DCT-II basis [n_out, n_in] (scipy.fftpack.dct(type=2, norm='ortho')[:n_out])."""
    k = np.arange(n_out)[:, None]
    n = np.arange(n_in)[None, :]
    basis = np.cos(np.pi * k * (2 * n + 1) / (2.0 * n_in)) * np.sqrt(2.0 / n_in)
    basis[0] *= 1.0 / np.sqrt(2.0)
    return basis


def mfcc_num_frames(n, sr):
    _, hop = frame_params(sr)
    return 1 + n // hop


def mfcc(sig, sr, max_input_seq_length=None, n_mfcc=20, n_mels=128, top_db=80.0):
    """util/audioprocessor.py:63-75 (_extract_mfcc)."""
    sig = np.asarray(sig, dtype=np.float64)
    n_fft, hop = frame_params(sr)
    pad = n_fft // 2
    y = np.pad(sig, pad, mode="reflect")                       # stft(center=True, pad_mode='reflect')
    T = 1 + (len(y) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + (np.arange(T) * hop)[:, None]
    # scipy.signal.get_window('hann', n_fft, fftbins=True): periodic Hann
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)
    spec = np.fft.rfft(y[idx] * win, n_fft)                    # [T, 1+n_fft/2]
    power = np.abs(spec) ** 2                                  # melspectrogram(power=2)
    mel = power @ mel_filterbank_slaney(sr, n_fft, n_mels).T   # [T, n_mels]
    log_spec = 10.0 * np.log10(np.maximum(1e-10, mel))         # power_to_db(ref=1, amin=1e-10)
    log_spec = np.maximum(log_spec, log_spec.max() - top_db)   # top_db=80 over the whole utterance
    out = log_spec @ dct_ortho(n_mfcc, n_mels).T               # [T, n_mfcc]
    if max_input_seq_length is not None and T > max_input_seq_length:
        out = out[:max_input_seq_length]
    return out, T
