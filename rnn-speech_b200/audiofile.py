"""Audio front door.

The reference calls ``librosa.load(file_name, mono=True)`` (util/audioprocessor.py:49):
decode the container, scale to float32, down-mix to mono, resample to 22 050 Hz with
resampy's 'kaiser_best' filter.  Here the container is parsed on the host (RIFF/WAV in
this module, FLAC by the library's ``rs_flac_decode_host``); the decoded int16 samples
go to the device as they are and everything after that -- int16 -> float32, mono mix,
band-limited sinc resampling -- runs in csrc/resample.cu through the C ABI, the result
staying on the device for the feature kernels.  No CPU fallback for those steps.

librosa / resampy / libFLAC are third-party and absent: the behaviour is restated from
their published algorithms (oracle/resample.py is the checker) -- parity unpinned upstream.
"""
import ctypes
import hashlib
import struct
import threading

import numpy as np

from . import _lib

TARGET_SR = 22050      # librosa.load's default sr


class DecodedAudio(object):
    """samples: interleaved [frames * channels], int16 (fmt 's16', scaled by 2^-15 on the device) or float32 already
    scaled to [-1, 1) (fmt 'f32': sample formats other than 16-bit).  The mono mix always happens on the device."""
    __slots__ = ("samples", "fmt", "channels", "sr", "frames")

    def __init__(self, samples, fmt, channels, sr, frames):
        self.samples, self.fmt, self.channels, self.sr, self.frames = samples, fmt, channels, sr, frames


def _to_float(x, scale):
    """Sample-format conversion of the decode step (what audioread / soundfile hand librosa): integer -> float32."""
    return np.ascontiguousarray(x.astype(np.float32) * np.float32(scale), dtype=np.float32)


def decode_wav(data):
    """RIFF/WAVE PCM (format 1, or WAVE_FORMAT_EXTENSIBLE carrying PCM) and IEEE float (format 3)."""
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, fmt, body = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        chunk = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = chunk
        elif cid == b"data":
            body = chunk
            break
        pos += 8 + size + (size & 1)
    if fmt is None or body is None or len(fmt) < 16:
        raise ValueError("WAVE file without fmt / data chunks")
    tag, nch, sr, _, _, bits = struct.unpack("<HHIIHH", fmt[:16])
    if tag == 0xFFFE and len(fmt) >= 26:
        tag = struct.unpack("<H", fmt[24:26])[0]
    if nch < 1 or sr < 1:
        raise ValueError("bad WAVE header (channels %d, rate %d)" % (nch, sr))
    width = bits // 8
    frames = len(body) // (width * nch) if width else 0
    body = body[:frames * width * nch]
    if tag == 1 and bits == 16:
        return DecodedAudio(np.frombuffer(body, dtype="<i2"), "s16", nch, sr, frames)
    if tag == 1 and bits == 8:
        x = np.frombuffer(body, dtype=np.uint8).astype(np.int16) - 128
        return DecodedAudio(_to_float(x, 1.0 / 128.0), "f32", nch, sr, frames)
    if tag == 1 and bits == 24:
        b = np.frombuffer(body, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        x = (b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16))
        x = np.where(x >= 1 << 23, x - (1 << 24), x)
        return DecodedAudio(_to_float(x, 1.0 / (1 << 23)), "f32", nch, sr, frames)
    if tag == 1 and bits == 32:
        return DecodedAudio(_to_float(np.frombuffer(body, dtype="<i4"), 1.0 / (1 << 31)), "f32", nch, sr, frames)
    if tag == 3 and bits == 32:
        return DecodedAudio(np.ascontiguousarray(np.frombuffer(body, dtype="<f4")), "f32", nch, sr, frames)
    raise ValueError("unsupported WAVE encoding (format tag %d, %d bits)" % (tag, bits))


def decode_flac(data, verify_md5=True):
    """FLAC through the library's host decoder (csrc/flac.cu); every frame CRC is checked there, and the
    STREAMINFO MD5 of the unencoded audio is checked here for 16-bit streams."""
    buf = np.frombuffer(data, dtype=np.uint8)
    sr, nch, bps, frames = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
    md5 = (ctypes.c_uint8 * 16)()
    args = (ctypes.byref(sr), ctypes.byref(nch), ctypes.byref(bps), ctypes.byref(frames), md5)
    _lib.call("rs_flac_decode_host", buf.ctypes.data, buf.size, None, 0, *args)
    if frames.value * nch.value > 64 * buf.size:
        # more than 64 samples per stored byte: silence, or a damaged header -- count the frames before allocating
        _lib.call("rs_flac_decode_host", buf.ctypes.data, buf.size, None, -1, *args)
    # int32 scratch kept per thread: a fresh large array per file costs page faults and an munmap, which serialise the
    # decoder threads on the address-space lock (measured: 1.1 -> 4.5 ms per file with four threads)
    need = frames.value * nch.value
    scratch = getattr(_TLS, "scratch", None)
    if scratch is None or scratch.size < need:
        scratch = _TLS.scratch = np.empty(max(need, 1 << 18), dtype=np.int32)
    _lib.call("rs_flac_decode_host", buf.ctypes.data, buf.size, scratch.ctypes.data, scratch.size, *args)
    out = scratch[:frames.value * nch.value]
    if bps.value == 16:
        pcm = out.astype("<i2")
        want = bytes(md5)
        if verify_md5 and any(want) and hashlib.md5(pcm.tobytes()).digest() != want:
            raise ValueError("FLAC stream decodes, but its audio MD5 does not match STREAMINFO")
        return DecodedAudio(pcm, "s16", nch.value, sr.value, frames.value)
    return DecodedAudio(_to_float(out, 1.0 / (1 << (bps.value - 1))), "f32", nch.value, sr.value, frames.value)


def decode_file(path):
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:4] == b"RIFF":
        return decode_wav(data)
    if data[:4] == b"fLaC" or data[:3] == b"ID3":
        return decode_flac(data)
    raise NotImplementedError("%r: only RIFF/WAVE and FLAC containers are decoded here (the reference's librosa.load "
                              "also opens whatever audioread/ffmpeg can)" % (path,))


_POOL = None
_TLS = threading.local()


def decode_files(paths, workers=None):
    """decode_file over a list, on a small thread pool: reading, the FLAC decoder (a ctypes call), numpy conversions
    and hashlib all release the GIL, so the files of a mini-batch decode in parallel (one core decodes ~200 ten-second
    FLAC utterances per second; a B200 trains on ~1600)."""
    global _POOL
    paths = list(paths)
    if len(paths) < 2 or workers == 1:
        return [decode_file(p) for p in paths]
    if _POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=workers or max(2, min(16, (os.cpu_count() or 2))),
                                   thread_name_prefix="rs-decode")
    return list(_POOL.map(decode_file, paths))


def duration_seconds(path):
    """Length of an audio file in seconds from its header alone (what the reference asks mutagen for,
    util/dataprocessor.py:232-241, to order the training set by duration); 0 if the file is not recognised."""
    try:
        with open(path, "rb") as fh:
            head = fh.read(1 << 16)
        if head[:4] == b"RIFF" and head[8:12] == b"WAVE":
            pos, block_align, sr = 12, None, None
            while pos + 8 <= len(head):
                cid, size = head[pos:pos + 4], struct.unpack("<I", head[pos + 4:pos + 8])[0]
                if cid == b"fmt " and pos + 24 <= len(head):
                    _, nch, sr, _, block_align, _ = struct.unpack("<HHIIHH", head[pos + 8:pos + 24])
                elif cid == b"data":
                    if not block_align or not sr:
                        return 0
                    import os
                    size = min(size, os.path.getsize(path) - (pos + 8))
                    return (size // block_align) / float(sr)
                pos += 8 + size + (size & 1)
            return 0
        if head[:4] == b"fLaC" or head[:3] == b"ID3":
            sr, nch, bps, frames = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
            buf = np.frombuffer(head, dtype=np.uint8)
            if head[:3] == b"ID3":                      # the tag may be longer than what was read
                with open(path, "rb") as fh:
                    buf = np.frombuffer(fh.read(), dtype=np.uint8)
            try:
                _lib.call("rs_flac_decode_host", buf.ctypes.data, buf.size, None, 0, ctypes.byref(sr), ctypes.byref(nch),
                          ctypes.byref(bps), ctypes.byref(frames), None)
            except ValueError:
                # total_samples unknown in STREAMINFO (the header-only call then walks frames and runs out of data)
                d = decode_file(path)
                return d.frames / float(d.sr)
            return frames.value / float(sr.value)
    except (OSError, ValueError, RuntimeError, struct.error):
        pass
    return 0


def load_batch_device(decoded, device, sr=TARGET_SR):
    """librosa.load's post-decode half for a list of DecodedAudio, on the device.

    Returns (pcm_d float32 [sum n], offsets_d int64 [B+1], lengths list, sr).  One H2D copy per group of files
    that share (format, channels, native rate), one resample launch sequence per group."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("rnnspeech_b200.audiofile needs a CUDA device (no CPU fallback)")
    dev = torch.device(device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    out_len = [int(_lib.raw("rs_resample_num_samples")(d.frames, d.sr, sr)) if sr is not None else d.frames
               for d in decoded]
    for d, n in zip(decoded, out_len):
        if d.frames < 1 or n < 1:
            raise ValueError("empty audio")
    out_off = np.zeros(len(decoded) + 1, dtype=np.int64)
    np.cumsum(out_len, out=out_off[1:])
    pcm_d = torch.empty((int(out_off[-1]),), dtype=torch.float32, device=dev)
    target = {}
    for i, d in enumerate(decoded):
        target.setdefault((d.fmt, d.channels, d.sr), []).append(i)
    keep = []
    for (fmt, nch, sr_native), idxs in target.items():
        same_rate = sr is None or sr_native == sr
        runs = _contiguous_runs(idxs)
        for run in runs:
            host = np.concatenate([decoded[i].samples for i in run])
            src_d = torch.from_numpy(host).pin_memory().to(dev, non_blocking=True)
            keep.append(src_d)
            first, last = run[0], run[-1]
            if same_rate:
                dst = pcm_d[int(out_off[first]):int(out_off[last + 1])]
                _lib.call("rs_pcm16_to_f32" if fmt == "s16" else "rs_pcm_f32_to_mono", src_d.data_ptr(),
                          int(sum(decoded[i].frames for i in run)), nch, dst.data_ptr(), stream)
                continue
            in_off = np.zeros(len(run) + 1, dtype=np.int64)
            np.cumsum([decoded[i].frames for i in run], out=in_off[1:])
            o_off = out_off[first:last + 2] - out_off[first]
            offs_d = torch.from_numpy(np.concatenate([in_off, o_off])).to(dev)
            max_out = int(max(out_len[i] for i in run))
            ws = torch.empty((int(_lib.raw("rs_resample_workspace_bytes")(len(run), max_out)),), dtype=torch.uint8,
                             device=dev)
            keep += [offs_d, ws]
            dst = pcm_d[int(out_off[first]):]
            _lib.call("rs_resample_forward", src_d.data_ptr(), _lib.PCM_S16 if fmt == "s16" else _lib.PCM_F32, nch,
                      offs_d.data_ptr(), len(run), max_out, int(sr_native), int(sr), dst.data_ptr(),
                      offs_d[len(run) + 1:].data_ptr(), ws.data_ptr(), ws.numel(), stream)
    offsets_d = torch.from_numpy(out_off).to(dev)
    for t in keep:                       # allocator: these buffers are in use by work queued on the current stream
        t.record_stream(torch.cuda.current_stream(dev))
    return pcm_d, offsets_d, out_len, (sr if sr is not None else decoded[0].sr)


def _contiguous_runs(idxs):
    runs, cur = [], [idxs[0]]
    for i in idxs[1:]:
        if i == cur[-1] + 1:
            cur.append(i)
        else:
            runs.append(cur)
            cur = [i]
    runs.append(cur)
    return runs


def load_audio(path, sr=TARGET_SR, device="cuda"):
    """librosa.load(path, sr=sr, mono=True) -> (float32 ndarray, sr); the arithmetic runs on the device."""
    pcm_d, _, _, out_sr = load_batch_device([decode_file(path)], device, sr=sr)
    return pcm_d.cpu().numpy(), out_sr
