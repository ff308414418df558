# coding=utf-8
"""B200 replacement for the reference's ``util/audioprocessor.py``.

Same class name, constructor, attributes and return values as
``util.audioprocessor.AudioProcessor`` (/root/reference/util/audioprocessor.py:10-61)
so callers (stt.py:24-27, :240, :335-354; models/AcousticModel.py:726, :812-813)
need no change.  The arithmetic runs in the CUDA feature kernels
(csrc/features.cu) through the C ABI; there is no CPU fallback.

Added (not in the reference): ``process_batch`` -- a whole batch of utterances in
one launch, result left on the device, optionally time-major, which is what the
training path consumes.
"""
import os

import numpy as np
import torch

from . import _lib

# GLOBALS (util/audioprocessor.py:6-7)
FRAME_STRIDE = 0.01
FRAME_SIZE = 0.025

_DELTA_MODES = {"interp": _lib.DELTA_INTERP, "edge": _lib.DELTA_EDGE}


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


_STAGE_POOL = None


def _stage_pool():
    """Threads that assemble a mini-batch in the pinned staging buffer (created on first use)."""
    global _STAGE_POOL
    if _STAGE_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _STAGE_POOL = ThreadPoolExecutor(max_workers=4, thread_name_prefix="rs-stage")
    return _STAGE_POOL


class AudioProcessor(object):
    def __init__(self, max_input_seq_length, feature_type="mfcc", delta_mode="interp", device=None, n_mfcc=20):
        """
        feature_type - string options are: mfcc, fbank
        mfcc is a 20-dim input
        fbank is 120-dim input (mel filterbank with delta and double delta)

        n_mfcc - number of cepstral coefficients kept by the mfcc extractor.  The reference calls
            librosa.feature.mfcc without n_mfcc (util/audioprocessor.py:65-66), i.e. 20 (feature_size = 20, :21);
            its README speaks of 40, and BASELINE config 1 is quoted at 40: pass n_mfcc=40 for that.

        delta_mode - which librosa.feature.delta the fbank path reproduces:
            "interp" (librosa >= 0.6.1, scipy savgol_filter mode='interp') or
            "edge"   (librosa <= 0.6.0, edge-replicated FIR, window / sum|w|).
        """
        self.max_input_seq_length = max_input_seq_length
        self.feature_type = feature_type
        if self.feature_type == "mfcc":
            if not 1 <= int(n_mfcc) <= 128:
                raise ValueError("n_mfcc must be in [1, 128] (128 mel bands), got %r" % (n_mfcc,))
            self.feature_size = int(n_mfcc)
        elif self.feature_type == "fbank":
            self.feature_size = 120
        else:
            raise ValueError("{0} is not a valid extraction function, \
            only fbank and mfcc are accepted.".format(self.feature_type))
        if delta_mode not in _DELTA_MODES:
            raise ValueError("delta_mode must be 'interp' or 'edge', got %r" % (delta_mode,))
        self.delta_mode = delta_mode
        self._device = device
        self._ws = None

    @staticmethod
    def get_mfcc_length_from_duration(duration):
        """util/audioprocessor.py:29-39"""
        length = int(duration // FRAME_STRIDE) - 1
        return length

    # ------------------------------------------------------------------ helpers
    def _dev(self):
        if not torch.cuda.is_available():
            raise RuntimeError("rnnspeech_b200.AudioProcessor needs a CUDA device (no CPU fallback)")
        return torch.device(self._device if self._device is not None else "cuda")

    def num_frames(self, n_samples, sr):
        if self.feature_type == "fbank":
            return int(_lib.raw("rs_fbank_num_frames")(int(n_samples), int(sr)))
        return int(_lib.raw("rs_mfcc_num_frames")(int(n_samples), int(sr)))

    def _workspace(self, nbytes, dev):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        return self._ws

    def features_device(self, pcm_d, offsets_d, batch, max_samples, sr, time_major=True, out=None, nframes=None):
        """Device-resident entry: pcm_d float32 [sum n], offsets_d int64 [B+1].
        Returns (features [Tmax,B,F] or [B,Tmax,F] float32, nframes int32 [B]),
        both on the device; nframes is the PRE-truncation frame count."""
        dev = pcm_d.device
        Tmax, F = int(self.max_input_seq_length), self.feature_size
        shape = (Tmax, batch, F) if time_major else (batch, Tmax, F)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=dev)
        if nframes is None:
            nframes = torch.empty((batch,), dtype=torch.int32, device=dev)
        assert out.is_contiguous() and tuple(out.shape) == shape and out.dtype == torch.float32
        if self.feature_type == "fbank":
            ws_bytes = _lib.raw("rs_fbank_workspace_bytes")(batch, int(max_samples), int(sr))
            ws = self._workspace(ws_bytes, dev)
            _lib.call("rs_fbank_forward", pcm_d.data_ptr(), offsets_d.data_ptr(), batch, int(max_samples), int(sr),
                      Tmax, _DELTA_MODES[self.delta_mode], 1 if time_major else 0, out.data_ptr(),
                      nframes.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr())
        else:
            ws_bytes = _lib.raw("rs_mfcc_workspace_bytes")(batch, int(max_samples), int(sr))
            ws = self._workspace(ws_bytes, dev)
            _lib.call("rs_mfcc_forward", pcm_d.data_ptr(), offsets_d.data_ptr(), batch, int(max_samples), int(sr),
                      Tmax, F, 1 if time_major else 0, out.data_ptr(), nframes.data_ptr(), ws.data_ptr(),
                      ws.numel(), _stream_ptr())
        return out, nframes

    def process_batch(self, signals, sr, time_major=True):
        """signals: list of 1-D float arrays (host).  One H2D copy, one launch
        sequence; returns device tensors (features, nframes)."""
        pcm_d, off_d, lens = self.stage_batch(signals, sr)
        return self.features_device(pcm_d, off_d, len(signals), max(lens), sr, time_major=time_major)

    def stage_batch(self, signals, sr):
        """Host PCM of a mini-batch -> device: returns (pcm_d float32 [sum n], offsets_d int64 [B+1], lengths)."""
        dev = self._dev()
        lens = [int(len(s)) for s in signals]
        if min(lens) < 1:
            raise ValueError("empty signal")
        if self.feature_type == "fbank" and self.delta_mode == "interp":
            for n in lens:
                if self.num_frames(n, sr) < 9:
                    # librosa.feature.delta(mode='interp') raises when width 9 > number of frames
                    raise ValueError("delta(mode='interp') needs at least 9 frames, got %d" % self.num_frames(n, sr))
        offsets = np.zeros(len(signals) + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        total = int(offsets[-1])
        # Pinned staging, allocated once and reused (page-locking 20 MB per call costs more than the copy); two
        # buffers alternate so that the copy of the previous call may still be in flight.  The whole mini-batch --
        # PCM of every utterance, then the int64 offsets -- is assembled in ONE pinned buffer and goes to the
        # device with ONE cudaMemcpyAsync issued by the library (rs_memcpy_h2d_async).
        slot = self._stage_next = (getattr(self, "_stage_next", 0) + 1) % 2
        stage = getattr(self, "_stage", None)
        if stage is None:
            stage = self._stage = [None, None]
            self._stage_ev = [None, None]
        opos = (total + 1) // 2 * 2                          # 8-byte aligned home of the int64 offsets
        nfloats = opos + len(offsets) * 2
        if stage[slot] is None or stage[slot].numel() < nfloats:
            stage[slot] = torch.empty((max(nfloats, 1 << 16),), dtype=torch.float32, pin_memory=True)
            self._stage_ev[slot] = torch.cuda.Event()
        else:
            self._stage_ev[slot].synchronize()              # the copy that last used this buffer has finished
        host = stage[slot]
        hview = host.numpy()
        def put(lo, hi):
            for k in range(lo, hi):
                o = offsets[k]
                hview[o:o + lens[k]] = signals[k]           # (casts to float32 when the source is not)
        if total >= (1 << 21) and len(signals) >= 8:
            # a large batch (cfg-5: 82 MB) is copied by a few threads: numpy releases the GIL inside the copy, and one
            # core moves ~10 GB/s into pinned memory
            nt = 4
            bounds = [len(signals) * i // nt for i in range(nt + 1)]
            for f in [_stage_pool().submit(put, bounds[i], bounds[i + 1]) for i in range(nt)]:
                f.result()
        else:
            put(0, len(signals))
        hview[opos:nfloats].view(np.int64)[:] = offsets
        batch_d = torch.empty((nfloats,), dtype=torch.float32, device=dev)
        _lib.call("rs_memcpy_h2d_async", batch_d.data_ptr(), host.data_ptr(), 4 * nfloats, _stream_ptr())
        self._stage_ev[slot].record()
        pcm_d = batch_d[:total]
        off_d = batch_d[opos:nfloats].view(torch.int64)
        return pcm_d, off_d, lens

    # ---------------------------------------------------------- reference API
    def process_signal(self, sig, sr):
        """
        :param sig: audio signal to process
        :param sr: audio signal rate
        :returns: mfcc: feature array [T', F] (truncated to max_input_seq_length)
        :returns: mfcc_length: original length of the features before truncation
        (util/audioprocessor.py:52-61)
        """
        feat, nframes = self.process_batch([np.asarray(sig)], sr, time_major=False)
        length = int(nframes.cpu()[0])
        keep = min(length, int(self.max_input_seq_length))
        return feat[0, :keep].cpu().numpy(), length

    def process_audio_file(self, file_name):
        """util/audioprocessor.py:41-50: ``sig, sr = librosa.load(file_name, mono=True)`` then process_signal.
        The container (RIFF/WAVE, FLAC) is parsed on the host; the int16 -> float32 conversion, the mono mix, the
        resampling to 22 050 Hz (resampy 'kaiser_best' restated, csrc/resample.cu) and the features all run on
        the device, with the decoded PCM as the only host-to-device copy."""
        feats, lengths = self.process_audio_files([file_name], time_major=False)
        length = int(lengths.cpu()[0])
        keep = min(length, int(self.max_input_seq_length))
        return feats[0, :keep].cpu().numpy(), length

    def process_audio_files(self, file_names, time_major=True, sr=None):
        """A batch of files in one go: returns device tensors (features [Tmax,B,F] or [B,Tmax,F], nframes [B]).
        sr: target rate (default 22 050 like librosa.load; pass the corpus rate, e.g. 16 000, to skip resampling)."""
        from . import audiofile
        dev = self._dev()
        decoded = audiofile.decode_files(file_names)
        target = audiofile.TARGET_SR if sr is None else int(sr)
        pcm_d, off_d, lens, out_sr = audiofile.load_batch_device(decoded, dev, sr=target)
        if self.feature_type == "fbank" and self.delta_mode == "interp":
            for n in lens:
                if self.num_frames(n, out_sr) < 9:
                    raise ValueError("delta(mode='interp') needs at least 9 frames, got %d" % self.num_frames(n, out_sr))
        return self.features_device(pcm_d, off_d, len(lens), max(lens), out_sr, time_major=time_major)


class BatchPrefetcher(object):
    """Input prefetch, the counterpart of the reference pipeline's ``.map(..., num_parallel_calls=2).prefetch(30)``
    (models/AcousticModel.py:819-822): staging, the host-to-device copy and the feature kernels of the NEXT
    mini-batch run on a worker thread and a low-priority side stream while the model trains on the current one.

        pre = BatchPrefetcher(audio_processor)
        ticket = pre.submit(signals, sr)            # returns at once
        ...                                         # train on the previous batch
        feats, nframes = ticket.result()            # current stream waits for the side stream's event
    """

    class _Ticket(object):
        def __init__(self, future, owner=None, deferred=None):
            self._future = future
            self._owner = owner
            self._deferred = deferred            # (sr, time_major): the feature kernels have not been enqueued yet
            self._out = None

        def launch_features(self):
            """Deferred ticket: enqueue the feature kernels NOW, on the prefetcher's low-priority stream, behind the
            host-to-device copy and behind whatever the calling stream holds at this moment.  The caller runs ahead of
            the device, so kernels enqueued by the worker thread as soon as its copy is issued land wherever the device
            happens to be -- measured at cfg-2 (profiles/r02d_e2e_phases.txt): in the middle of the previous step's
            backward pass, whose recurrent launches they hold up (8.55 -> 9.18 ms).  Called right in front of a step's
            forward pass they run beside its head (input dense, first chunk GEMM: most SMs idle) and cost nothing."""
            if self._deferred is None or self._out is not None:
                return
            pcm_d, off_d, lens, h2d_ev = self._future.result()
            own = self._owner
            sr, time_major = self._deferred
            cur = torch.cuda.current_stream(own.device)
            gate = torch.cuda.Event()
            gate.record(cur)
            own.stream.wait_event(h2d_ev)
            own.stream.wait_event(gate)
            with torch.cuda.stream(own.stream):
                feats, nframes = own.audio_processor.features_device(pcm_d, off_d, len(lens), max(lens), sr,
                                                                     time_major=time_major)
                ev = torch.cuda.Event()
                ev.record(own.stream)
            self._out = (feats, nframes, ev)

        def staged(self):
            """Deferred ticket: the staged PCM itself -- (pcm_d float32 [sum n], offsets_d int64 [B + 1], lengths) --
            with the calling stream ordered behind the host-to-device copy (AcousticModel.infer_ticket)."""
            pcm_d, off_d, lens, h2d_ev = self._future.result()
            cur = torch.cuda.current_stream(pcm_d.device)
            cur.wait_event(h2d_ev)
            pcm_d.record_stream(cur)
            off_d.record_stream(cur)
            return pcm_d, off_d, lens

        def result(self):
            if self._deferred is not None:
                self.launch_features()
                feats, nframes, ev = self._out
            else:
                feats, nframes, ev = self._future.result()
            cur = torch.cuda.current_stream(feats.device)
            cur.wait_event(ev)
            feats.record_stream(cur)
            nframes.record_stream(cur)
            return feats, nframes

    def __init__(self, audio_processor):
        from concurrent.futures import ThreadPoolExecutor
        self.audio_processor = audio_processor
        self.device = audio_processor._dev()
        self.stream = torch.cuda.Stream(device=self.device, priority=0)      # 0 = least urgent
        self._pool = ThreadPoolExecutor(max_workers=1)

    def _work(self, signals, sr, time_major):
        torch.cuda.set_device(self.device)
        with torch.cuda.stream(self.stream):
            feats, nframes = self.audio_processor.process_batch(signals, sr, time_major=time_major)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return feats, nframes, ev

    def _work_stage(self, signals, sr, after=None):
        torch.cuda.set_device(self.device)
        if after is not None:
            self.stream.wait_event(after)          # the copy itself waits for this point of the caller's stream
        with torch.cuda.stream(self.stream):
            pcm_d, off_d, lens = self.audio_processor.stage_batch(signals, sr)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return pcm_d, off_d, lens, ev

    def _work_files(self, file_names, time_major):
        torch.cuda.set_device(self.device)
        with torch.cuda.stream(self.stream):
            feats, nframes = self.audio_processor.process_audio_files(file_names, time_major=time_major)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return feats, nframes, ev

    def submit(self, signals, sr, time_major=True, defer_features=False, copy_after_current=False):
        """defer_features: the worker only stages and copies; the feature kernels are enqueued by
        ticket.launch_features() (or by ticket.result(), whichever comes first).
        copy_after_current (with defer_features): the host-to-device copy waits for everything the calling stream holds
        at this moment (a caller that runs ahead of the device submits while the previous step's backward pass is still
        running).  Measured at cfg-2 (profiles/r02d_sweep11.log): the 20 MB copy costs the pass it lands in ~0.4 ms
        either way -- backward 8.5 -> 8.9 ms without the gate, forward 5.0 -> 5.4 ms with it -- so it is off by default."""
        if defer_features:
            after = None
            if copy_after_current:
                after = torch.cuda.Event()
                after.record(torch.cuda.current_stream(self.device))
            return BatchPrefetcher._Ticket(self._pool.submit(self._work_stage, signals, sr, after), self, (sr, time_major))
        return BatchPrefetcher._Ticket(self._pool.submit(self._work, signals, sr, time_major))

    def submit_files(self, file_names, time_major=True):
        """Decode (host), then convert / resample / featurise (device, side stream) a mini-batch of audio files."""
        return BatchPrefetcher._Ticket(self._pool.submit(self._work_files, file_names, time_major))

    def close(self):
        self._pool.shutdown(wait=True)
