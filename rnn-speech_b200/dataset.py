# coding=utf-8
"""Input pipeline contract of AcousticModel.build_dataset
(/root/reference/models/AcousticModel.py:801-840) without tf.data: an iterable of
mini-batches ``(features [Tmax,B,F] float32 device, lengths [B] int32 device,
dense labels [B, Lmax_in_batch] int32 host, zero-padded)``.

The last, incomplete batch is padded to ``batch_size`` with zero features /
length 0 / empty labels, as create_training_rnn does (:147-159).  Feature
extraction for the whole mini-batch is one GPU launch sequence
(AudioProcessor.process_batch).
"""
import logging

import numpy as np
import torch

from . import labels as labelcodec
from .audioprocessor import AudioProcessor


class AudioBatchDataset(object):
    def __init__(self, input_set, batch_size, max_input_seq_length, max_target_seq_length, signal_processing,
                 char_map, device=None, delta_mode="interp"):
        self.items = [[item[0], item[1]] for item in input_set]
        self.batch_size = batch_size
        self.max_input_seq_length = max_input_seq_length
        self.max_target_seq_length = max_target_seq_length
        self.char_map = char_map
        self.audio_processor = AudioProcessor(max_input_seq_length, signal_processing, delta_mode=delta_mode,
                                              device=device)

    def __len__(self):
        return (len(self.items) + self.batch_size - 1) // self.batch_size

    def _load(self, audio):
        if isinstance(audio, (tuple, list)):
            return np.asarray(audio[0], dtype=np.float32), int(audio[1])
        from .audiofile import load_audio
        return load_audio(audio)

    def _submit(self, prefetcher, start):
        """Load one mini-batch on the host and hand its staging / H2D / feature kernels to the prefetcher."""
        chunk = self.items[start:start + self.batch_size]
        if all(not isinstance(audio, (tuple, list)) for audio, _ in chunk):
            # files: decoded on the worker thread, resampled and featurised on the device without a host round trip
            labs = [labelcodec.get_str_labels(self.char_map, label) if isinstance(label, str) else list(label)
                    for _, label in chunk]
            return prefetcher.submit_files([audio for audio, _ in chunk], time_major=True), labs, len(chunk)
        sigs, srs, labs = [], [], []
        for audio, label in chunk:
            sig, sr = self._load(audio)
            sigs.append(sig)
            srs.append(sr)
            labs.append(labelcodec.get_str_labels(self.char_map, label) if isinstance(label, str)
                        else list(label))
        if len(set(srs)) != 1:
            raise ValueError("all utterances of a mini-batch must share one sample rate, got %s" % sorted(set(srs)))
        # (staged and copied by the worker at once; the feature kernels are enqueued by launch_pending_features --
        #  AcousticModel.run_step calls it in front of its forward pass -- or by the next iteration, whichever comes first)
        return prefetcher.submit(sigs, srs[0], time_major=True, defer_features=True), labs, len(sigs)

    def launch_pending_features(self):
        """Enqueue the feature kernels of the mini-batch the iterator has in flight (no-op when there is none, when
        they are already enqueued, or when the batch comes from files: those are featurised by the worker)."""
        ticket = getattr(self, "_pending_ticket", None)
        launch = getattr(ticket, "launch_features", None)
        if launch is not None:
            launch()

    def __iter__(self):
        """The next mini-batch's features are computed on a side stream while the caller trains on the current one
        (the reference pipeline prefetches too: models/AcousticModel.py:819-822)."""
        from .audioprocessor import BatchPrefetcher
        B = self.batch_size
        prefetcher = BatchPrefetcher(self.audio_processor)
        starts = list(range(0, len(self.items), B))
        pending = self._submit(prefetcher, starts[0]) if starts else None
        for i in range(len(starts)):
            ticket, labs, n_real = pending
            pending = self._submit(prefetcher, starts[i + 1]) if i + 1 < len(starts) else None
            self._pending_ticket = pending[0] if pending is not None else None
            feats, nframes = ticket.result()
            if n_real < B:      # pad the batch (features 0, length 0)
                full = torch.zeros((feats.shape[0], B, feats.shape[2]), dtype=feats.dtype, device=feats.device)
                full[:, :n_real] = feats
                lens = torch.zeros((B,), dtype=torch.int32, device=feats.device)
                lens[:n_real] = nframes
                feats, nframes = full, lens
            over = nframes > self.max_input_seq_length
            if bool(over.any()):
                # the reference would hand tf.nn.ctc_loss a length > Tmax and fail (TODO at
                # models/AcousticModel.py:837-838); clamp instead and say so
                logging.warning("utterance longer than max_input_seq_length: truncated to %d frames",
                                self.max_input_seq_length)
                nframes = torch.clamp(nframes, max=self.max_input_seq_length)
            width = max([len(l) for l in labs] + [1])
            dense = np.zeros((n_real, width), dtype=np.int32)
            for i2, l in enumerate(labs):
                dense[i2, :len(l)] = l
            yield feats, nframes, dense
        prefetcher.close()
