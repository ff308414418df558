# coding=utf-8
"""B200 replacement for the reference's ``models/AcousticModel.py``.

Acoustic RNN trained with CTC loss: input dense -> L x LSTM -> output dense ->
CTC.  Same class name, constructor and method names / return values as
``models.AcousticModel.AcousticModel`` (/root/reference/models/AcousticModel.py:28-939)
so that ``stt.py`` keeps its structure; every method that took a ``tf.Session``
still accepts (and ignores) a ``sess`` argument.  All arithmetic runs in the CUDA
kernels behind the C ABI (include/rnnspeech_b200.h); torch tensors are used only
as device-memory holders and for NCCL plumbing.  There is no CPU fallback.

Data-parallel training (not in the reference): one process
per GPU, each rank runs its own mini-batches, one NCCL all-reduce(SUM) over the
flat gradient buffer per step (the reference differentiates the SUM of the
per-item losses, models/AcousticModel.py:386-388, so SUM -- not mean -- keeps the
update identical to the reference run with mini_batch_size = world size).
"""
import logging
import os
import pickle
import time
from random import randint

import numpy as np
import torch

from . import _lib
from . import labels as labelcodec
from .audioprocessor import AudioProcessor
from .dist import allreduce_sum_, world as dist_world


def _stream_ptr():
    return torch.cuda.current_stream().cuda_stream


class OutOfRangeError(Exception):
    """End of dataset (stands in for tf.errors.OutOfRangeError)."""


def levenshtein(a, b):
    """Edit distance between two integer sequences (numpy row DP)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if len(a) == 0:
        return len(b)
    if len(b) == 0:
        return len(a)
    idx = np.arange(len(b) + 1)
    prev = idx.copy()
    for i in range(1, len(a) + 1):
        cost = (b != a[i - 1]).astype(np.int64)
        cur = np.empty_like(prev)
        cur[0] = i
        cur[1:] = np.minimum(prev[1:] + 1, prev[:-1] + cost)
        cur = np.minimum.accumulate(cur - idx) + idx
        prev = cur
    return int(prev[-1])


# Utterances per launch of the tensor-core recurrent kernels: a larger mini-batch runs as batch tiles of this size.
# csrc/lstm_rec_ts.cu takes up to 64 rows, but its fastest kernels (two 16-row chains per CTA, validated exchange) take
# 32: measured at BASELINE config 5 (256 clips x 5 s), tiles of 32 give 11 588 clips/s against 10 546 with tiles of 64
# (profiles/r02c_sweep5.log).  RS_TC_MAX_BATCH overrides.
TC_MAX_BATCH = int(os.environ.get("RS_TC_MAX_BATCH", "32"))


class AcousticModel(object):
    def __init__(self, num_layers, hidden_size, batch_size, max_input_seq_length,
                 max_target_seq_length, input_dim, normalization, num_labels, device=None, seed=0):
        """
        Initialize the acoustic rnn model parameters (models/AcousticModel.py:29-94)

        :param num_layers: number of lstm layers
        :param hidden_size: size of hidden layers
        :param batch_size: number of training examples fed at once
        :param max_input_seq_length: maximum length of input vector sequence
        :param max_target_seq_length: maximum length of ouput vector sequence
        :param input_dim: dimension of input vector
        :param normalization: boolean indicating whether or not to normalize data in a input batch
        :param num_labels: the numbers of output labels
        """
        self.num_layers = num_layers
        self.hidden_size = hidden_size
        self.batch_size = batch_size
        self.max_input_seq_length = max_input_seq_length
        self.max_target_seq_length = max_target_seq_length
        self.input_dim = input_dim
        self.normalization = normalization
        self.num_labels = num_labels
        if not torch.cuda.is_available():
            raise RuntimeError("rnnspeech_b200.AcousticModel needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.seed = seed
        self.beta_skip = _lib.CTC_BETA_SOURCE
        self.decoder = "beam"            # the reference's prediction op; "greedy" = per-frame argmax path

        self.input_keep_prob = self.output_keep_prob = 1.0
        self.grad_clip = None
        self.lr_decay_factor = None
        self.learning_rate_var = None
        self.global_step = 0
        self.is_training = False
        self.rnn_created = False
        self.training_created = False
        self.tensorboard_dir = None
        self.timeline_enabled = False
        self._train_iter = self._valid_iter = None
        self._train_dataset = self._valid_dataset = None

        self._handle = None
        self.params = self.grads = self.adam_m = self.adam_v = None
        self.rnn_state = None
        self._dropout_calls = 0
        # TF re-initialises Adam's beta powers and slots on restore (they are not in the reference's save_list,
        # models/AcousticModel.py:515-527), so the bias-correction step restarts with the slots
        self._adam_step = 0
        self._mini_batches = 0           # host copy of the mini-batch counter (models/AcousticModel.py:378-380)
        self._err_stream = None          # side stream of the prediction + edit-distance ops
        self._err_event = None
        self._acc_event, self._acc_recorded = None, False         # end_batch's early read of the accumulators
        self._read_stream = self._acc_host = self._read_done = None
        self._dataset_empty = False

    # ------------------------------------------------------------ construction
    def _create_common(self):
        if self.rnn_created:
            logging.fatal("Trying to create the acoustic RNN but it is already.")
            return
        def create(batch):
            hh = _lib.c_void_p()
            _lib.call("rs_am_create", _lib.ctypes.byref(hh), self.num_layers, self.hidden_size, self.input_dim,
                      self.num_labels, batch, self.max_input_seq_length)
            if self.normalization:       # models/AcousticModel.py:253-259 (config.ini batch_normalization)
                _lib.call("rs_am_set_normalization", hh, 1)
            return hh
        # The tensor-core recurrent kernels take up to TC_MAX_BATCH utterances; a larger mini-batch (BASELINE
        # config 5: 256 clips) runs as batch tiles of that size, one handle per tile size, sharing parameters,
        # gradients and workspace (utterances are independent: SURVEY 8e).
        self._tiles = None
        h = None
        # (batch normalisation takes its statistics over the whole mini-batch: no tiling then)
        if self.batch_size > TC_MAX_BATCH and not self.normalization:
            probe = create(TC_MAX_BATCH)
            if bool(_lib.raw("rs_am_uses_tensor_cores")(probe)):
                bounds = list(range(0, self.batch_size, TC_MAX_BATCH)) + [self.batch_size]
                self._tiles = []
                rest = None
                for b0, b1 in zip(bounds[:-1], bounds[1:]):
                    if b1 - b0 == TC_MAX_BATCH:
                        th = probe
                    else:
                        rest = rest if rest is not None else create(b1 - b0)
                        th = rest
                    self._tiles.append({"b0": b0, "b1": b1, "handle": th, "reserve": None})
                self._tile_handles = [probe] + ([rest] if rest is not None else [])
                h = probe
            else:
                _lib.raw("rs_am_destroy")(probe)
        if h is None:
            h = create(self.batch_size)
        self._handle = h
        self.uses_tensor_cores = bool(_lib.raw("rs_am_uses_tensor_cores")(h))
        n = _lib.raw("rs_am_param_count")(h)
        self.n_params = int(n)
        self.params = torch.zeros(self.n_params, dtype=torch.float32, device=self.device)
        ws_bytes = max(int(_lib.raw("rs_am_workspace_bytes")(th)) for th in (getattr(self, "_tile_handles", None) or [h]))
        self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        self.rnn_state = torch.zeros((self.num_layers, 2, self.batch_size, self.hidden_size), dtype=torch.float32,
                                     device=self.device)
        self._ctc_ws = None
        # version of the parameter buffer (rs_am_set_params_version): the kernels' weight planes are re-packed when it
        # changes -- once per optimizer step -- instead of on every forward / backward call.  Every method of this
        # class that writes self.params bumps it; code that writes the buffer directly calls params_changed().  With
        # batch tiles of two sizes the tiles' handles lay the shared workspace out differently: no caching then.
        self._params_version = 1
        self._cache_planes = len(getattr(self, "_tile_handles", None) or [h]) == 1
        self.rnn_created = True

    def create_forward_rnn(self):
        """Create the forward-only RNN (models/AcousticModel.py:96-120)."""
        self._create_common()

    def create_training_rnn(self, input_keep_prob, output_keep_prob, grad_clip, learning_rate, lr_decay_factor,
                            use_iterator=False):
        """Create the training RNN (models/AcousticModel.py:122-187)."""
        self._create_common()
        self.input_keep_prob = float(input_keep_prob)
        self.output_keep_prob = float(output_keep_prob)
        self.grad_clip = float(grad_clip)
        self.learning_rate_var = float(learning_rate)
        self.lr_decay_factor = float(lr_decay_factor)
        self.use_iterator = use_iterator
        h = self._handle
        if self._tiles is None:
            self._reserve = torch.empty(int(_lib.raw("rs_am_reserve_bytes")(h)), dtype=torch.uint8, device=self.device)
        else:
            self._reserve = None
            for t in self._tiles:       # one backward per forward PER TILE: each tile keeps its own activations
                t["reserve"] = torch.empty(int(_lib.raw("rs_am_reserve_bytes")(t["handle"])), dtype=torch.uint8,
                                           device=self.device)
        self.grads = torch.zeros_like(self.params)
        self.adam_m = torch.zeros_like(self.params)
        self.adam_v = torch.zeros_like(self.params)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=self.device)
        # accumulators (models/AcousticModel.py:364-383): [mean_loss, error_rate, mini_batch, ranks out of data]
        self._acc = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._one = torch.ones(1, dtype=torch.float32, device=self.device)
        self.training_created = True

    def __del__(self):
        try:
            if self._handle is not None:
                for th in (getattr(self, "_tile_handles", None) or [self._handle]):
                    _lib.raw("rs_am_destroy")(th)
                self._handle = None
        except Exception:
            pass

    # ------------------------------------------------------------- parameters
    def param_views(self):
        """name -> view into the flat parameter buffer, with the reference's
        checkpoint variable names (models/AcousticModel.py:515-527)."""
        h, H, F, C = self._handle, self.hidden_size, self.input_dim, self.num_labels
        off = lambda which, layer=0: int(_lib.raw("rs_am_param_offset")(h, which, layer))
        views = {}

        def view(buf, o, shape):
            n = int(np.prod(shape))
            return buf[o:o + n].view(*shape)
        def build(buf):
            out = {"Input_Layer/input_w": view(buf, off(0), (F, H)), "Input_Layer/input_b": view(buf, off(1), (H,))}
            for l in range(self.num_layers):
                out["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % l] = view(buf, off(2, l), (2 * H, 4 * H))
                out["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % l] = view(buf, off(3, l), (4 * H,))
            out["Output_layer/output_w"] = view(buf, off(4), (H, C))
            out["Output_layer/output_b"] = view(buf, off(5), (C,))
            return out
        views = build(self.params)
        return views

    def grad_views(self):
        saved = self.params
        try:
            self.params = self.grads
            return self.param_views()
        finally:
            self.params = saved

    def params_changed(self):
        """Tell the library that self.params was modified (see _create_common)."""
        self._params_version += 1

    def _sync_params_version(self):
        v = self._params_version if self._cache_planes else 0
        for th in (getattr(self, "_tile_handles", None) or [self._handle]):
            _lib.call("rs_am_set_params_version", th, v)

    def initialize(self, sess=None):
        """Xavier / glorot-uniform weights, zero biases (models/AcousticModel.py:242-245,
        :303-306, BasicLSTMCell defaults); deterministic in self.seed."""
        gen = torch.Generator(device="cpu")
        gen.manual_seed(int(self.seed))
        for name, v in self.param_views().items():
            if v.dim() == 2:
                lim = float(np.sqrt(6.0 / (v.shape[0] + v.shape[1])))
                w = (torch.rand(v.shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * lim
                v.copy_(w.to(self.device))
            else:
                v.zero_()
        self.global_step = 0
        self.rnn_state.zero_()
        self.params_changed()

    def load_flat_params(self, flat):
        flat = torch.as_tensor(np.asarray(flat, dtype=np.float32))
        assert flat.numel() == self.n_params
        self.params.copy_(flat.to(self.device))
        self.params_changed()

    def get_learning_rate(self):
        return self.learning_rate_var

    def set_learning_rate(self, sess, learning_rate):
        self.learning_rate_var = float(learning_rate)

    def learning_rate_decay_op(self, sess=None):
        """models/AcousticModel.py:354 -- learning_rate *= lr_decay_factor"""
        self.learning_rate_var = self.learning_rate_var * self.lr_decay_factor
        return self.learning_rate_var

    def set_is_training(self, sess, is_training):
        self.is_training = bool(is_training)

    # ----------------------------------------------------------- checkpointing
    def save(self, session, checkpoint_dir, fmt="npz"):
        """models/AcousticModel.py:483-487.  Writes the same variable set (weights, global_step, learning_rate; no
        Adam slots, no RNN state) under the TF checkpoint prefix: fmt="npz" a numpy archive, fmt="tf" the reference's
        own tensor-bundle format (acousticmodel.ckpt-N.index / .data-00000-of-00001, tf_checkpoint.write_bundle)."""
        arrays = {k: v.detach().cpu().numpy() for k, v in self.param_views().items()}
        arrays["global_step"] = np.array(self.global_step, np.int32)
        arrays["learning_rate"] = np.array(self.learning_rate_var if self.learning_rate_var is not None else 0.0, np.float32)
        # every file is written under a temporary name and renamed: a reader (or a crash) never sees half a checkpoint
        if fmt == "tf":
            from .tf_checkpoint import write_bundle
            name = "acousticmodel.ckpt-%d" % self.global_step
            path = os.path.join(checkpoint_dir, name)
            tmp = os.path.join(checkpoint_dir, ".tmp-" + name)
            write_bundle(tmp, arrays, accel=_lib.raw("rs_crc32c"))
            for suffix in (".data-00000-of-00001", ".index"):
                os.replace(tmp + suffix, path + suffix)
        else:
            name = "acousticmodel.ckpt-%d.npz" % self.global_step
            path = os.path.join(checkpoint_dir, name)
            tmp = os.path.join(checkpoint_dir, ".tmp-" + name)
            with open(tmp, "wb") as fh:
                np.savez(fh, **arrays)
            os.replace(tmp, path)
        state_tmp = os.path.join(checkpoint_dir, ".tmp-checkpoint")
        with open(state_tmp, "w") as fh:
            fh.write('model_checkpoint_path: "%s"\n' % name)
        os.replace(state_tmp, os.path.join(checkpoint_dir, "checkpoint"))
        logging.info("Checkpoint saved")
        return path

    def restore(self, session, checkpoint_dir):
        """models/AcousticModel.py:489-499.  Reads this library's .npz archives and the reference's own TensorFlow
        checkpoints (tensor bundle, or the constants embedded in the shipped .meta: tf_checkpoint.py)."""
        state = os.path.join(checkpoint_dir, "checkpoint")
        if os.path.exists(state):
            line = open(state).readline()
            name = line.split('"')[1]
            path = name if os.path.isabs(name) else os.path.join(checkpoint_dir, name)
            if name.endswith(".npz"):
                data = np.load(path)
                source = "npz"
            else:
                from .tf_checkpoint import load_reference_checkpoint
                data, source = load_reference_checkpoint(path)
            for k, v in self.param_views().items():
                if tuple(data[k].shape) != tuple(v.shape):
                    raise ValueError("checkpoint variable %s has shape %s, the model expects %s"
                                     % (k, tuple(data[k].shape), tuple(v.shape)))
                v.copy_(torch.from_numpy(np.ascontiguousarray(data[k], dtype=np.float32)).to(self.device))
            self.global_step = int(np.asarray(data["global_step"]).reshape(-1)[0])
            self.params_changed()
            self._adam_step = 0
            if self.adam_m is not None:
                self.adam_m.zero_()
                self.adam_v.zero_()
            if self.learning_rate_var is not None:
                self.learning_rate_var = float(np.asarray(data["learning_rate"]).reshape(-1)[0])
            logging.info("Restored model parameters from %s [%s] (global_step id %d)", name, source, self.global_step)
        else:
            logging.info("Created model with fresh parameters.")

    # -------------------------------------------------------------- hot path
    def _labels_to_device(self, label_rows):
        lens = np.array([len(r) for r in label_rows], dtype=np.int32)
        offs = np.zeros(len(label_rows) + 1, dtype=np.int32)
        np.cumsum(lens, out=offs[1:])
        flat = np.concatenate([np.asarray(r, dtype=np.int32) for r in label_rows]) if offs[-1] > 0 \
            else np.zeros((1,), np.int32)
        if flat.size and (flat.min() < 0 or flat.max() >= self.num_labels):
            raise ValueError("labels must be in [0, num_labels)")
        return (torch.from_numpy(flat).to(self.device, non_blocking=True),
                torch.from_numpy(offs).to(self.device, non_blocking=True), int(lens.max()) if len(lens) else 0)

    def sparse_labels_from_dense(self, dense_labels, fill_empty):
        """models/AcousticModel.py:151-156 / :174-178: drop every 0 of the dense
        zero-padded label batch; (iterator path) fill empty rows with [num_labels-1]."""
        dense = np.asarray(dense_labels)
        rows = [row[row != 0].astype(np.int32) for row in dense]
        while fill_empty and len(rows) < self.batch_size:
            rows.append(np.zeros((0,), np.int32))
        if fill_empty:
            rows = [r if len(r) else np.array([self.num_labels - 1], np.int32) for r in rows]
        return rows

    def forward(self, x_d, len_d, training=False, keep_state=True, T=None, logits=None):
        """x_d float32 [T,B,F] (device), len_d int32 [B].  Returns logits [T,B,C].
        training=True keeps the activations for one backward call and applies the
        configured dropout."""
        assert x_d.dtype == torch.float32 and x_d.is_contiguous() and x_d.dim() == 3
        T = int(x_d.shape[0]) if T is None else int(T)
        assert x_d.shape[1] == self.batch_size and x_d.shape[2] == self.input_dim
        if logits is None:
            logits = torch.empty((T, self.batch_size, self.num_labels), dtype=torch.float32, device=self.device)
        keep_in = self.input_keep_prob if training else 1.0
        keep_out = self.output_keep_prob if training else 1.0
        self._dropout_calls += 1
        # fresh masks per call, and per rank under data parallelism (N ranks = N independent mini-batches)
        seed = (int(self.seed) * 1000003 + self._dropout_calls + dist_world()[0] * 0x632BE59BD9B4E019) & 0xFFFFFFFFFFFFFFFF
        self._last_fwd = (keep_in, keep_out, seed, T)
        self._sync_params_version()
        if self._tiles is None:
            _lib.call("rs_am_forward", self._handle, self.params.data_ptr(), x_d.data_ptr(), len_d.data_ptr(), T,
                      self.rnn_state.data_ptr(), self.rnn_state.data_ptr() if keep_state else None,
                      keep_in, keep_out, seed, logits.data_ptr(),
                      self._reserve.data_ptr() if training else None, self._ws.data_ptr(), self._ws.numel(),
                      _stream_ptr())
            return logits
        for i, t in enumerate(self._tiles):
            b0, b1 = t["b0"], t["b1"]
            x_t = x_d[:T, b0:b1].contiguous()
            len_t = len_d[b0:b1].contiguous()
            st_t = self.rnn_state[:, :, b0:b1].contiguous()
            lg_t = torch.empty((T, b1 - b0, self.num_labels), dtype=torch.float32, device=self.device)
            _lib.call("rs_am_forward", t["handle"], self.params.data_ptr(), x_t.data_ptr(), len_t.data_ptr(), T,
                      st_t.data_ptr(), st_t.data_ptr() if keep_state else None,
                      keep_in, keep_out, (seed + 0x9E3779B97F4A7C15 * (i + 1)) & 0xFFFFFFFFFFFFFFFF, lg_t.data_ptr(),
                      t["reserve"].data_ptr() if training else None, self._ws.data_ptr(), self._ws.numel(),
                      _stream_ptr())
            logits[:T, b0:b1] = lg_t
            if keep_state:
                self.rnn_state[:, :, b0:b1] = st_t
            t["x"], t["len"] = (x_t, len_t) if training else (None, None)
        return logits

    def backward(self, x_d, len_d, dlogits):
        keep_in, keep_out, seed, T = self._last_fwd
        self._sync_params_version()
        if self._tiles is None:
            _lib.call("rs_am_backward", self._handle, self.params.data_ptr(), x_d.data_ptr(), len_d.data_ptr(), T,
                      keep_in, keep_out, seed, dlogits.data_ptr(), self._reserve.data_ptr(), self.grads.data_ptr(),
                      self._ws.data_ptr(), self._ws.numel(), _stream_ptr())
            return
        for i, t in enumerate(self._tiles):
            b0, b1 = t["b0"], t["b1"]
            dl_t = dlogits[:T, b0:b1].contiguous()
            _lib.call("rs_am_backward", t["handle"], self.params.data_ptr(), t["x"].data_ptr(), t["len"].data_ptr(), T,
                      keep_in, keep_out, (seed + 0x9E3779B97F4A7C15 * (i + 1)) & 0xFFFFFFFFFFFFFFFF, dl_t.data_ptr(),
                      t["reserve"].data_ptr(), self.grads.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                      _stream_ptr())

    def ctc_loss(self, logits, label_rows, len_d, want_grad=True):
        """tf.nn.ctc_loss(..., ignore_longer_outputs_than_inputs=True) + gradient.
        Returns (loss [B] device, grad [T,B,C] device or None)."""
        T, B, C = logits.shape
        flat, offs, maxlen = self._labels_to_device(label_rows)
        need = int(_lib.raw("rs_ctc_workspace_bytes")(T, B, C, maxlen))
        if self._ctc_ws is None or self._ctc_ws.numel() < need:
            self._ctc_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        loss = torch.empty((B,), dtype=torch.float32, device=self.device)
        grad = torch.empty_like(logits) if want_grad else None
        _lib.call("rs_ctc_loss_grad", logits.data_ptr(), flat.data_ptr(), offs.data_ptr(), len_d.data_ptr(), T, B, C,
                  maxlen, self.num_labels - 1, self.beta_skip, loss.data_ptr(),
                  grad.data_ptr() if want_grad else None, self._ctc_ws.data_ptr(), self._ctc_ws.numel(),
                  _stream_ptr())
        return loss, grad

    def greedy_decode(self, logits, len_d):
        """Per-frame argmax, collapse repeats, drop blank.  Returns (ids [B,T] int32
        padded with -1, lengths [B]) on the device."""
        T, B, C = logits.shape
        out = torch.empty((B, T), dtype=torch.int32, device=self.device)
        out_len = torch.empty((B,), dtype=torch.int32, device=self.device)
        _lib.call("rs_ctc_greedy_decode", logits.data_ptr(), len_d.data_ptr(), T, B, C, self.num_labels - 1,
                  out.data_ptr(), out_len.data_ptr(), _stream_ptr())
        return out, out_len

    def beam_search_decode(self, logits, len_d, beam_width=100, merge_repeated=True, normalize=True):
        """tf.nn.ctc_beam_search_decoder(logits, seq_len) with its defaults (models/AcousticModel.py:312): the top
        path per item.  Returns (ids [B,T] int32 padded with -1, lengths [B], log scores [B]) on the device."""
        T, B, C = logits.shape
        out = torch.empty((B, T), dtype=torch.int32, device=self.device)
        out_len = torch.empty((B,), dtype=torch.int32, device=self.device)
        score = torch.empty((B,), dtype=torch.float32, device=self.device)
        need = int(_lib.raw("rs_ctc_beam_workspace_bytes")(T, B))
        if getattr(self, "_beam_ws", None) is None or self._beam_ws.numel() < need:
            self._beam_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        _lib.call("rs_ctc_beam_search", logits.data_ptr(), len_d.data_ptr(), T, B, C, int(beam_width),
                  1 if merge_repeated else 0, 1 if normalize else 0, out.data_ptr(), out_len.data_ptr(),
                  score.data_ptr(), self._beam_ws.data_ptr(), self._beam_ws.numel(), _stream_ptr())
        return out, out_len, score

    def predict(self, logits, len_d):
        """The reference's `prediction` (models/AcousticModel.py:312-314): beam search, width 100, top path;
        self.decoder = "greedy" selects the per-frame argmax path instead."""
        if self.decoder == "greedy":
            return self.greedy_decode(logits, len_d)
        ids, lens, _ = self.beam_search_decode(logits, len_d)
        return ids, lens

    def edit_distance(self, ids, lens, label_rows):
        """Levenshtein distance of every decoded row to its truth, on the device.  Returns (distance int32 [B],
        rate float32 [B] = distance / len(truth))."""
        flat, offs, maxlen = self._labels_to_device(label_rows)
        B = int(ids.shape[0])
        dist = torch.empty((B,), dtype=torch.int32, device=self.device)
        rate = torch.empty((B,), dtype=torch.float32, device=self.device)
        _lib.call("rs_edit_distance", ids.data_ptr(), lens.data_ptr(), int(ids.shape[1]), flat.data_ptr(), offs.data_ptr(),
                  B, int(maxlen), dist.data_ptr(), rate.data_ptr(), _stream_ptr())
        return dist, rate

    def _error_rate(self, logits, len_d, label_rows):
        """mean over the batch of edit_distance(prediction, truth) / len(truth)
        (tf.edit_distance(normalize=True), models/AcousticModel.py:370): decoder and distance on the device, the
        result stays there (no host round trip inside the step)."""
        ids, lens = self.predict(logits, len_d)
        _, rate = self.edit_distance(ids, lens, label_rows)
        return rate

    def _accumulate_error_rate_async(self, logits, len_d, label_rows):
        """acc_error_rate_op (models/AcousticModel.py:370-376, fetched by every run_step :641): the prediction (beam
        search) and the edit distance only need the logits, so they run on a low-priority side stream while the
        caller's stream goes on with the CTC loss and the backward pass; end_batch waits for them."""
        cur = torch.cuda.current_stream(self.device)
        if self._err_stream is None:
            self._err_stream = torch.cuda.Stream(device=self.device, priority=0)
        side = self._err_stream
        ready = torch.cuda.Event()
        ready.record(cur)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            rate = self._error_rate(logits, len_d, label_rows)
            _lib.call("rs_accumulate_mean", rate.data_ptr(), None, int(rate.numel()), self._acc.data_ptr() + 4, None,
                      side.cuda_stream)
            done = torch.cuda.Event()
            done.record(side)
        logits.record_stream(side)
        len_d.record_stream(side)
        self._err_event = done

    def _phase_mark(self, name):
        ev = getattr(self, "_phase_events", None)
        if ev is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(self.device))
            ev.append((name, e))

    def step_on_batch(self, x_d, len_d, label_rows, compute_gradients=True, compute_error_rate=True):
        """One mini-batch of run_step (models/AcousticModel.py:634-660) on explicit
        device tensors: forward, CTC, (backward + gradient accumulation),
        accumulate mean loss / error rate / mini-batch count, keep the RNN state."""
        self._phase_mark("forward")
        logits = self.forward(x_d, len_d, training=compute_gradients, keep_state=True)
        if compute_error_rate:
            self._accumulate_error_rate_async(logits, len_d, label_rows)
        self._phase_mark("ctc")
        loss, grad = self.ctc_loss(logits, label_rows, len_d, want_grad=compute_gradients)
        self._phase_mark("bookkeeping")
        # display loss: mean(loss[b] / len[b]), and the mini-batch counter        (models/AcousticModel.py:361-383).
        # In FRONT of the backward pass: the accumulators are final once the CTC kernel has run, and end_batch reads
        # them from a side stream behind this event while the backward pass is still running.
        _lib.call("rs_accumulate_mean", loss.data_ptr(), len_d.data_ptr(), int(loss.numel()), self._acc.data_ptr(),
                  self._acc.data_ptr() + 8, _stream_ptr())
        if self._acc_event is None:
            self._acc_event = torch.cuda.Event()
        self._acc_event.record(torch.cuda.current_stream(self.device))
        self._acc_recorded = True
        self._phase_mark("backward")
        if compute_gradients:
            self.backward(x_d, len_d, grad)
        self._mini_batches += 1
        return loss

    def apply_gradients(self):
        """train_step_op (models/AcousticModel.py:404-406): all-reduce (data parallel),
        clip the ACCUMULATED gradient by its global norm, Adam, global_step += 1."""
        self._phase_mark("allreduce")
        allreduce_sum_(self.grads)
        self._phase_mark("clip_adam")
        self.global_step += 1
        self._adam_step += 1
        _lib.call("rs_sumsq", self.grads.data_ptr(), self.n_params, self._sumsq.data_ptr(), _stream_ptr())
        _lib.call("rs_clip_adam_step", self.params.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(),
                  self.adam_v.data_ptr(), self.n_params, self._sumsq.data_ptr(), self.grad_clip,
                  self.learning_rate_var, 0.9, 0.999, 1e-8, self._adam_step, _stream_ptr())
        self.params_changed()
        self._phase_mark("end")

    # ---------------------------------------------------------- measurement
    def enable_timing(self):
        """CUDA events on the launch stream around every recurrent kernel."""
        _lib.call("rs_am_enable_timing", self._handle, 1)

    def recurrent_ms(self):
        """(forward_ms, backward_ms): per-layer durations of the last step's recurrent kernels."""
        out = ([], [])
        for d in (0, 1):
            for l in range(self.num_layers):
                ms = _lib.ctypes.c_float()
                rc = _lib.raw("rs_am_recurrent_ms")(self._handle, d, l, _lib.ctypes.byref(ms))
                out[d].append(ms.value if rc == 0 else 0.0)       # (a direction that has not run: inference)
        return out

    def recurrent_trace(self, max_launches=256):
        """[direction][layer] -> list of (start_ms, stop_ms) of every recurrent launch of the last step,
        measured from the top of the forward / backward call (the pipelined schedule's timeline)."""
        out = ([], [], [], [])       # forward rec, backward rec, backward weight-gradient GEMM pairs, backward dx GEMMs
        buf = (_lib.ctypes.c_float * (2 * max_launches))()
        for d in (0, 1, 2, 3):
            for l in range(self.num_layers):
                n = _lib.raw("rs_am_recurrent_trace")(self._handle, d, l, buf, max_launches)
                if n < 0:
                    _lib.check(n)
                out[d].append([(buf[2 * i], buf[2 * i + 1]) for i in range(n)])
        return out

    # ------------------------------------------------------- step protocol
    def start_batch(self, session=None, is_training=True, run_options=None, run_metadata=None):
        """models/AcousticModel.py:662-670"""
        _lib.call("rs_memset_zero", self._acc.data_ptr(), 16, _stream_ptr())
        self._mini_batches = 0
        self._dataset_empty = False
        self.set_is_training(session, is_training)
        if is_training:
            _lib.call("rs_memset_zero", self.grads.data_ptr(), 4 * self.n_params, _stream_ptr())

    def _next_batch(self):
        it = self._train_iter if self.is_training else self._valid_iter
        if it is None:
            raise RuntimeError("no dataset attached: call add_datasets_input / add_dataset_input first")
        try:
            return next(it)
        except StopIteration:
            raise OutOfRangeError()

    def run_step(self, session=None, compute_gradients=True, run_options=None, run_metadata=None,
                 compute_error_rate=True):
        """models/AcousticModel.py:634-660: one mini-batch pulled from the attached dataset.  Returns the number of
        mini-batches accumulated so far (kept on the host: no device read inside the step)."""
        start_time = time.time()
        x_d, len_d, dense_labels = self._next_batch()
        rows = self.sparse_labels_from_dense(dense_labels, fill_empty=True)
        # the NEXT mini-batch's feature kernels go in front of this forward pass (see BatchPrefetcher._Ticket.launch_features)
        launch = getattr(self._train_dataset if self.is_training else self._valid_dataset, "launch_pending_features", None)
        if launch is not None:
            launch()
        self.step_on_batch(x_d, len_d, rows, compute_gradients, compute_error_rate)
        logging.debug("Step duration : %.2f", time.time() - start_time)
        return float(self._mini_batches)

    def end_batch(self, session=None, is_training=True, run_options=None, run_metadata=None,
                  rnn_state_reset_ratio=1.0):
        """models/AcousticModel.py:672-703.  Data parallel: the four accumulators (mean loss, error rate, mini-batch
        count, ranks whose dataset ran dry) are summed over the ranks FIRST, so that every rank takes the same
        decision -- apply the (all-reduced) gradient iff any rank accumulated a mini-batch -- and the collectives
        of all ranks stay in lockstep whatever the shard sizes."""
        cur = torch.cuda.current_stream(self.device)
        n_ranks = dist_world()[1]
        early = None
        if os.environ.get("RS_EARLY_READ", "1") != "0" and (n_ranks == 1 or os.environ.get("RS_EARLY_READ_DP", "1") != "0"):
            # The accumulators are final right behind the last CTC kernel (step_on_batch records an event there), long
            # before the backward pass has finished.  Their all-reduce (data parallel) and their device-to-host copy run
            # on a side stream behind that event (and the decoder's, if it ran); the optimizer is enqueued behind the
            # backward pass meanwhile, and the call returns as soon as the copy has landed -- the backward pass and the
            # optimizer still running, the caller already enqueueing its next step -- instead of draining the device.
            # One process: the decision below needs nothing from the device (the mini-batch count is known here).
            if self._read_stream is None:
                self._read_stream = torch.cuda.Stream(device=self.device, priority=-1)
                self._acc_host = torch.empty(4, dtype=torch.float32).pin_memory()
                self._read_done = torch.cuda.Event()
            if self._acc_event is None:
                self._acc_event = torch.cuda.Event()
            rd = self._read_stream
            if not self._acc_recorded:
                self._acc_event.record(cur)               # no mini-batch here: behind start_batch's clear
            if n_ranks > 1 or self._mini_batches > 0:
                rd.wait_event(self._acc_event)
                if self._err_event is not None:
                    rd.wait_event(self._err_event)
                with torch.cuda.stream(rd):
                    if n_ranks > 1:
                        if self._dataset_empty:
                            _lib.call("rs_accumulate_mean", self._one.data_ptr(), None, 1, self._acc.data_ptr() + 12, None,
                                      _stream_ptr())
                        allreduce_sum_(self._acc)
                    self._acc_host.copy_(self._acc, non_blocking=True)
                    self._read_done.record(rd)
                early = self._read_done
            if n_ranks == 1:
                batchs_count = float(self._mini_batches)
                self._ranks_empty = 1 if self._dataset_empty else 0
            else:
                early.synchronize()
                acc = self._acc_host.numpy().copy()
                batchs_count = float(acc[2])
                self._ranks_empty = int(round(float(acc[3])))
            if self._err_event is not None:
                cur.wait_event(self._err_event)       # (the next start_batch clears the accumulators on this stream)
                self._err_event = None
            if early is not None:
                cur.wait_event(early)
        else:
            if self._err_event is not None:
                cur.wait_event(self._err_event)
                self._err_event = None
            if self._dataset_empty:
                _lib.call("rs_accumulate_mean", self._one.data_ptr(), None, 1, self._acc.data_ptr() + 12, None, _stream_ptr())
            acc = allreduce_sum_(self._acc).cpu().numpy()
            batchs_count = float(acc[2])
            self._ranks_empty = int(round(float(acc[3])))
        self._acc_recorded = False
        if is_training and batchs_count > 0:
            self.apply_gradients()
            # Reset the hidden state at the given random ratio (default to always)   (:681-682)
            if randint(1, int(1 // rnn_state_reset_ratio)) == 1:
                _lib.call("rs_memset_zero", self.rnn_state.data_ptr(), 4 * self.rnn_state.numel(), _stream_ptr())
        if batchs_count <= 0:
            return 0.0, 0.0, self.global_step
        if early is not None:
            early.synchronize()
            acc = self._acc_host.numpy().copy()
        mean_loss = acc[0] / batchs_count
        mean_error_rate = acc[1] / batchs_count
        return mean_loss, mean_error_rate, self.global_step

    def run_train_step(self, sess=None, mini_batch_size=1, rnn_state_reset_ratio=1.0, run_options=None,
                       run_metadata=None, compute_error_rate=True):
        """models/AcousticModel.py:887-939.  Returns (mean_loss, mean_error_rate,
        current_step, dataset_empty)."""
        start_time = time.time()
        dataset_empty = False
        self.start_batch(sess, True)
        mini_batch_num = 0
        try:
            for _ in range(mini_batch_size):
                mini_batch_num = self.run_step(sess, True, compute_error_rate=compute_error_rate)
        except OutOfRangeError:
            logging.debug("Dataset empty, exiting train step")
            dataset_empty = True
        self._dataset_empty = dataset_empty
        if mini_batch_num > 0 or dist_world()[1] > 1:
            # (data parallel: a rank without a mini-batch still joins the collectives, with a zero gradient, and
            #  every rank ends its epoch as soon as any rank's shard ran dry)
            mean_loss, mean_error_rate, current_step = self.end_batch(sess, True,
                                                                      rnn_state_reset_ratio=rnn_state_reset_ratio)
            dataset_empty = dataset_empty or self._ranks_empty > 0
            logging.info("Batch %d : loss %.5f - error_rate %.5f - duration %.2f",
                         current_step, mean_loss, mean_error_rate, time.time() - start_time)
            return mean_loss, mean_error_rate, current_step, dataset_empty
        return 0.0, 0.0, self.global_step, dataset_empty

    def run_evaluation(self, sess=None, run_options=None, run_metadata=None):
        """models/AcousticModel.py:779-799"""
        start_time = time.time()
        logging.info("Start evaluating...")
        self.start_batch(sess, False)
        if self._valid_dataset is not None:
            self._valid_iter = iter(self._valid_dataset)
        try:
            while True:
                self.run_step(sess, False)
        except OutOfRangeError:
            logging.debug("Dataset empty, exiting evaluation step")
        mean_loss, mean_error_rate, current_step = self.end_batch(sess, False, rnn_state_reset_ratio=1.0)
        # always reset the RNN state after evaluation (:793-795)
        _lib.call("rs_memset_zero", self.rnn_state.data_ptr(), 4 * self.rnn_state.numel(), _stream_ptr())
        logging.info("Evaluation at step %d : loss %.5f - error_rate %.5f - duration %.2f",
                     current_step, mean_loss, mean_error_rate, time.time() - start_time)
        return mean_loss, mean_error_rate, current_step

    def process_input(self, session, inputs, input_seq_lengths, run_options=None, run_metadata=None):
        """models/AcousticModel.py:705-721: forward only; returns int32 [B, maxdecoded]
        padded with num_labels (which get_labels_str drops).  The RNN state is NOT
        carried over (the reference does not run rnn_keep_state_op here)."""
        x = torch.as_tensor(np.ascontiguousarray(np.asarray(inputs, dtype=np.float32))).to(self.device)
        lens = torch.as_tensor(np.asarray(input_seq_lengths, dtype=np.int32)).to(self.device)
        logits = self.forward(x, lens, training=False, keep_state=False)
        ids, out_len = self.predict(logits, lens)
        ids = ids.cpu().numpy()
        out_len = out_len.cpu().numpy()
        width = int(out_len.max()) if len(out_len) else 0
        pred = np.full((len(out_len), width), self.num_labels, dtype=np.int32)
        for b, n in enumerate(out_len):
            pred[b, :n] = ids[b, :n]
        return pred

    # ------------------------------------------------------------ inference
    def infer_pcm_device(self, audio_processor, pcm_d, offsets_d, batch, max_samples, sr, decoder="greedy"):
        """The stt.py --file / --evaluate path for a whole batch with the PCM already on the device
        (stt.py:239-264, models/AcousticModel.py:705-777): feature kernels -> forward -> decode, batch tile by batch
        tile (no gather / scatter copies between the stages: every tile's features, logits and decoded rows are
        produced where the next stage reads them).  The RNN state is not carried (process_input semantics).
        Returns (ids int32 [B, T] padded with -1, lengths int32 [B]) on the device."""
        assert batch == self.batch_size
        Tmax = int(self.max_input_seq_length)
        T = min(Tmax, audio_processor.num_frames(int(max_samples), sr))
        ids = torch.empty((batch, T), dtype=torch.int32, device=self.device)
        out_len = torch.empty((batch,), dtype=torch.int32, device=self.device)
        tiles = self._tiles if self._tiles is not None else [{"b0": 0, "b1": batch, "handle": self._handle}]
        clamp = audio_processor.num_frames(int(max_samples), sr) > Tmax
        self._sync_params_version()
        for t in tiles:
            b0, b1 = t["b0"], t["b1"]
            bt = b1 - b0
            feats, nframes = audio_processor.features_device(pcm_d, offsets_d[b0:b1 + 1], bt, max_samples, sr, time_major=True)
            lens = torch.clamp(nframes, max=Tmax) if clamp else nframes
            logits = torch.empty((T, bt, self.num_labels), dtype=torch.float32, device=self.device)
            _lib.call("rs_am_forward", t["handle"], self.params.data_ptr(), feats.data_ptr(), lens.data_ptr(), T,
                      None, None, 1.0, 1.0, 0, logits.data_ptr(), None, self._ws.data_ptr(), self._ws.numel(), _stream_ptr())
            if decoder == "greedy":
                _lib.call("rs_ctc_greedy_decode", logits.data_ptr(), lens.data_ptr(), T, bt, self.num_labels,
                          self.num_labels - 1, ids[b0:b1].data_ptr(), out_len[b0:b1].data_ptr(), _stream_ptr())
            else:
                need = int(_lib.raw("rs_ctc_beam_workspace_bytes")(T, bt))
                if getattr(self, "_beam_ws", None) is None or self._beam_ws.numel() < need:
                    self._beam_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
                _lib.call("rs_ctc_beam_search", logits.data_ptr(), lens.data_ptr(), T, bt, self.num_labels, 100, 1, 1,
                          ids[b0:b1].data_ptr(), out_len[b0:b1].data_ptr(), None, self._beam_ws.data_ptr(),
                          self._beam_ws.numel(), _stream_ptr())
        return ids, out_len

    def infer_signals(self, audio_processor, signals, sr, decoder="greedy"):
        """Host PCM in, decoded label ids on the host out (numpy int32 [B, T] padded with -1, lengths [B]): one pinned
        staging buffer and ONE host-to-device copy for the whole batch, then infer_pcm_device."""
        pcm_d, off_d, lens = audio_processor.stage_batch(signals, sr)
        ids, out_len = self.infer_pcm_device(audio_processor, pcm_d, off_d, len(signals), max(lens), sr, decoder=decoder)
        return ids.cpu().numpy(), out_len.cpu().numpy()

    def infer_ticket(self, audio_processor, ticket, sr, decoder="greedy"):
        """infer_signals for a batch that BatchPrefetcher.submit(signals, sr, defer_features=True) has already staged and
        copied on its own thread and stream: the staging and the copy of the next batch overlap this batch's kernels."""
        pcm_d, off_d, lens = ticket.staged()
        ids, out_len = self.infer_pcm_device(audio_processor, pcm_d, off_d, len(lens), max(lens), sr, decoder=decoder)
        return ids.cpu().numpy(), out_len.cpu().numpy()

    # ------------------------------------------------------------- datasets
    @staticmethod
    def build_dataset(input_set, batch_size, max_input_seq_length, max_target_seq_length,
                      signal_processing, char_map, sr=None, device=None, delta_mode="interp"):
        """models/AcousticModel.py:801-840.  Items are [audio, label, ...] where
        audio is a file name or an in-memory (signal, sample_rate) pair."""
        from .dataset import AudioBatchDataset
        return AudioBatchDataset(input_set, batch_size, max_input_seq_length, max_target_seq_length,
                                 signal_processing, char_map, device=device, delta_mode=delta_mode)

    def add_dataset_input(self, dataset):
        """models/AcousticModel.py:842-853"""
        self._train_dataset = self._valid_dataset = dataset
        self._train_iter = self._valid_iter = iter(dataset)
        return dataset

    def add_datasets_input(self, train_dataset, valid_dataset):
        """models/AcousticModel.py:855-871"""
        self._train_dataset, self._valid_dataset = train_dataset, valid_dataset
        self._train_iter, self._valid_iter = iter(train_dataset), iter(valid_dataset)
        return train_dataset, valid_dataset

    def reset_train_iterator(self):
        """stands in for sess.run(t_iterator.initializer) (stt.py:193-195)"""
        self._train_iter = iter(self._train_dataset)

    def add_tensorboard(self, session, tensorboard_dir, tb_run_name=None, timeline_enabled=False):
        """models/AcousticModel.py:409-465: kept as a no-op hook (scalar logging goes
        through `logging`; per-phase timing through --timeline JSON)."""
        self.tensorboard_dir = tensorboard_dir
        self.timeline_enabled = timeline_enabled

    # --------------------------------------------------------------- metrics
    @staticmethod
    def calculate_wer(first_string, second_string):
        """Word-level Levenshtein distance (models/AcousticModel.py:529-580).
        > calculate_wer("who is there", "is there") == 1"""
        r, h = first_string.split(), second_string.split()
        vocab = {w: i for i, w in enumerate(set(r) | set(h))}
        return levenshtein([vocab[w] for w in r], [vocab[w] for w in h])

    @staticmethod
    def calculate_cer(first_string, second_string):
        """Character-level Levenshtein distance ignoring spaces (models/AcousticModel.py:582-632).
        > calculate_cer("who is there", "who i thre") == 2"""
        r = [ord(c) for c in first_string.replace(" ", "")]
        h = [ord(c) for c in second_string.replace(" ", "")]
        return levenshtein(r, h)

    def _iter_features(self, eval_dataset, audio_processor):
        """(features [T', F], pre-truncation length) per item of eval_dataset, in order -- what the reference gets from
        process_audio_file one file at a time (models/AcousticModel.py:737-741).  Runs of file names are decoded,
        resampled and featurised batch_size at a time (one launch sequence and one device-to-host copy per chunk);
        in-memory (signal, sample_rate) items go through process_signal."""
        chunk = max(1, int(self.batch_size))
        i, n_items = 0, len(eval_dataset)
        while i < n_items:
            audio = eval_dataset[i][0]
            if isinstance(audio, (tuple, list)):
                yield audio_processor.process_signal(audio[0], audio[1])
                i += 1
                continue
            j = i
            while j < n_items and j - i < chunk and not isinstance(eval_dataset[j][0], (tuple, list)):
                j += 1
            feats, lens = audio_processor.process_audio_files([eval_dataset[k][0] for k in range(i, j)],
                                                              time_major=False)
            feats, lens = feats.cpu().numpy(), lens.cpu().numpy()
            for k in range(j - i):
                length = int(lens[k])
                yield feats[k, :min(length, int(audio_processor.max_input_seq_length))], length
            i = j

    def evaluate_full(self, sess, eval_dataset, input_seq_length, signal_processing, char_map,
                      run_options=None, run_metadata=None):
        """models/AcousticModel.py:723-777: WER / CER (percent) over a list of
        [audio, label, ...] items; audio = file name or (signal, sr)."""
        audio_processor = AudioProcessor(input_seq_length, signal_processing, device=self.device)
        wer_list, cer_list = [], []
        feats, lens, labs = [], [], []
        file_number = 0
        for item, (feat_vec, feat_len) in zip(eval_dataset, self._iter_features(eval_dataset, audio_processor)):
            audio, label = item[0], item[1]
            file_number += 1
            if len(label) > self.max_target_seq_length or feat_len > self.max_input_seq_length:
                logging.warning("Warning - sample too long : %s (input : %d / text : %s)", audio, feat_len, len(label))
            else:
                padded = np.zeros((self.max_input_seq_length, audio_processor.feature_size), np.float32)
                padded[:len(feat_vec)] = feat_vec
                feats.append(padded)
                lens.append(feat_len)
                labs.append(label)
            if file_number == len(eval_dataset):
                while len(feats) < self.batch_size and len(feats) > 0:
                    feats.append(np.zeros((self.max_input_seq_length, audio_processor.feature_size), np.float32))
                    lens.append(0)
                    labs.append("")
            if len(feats) == self.batch_size:
                batch = np.swapaxes(np.stack(feats), 0, 1)
                predictions = self.process_input(sess, batch, lens)
                for index, prediction in enumerate(predictions):
                    text = labelcodec.get_labels_str(char_map, prediction)
                    truth = labs[index]
                    if len(truth) > 0:
                        wer_list.append(self.calculate_wer(text, truth) / float(len(truth.split())))
                        cer_list.append(self.calculate_cer(text, truth) / float(len(truth.replace(" ", ""))))
                feats, lens, labs = [], [], []
        wer = (sum(wer_list) * 100) / float(len(wer_list))
        cer = (sum(cer_list) * 100) / float(len(cer_list))
        return wer, cer
