"""Host-side logic that needs no GPU: config surface, metrics, sparse labels."""
import os
import textwrap

import numpy as np
import pytest


CONFIG = textwrap.dedent("""
    [acoustic_network_params]
    num_layers : 3
    hidden_size : 768
    dropout_input_keep_prob : 0.8
    dropout_output_keep_prob : 0.5
    batch_size : 32
    mini_batch_size : 1
    learning_rate : 0.0003
    lr_decay_factor : 0.33
    grad_clip : 1
    signal_processing : fbank
    language : english
    rnn_state_reset_ratio : 0.25
    [general]
    use_config_file_if_checkpoint_exists : True
    steps_per_checkpoint : 100
    steps_per_evaluation : 1000
    checkpoint_dir : %s
    [training]
    training_dataset_dirs : data/train
    max_input_seq_length : 1000
    max_target_seq_length : 600
    [logging]
    log_level : INFO
    """)


def test_hyperparams_same_keys_as_reference(pkg, tmp_path):
    ckpt = tmp_path / "ckpt"
    cfg = tmp_path / "config.ini"
    cfg.write_text(CONFIG % ckpt)
    hp = pkg.HyperParameterHandler(str(cfg)).get_hyper_params()
    expected = {"num_layers", "hidden_size", "dropout_input_keep_prob", "dropout_output_keep_prob", "batch_size",
                "mini_batch_size", "learning_rate", "lr_decay_factor", "grad_clip", "signal_processing", "language",
                "rnn_state_reset_ratio", "use_config_file_if_checkpoint_exists", "steps_per_checkpoint",
                "steps_per_evaluation", "checkpoint_dir", "training_dataset_dirs", "training_filelist_cache",
                "test_dataset_dirs", "train_frac", "max_input_seq_length", "max_target_seq_length",
                "tensorboard_dir", "batch_normalization", "dataset_size_ordering", "log_file", "log_level"}
    assert set(hp) == expected                       # util/hyperparams.py:105-137
    assert hp["hidden_size"] == 768 and hp["grad_clip"] == 1 and hp["batch_normalization"] is False
    assert os.path.exists(ckpt / "hyperparams.p")
    # unchanged architecture -> same dir; changed -> forked dir (util/hyperparams.py:36-52)
    assert pkg.HyperParameterHandler(str(cfg)).get_hyper_params()["checkpoint_dir"] == str(ckpt)
    cfg.write_text((CONFIG % ckpt).replace("hidden_size : 768", "hidden_size : 512"))
    forked = pkg.HyperParameterHandler(str(cfg)).get_hyper_params()["checkpoint_dir"]
    assert forked != str(ckpt) and "hidden_size_512_numlayers_3_signal_processing_fbank" in forked


def test_reference_config_ini_parses(pkg, tmp_path):
    ref = "/root/reference/config.ini"
    if not os.path.exists(ref):
        pytest.skip("reference not present")
    hp = pkg.HyperParameterHandler.read_config_file(ref)
    assert hp["num_layers"] == 5 and hp["hidden_size"] == 1024 and hp["signal_processing"] == "fbank"
    assert hp["max_input_seq_length"] == 3510 and hp["max_target_seq_length"] == 600


def test_wer_cer_docstring_examples(pkg):
    AM = pkg.AcousticModel
    assert AM.calculate_wer("who is there", "is there") == 1        # models/AcousticModel.py:548-555
    assert AM.calculate_wer("who is there", "") == 3
    assert AM.calculate_wer("", "who is there") == 3
    assert AM.calculate_cer("who is there", "whois there") == 0     # :601-606
    assert AM.calculate_cer("who is there", "who i thre") == 2
    assert AM.calculate_cer("", "who is there") == 10


def test_levenshtein_against_plain_dp(pkg):
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.integers(0, 5, size=rng.integers(0, 12))
        b = rng.integers(0, 5, size=rng.integers(0, 12))
        d = np.zeros((len(a) + 1, len(b) + 1), int)
        d[:, 0] = np.arange(len(a) + 1)
        d[0, :] = np.arange(len(b) + 1)
        for i in range(1, len(a) + 1):
            for j in range(1, len(b) + 1):
                d[i, j] = min(d[i - 1, j] + 1, d[i, j - 1] + 1, d[i - 1, j - 1] + (a[i - 1] != b[j - 1]))
        assert pkg.levenshtein(a, b) == d[-1, -1]


def test_mfcc_length_estimate(pkg):
    assert pkg.AudioProcessor.get_mfcc_length_from_duration(10.0) == int(10.0 // 0.01) - 1    # util/audioprocessor.py:29-39


def test_dataset_iteration_over_file_names_host_logic(pkg, monkeypatch):
    """AudioBatchDataset with file-name items (the reference's [audio_file, label, length] rows,
    models/AcousticModel.py:801-840): the file branch hands whole mini-batches to submit_files, pads the last batch
    with zero features / length 0 and builds zero-padded dense labels.  The device work is replaced by a stand-in
    so that the host logic runs on a CPU box."""
    import torch
    from rnn_speech_b200 import audioprocessor, dataset

    calls = []

    class FakeTicket(object):
        def __init__(self, value):
            self.value = value

        def result(self):
            return self.value

    class FakePrefetcher(object):
        def __init__(self, audio_processor):
            self.ap = audio_processor

        def submit_files(self, file_names, time_major=True):
            calls.append(list(file_names))
            n = len(file_names)
            feats = torch.stack([torch.full((20, 120), float(len(f))) for f in file_names], dim=1)
            lens = torch.tensor([10 + i for i in range(n)], dtype=torch.int32)
            assert time_major and feats.shape == (20, n, 120)
            return FakeTicket((feats, lens))

        def submit(self, *a, **k):
            raise AssertionError("in-memory branch taken for file items")

        def close(self):
            calls.append("closed")

    monkeypatch.setattr(audioprocessor, "BatchPrefetcher", FakePrefetcher)
    items = [["a.flac", "hello", 1.0], ["bb.wav", "it's", 2.0], ["ccc.flac", "we", None]]
    ds = dataset.AudioBatchDataset(items, 2, 20, 600, "fbank", pkg.ENGLISH_CHAR_MAP)
    assert len(ds) == 2
    batches = list(ds)
    assert calls == [["a.flac", "bb.wav"], ["ccc.flac"], "closed"]
    (f0, l0, d0), (f1, l1, d1) = batches
    assert f0.shape == (20, 2, 120) and list(l0) == [10, 11]
    assert f1.shape == (20, 2, 120) and list(l1) == [10, 0]                  # padded to batch_size, length 0
    assert float(f1[:, 0].min()) == 8.0 and float(f1[:, 1].abs().max()) == 0.0
    hello = pkg.get_str_labels(pkg.ENGLISH_CHAR_MAP, "hello")
    its = pkg.get_str_labels(pkg.ENGLISH_CHAR_MAP, "it's")
    assert d0.shape == (2, max(len(hello), len(its)))
    assert list(d0[0, :len(hello)]) == hello and list(d0[1, :len(its)]) == its
    assert d1.shape[0] == 1 and list(d1[0]) == pkg.get_str_labels(pkg.ENGLISH_CHAR_MAP, "we")


def test_dataset_defers_the_feature_kernels_of_in_memory_batches(pkg, monkeypatch):
    """In-memory items: the iterator submits the NEXT mini-batch with defer_features=True (the worker stages and copies,
    the feature kernels wait for launch_pending_features -- AcousticModel.run_step calls it in front of its forward pass --
    or for the next iteration's result()), and launch_pending_features is a no-op for tickets without that method (file
    batches) and when nothing is in flight.  Device work replaced by stand-ins."""
    import torch
    from rnn_speech_b200 import audioprocessor, dataset

    log = []

    class FakeTicket(object):
        def __init__(self, k, n):
            self.k, self.n, self.launched = k, n, 0

        def launch_features(self):
            self.launched += 1
            log.append(("launch", self.k))

        def result(self):
            log.append(("result", self.k))
            return torch.zeros((20, self.n, 120)), torch.full((self.n,), 12, dtype=torch.int32)

    class FakePrefetcher(object):
        def __init__(self, audio_processor):
            self.count = 0

        def submit(self, signals, sr, time_major=True, defer_features=False):
            assert defer_features and time_major and sr == 16000
            self.count += 1
            log.append(("submit", self.count))
            return FakeTicket(self.count, len(signals))

        def close(self):
            log.append("closed")

    monkeypatch.setattr(audioprocessor, "BatchPrefetcher", FakePrefetcher)
    sig = np.zeros(1600, np.float32)
    items = [[(sig, 16000), [1, 2, 3]] for _ in range(5)]
    ds = dataset.AudioBatchDataset(items, 2, 20, 600, "fbank", pkg.ENGLISH_CHAR_MAP)
    ds.launch_pending_features()                              # nothing in flight yet
    it = iter(ds)
    next(it)
    assert log == [("submit", 1), ("submit", 2), ("result", 1)]
    ds.launch_pending_features()                              # run_step, in front of the forward pass of batch 1
    assert log[-1] == ("launch", 2)
    next(it)
    assert log[-2:] == [("submit", 3), ("result", 2)]
    f, l, d = next(it)                                        # last batch: one real item, padded; nothing left in flight
    assert list(l) == [12, 0] and d.shape == (1, 3)
    n = len(log)
    ds.launch_pending_features()
    assert len(log) == n
    assert list(it) == [] and log[-1] == "closed"


def test_evaluate_full_host_logic_with_file_batches(pkg, monkeypatch):
    """evaluate_full (models/AcousticModel.py:723-777): files are featurised batch_size at a time, too-long samples
    are skipped, the last batch is padded with empty items, WER / CER are averaged in percent.  Device work is
    replaced by stand-ins so the control flow runs on a CPU box."""
    import types
    import torch
    from rnn_speech_b200 import acoustic_model

    extract_calls, forward_calls = [], []
    truth = {"a.flac": "hello world", "b.wav": "it's", "long.flac": "we", "c.flac": "coffee"}
    frames = {"a.flac": 12, "b.wav": 7, "long.flac": 99, "c.flac": 9}

    class FakeAP(object):
        def __init__(self, max_input_seq_length, feature_type="mfcc", device=None):
            self.max_input_seq_length, self.feature_size = max_input_seq_length, 120

        def process_audio_files(self, names, time_major=True):
            extract_calls.append(list(names))
            assert not time_major
            feats = torch.zeros((len(names), self.max_input_seq_length, 120))
            for i, n in enumerate(names):
                feats[i, :min(frames[n], self.max_input_seq_length)] = float(frames[n])
            return feats, torch.tensor([frames[n] for n in names], dtype=torch.int32)

        def process_signal(self, sig, sr):
            return np.full((5, 120), 5.0, np.float32), 5

    def process_input(sess, batch, lens):
        forward_calls.append((batch.shape, list(lens), [float(batch[0, b, 0]) for b in range(batch.shape[1])]))
        # item 0 of every batch is recognised perfectly, the others come out empty
        first = [k for k, v in frames.items() if v == lens[0]] or ["sig"]
        text = truth.get(first[0], "ab")
        ids = pkg.get_str_labels(pkg.ENGLISH_CHAR_MAP, text)[:-1]
        out = np.full((batch.shape[1], max(len(ids), 1)), 80, dtype=np.int32)
        out[0, :len(ids)] = ids
        return out

    monkeypatch.setattr(acoustic_model, "AudioProcessor", FakeAP)
    am = acoustic_model.AcousticModel
    fake = types.SimpleNamespace(batch_size=2, max_input_seq_length=20, max_target_seq_length=600, device=None,
                                 process_input=process_input, calculate_wer=am.calculate_wer,
                                 calculate_cer=am.calculate_cer)
    fake._iter_features = types.MethodType(am._iter_features, fake)
    data = [["a.flac", truth["a.flac"], None], ["b.wav", truth["b.wav"], None], ["long.flac", truth["long.flac"], None],
            [(np.zeros(10, np.float32), 16000), "ab", None], ["c.flac", truth["c.flac"], None]]
    wer, cer = am.evaluate_full(fake, None, data, 20, "fbank", pkg.ENGLISH_CHAR_MAP)
    assert extract_calls == [["a.flac", "b.wav"], ["long.flac"], ["c.flac"]]           # runs of files, batch_size at a time
    assert [c[1] for c in forward_calls] == [[12, 7], [5, 9]]                          # long.flac skipped (99 > 20 frames)
    assert forward_calls[0][0] == (20, 2, 120) and forward_calls[1][2] == [5.0, 9.0]
    # batch 1: "hello world" right, "it's" -> "" ; batch 2: "ab" right, "coffee" -> ""
    assert wer == pytest.approx(50.0) and cer == pytest.approx(50.0)
