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
