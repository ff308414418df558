"""The stt.py command line keeps the reference's flag surface (stt.py:360-404)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
import stt  # noqa: E402


def test_flag_surface_matches_reference():
    p = stt.parse_args(["--train_acoustic", "--config", "x.ini", "--max_epoch", "3", "--learn_rate", "0.01",
                        "--tb_name", "run", "--timeline", "--XLA"])
    assert p["train_acoustic"] and p["config_file"] == "x.ini" and p["max_epoch"] == 3 and p["learn_rate"] == 0.01
    assert p["tb_name"] == "run" and p["timeline"] and p["XLA"]
    assert stt.parse_args(["--train"])["train_acoustic"]                  # stale alias used by README / scripts
    assert stt.parse_args(["--file", "a.wav"])["file"] == "a.wav"
    assert stt.parse_args(["--evaluate"])["evaluate"]
    with pytest.raises(SystemExit):
        stt.parse_args(["--train_acoustic", "--evaluate"])                # mutually exclusive group
    with pytest.raises(SystemExit):
        stt.parse_args([])                                                # one mode is required


def test_out_of_scope_modes_exit_cleanly():
    for flag in ("--train_language", "--generate_text", "--record"):
        assert stt.main([flag]) == 2


def test_manifest_loader(tmp_path):
    d = tmp_path / "ds"
    d.mkdir()
    (d / "manifest.tsv").write_text("a.wav\thello world\n/abs/b.wav\tit's\n")
    items = stt.load_dataset_dirs(str(d))
    assert items[0][0] == str(d / "a.wav") and items[0][1] == "hello world"
    assert items[1][0] == "/abs/b.wav"
    syn = stt.synthetic_dataset(3, 0.5, 16000)
    assert len(syn) == 3 and len(syn[0][0][0]) == 8000 and isinstance(syn[0][1], str)


def test_librispeech_tree_with_flac_durations_and_ordering(pkg, tmp_path):
    """util/dataprocessor.py:263-278 (walker), :232-249 (durations), models/SpeechRecognizer.py:80-95 (order / split)."""
    import wave
    import numpy as np
    import flac_writer
    from rnn_speech_b200 import audiofile
    d = tmp_path / "LibriSpeech" / "train" / "19" / "198"
    d.mkdir(parents=True)
    (d / "19-198.trans.txt").write_text("19-198-0000 NORTHANGER ABBEY\n19-198-0001 THIS LITTLE WORK, WAS FINISHED\n"
                                        "19-198-0002 MISSING FILE\n19-198-0003 A WAV ONE\n")
    rng = np.random.default_rng(0)
    (d / "19-198-0000.flac").write_bytes(flac_writer.encode(rng.integers(-900, 900, 24000).astype(np.int16), 16000))
    (d / "19-198-0001.flac").write_bytes(flac_writer.encode(rng.integers(-900, 900, 8000).astype(np.int16), 16000))
    with wave.open(str(d / "19-198-0003.wav"), "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(8000)
        w.writeframes(rng.integers(-900, 900, 2 * 8000).astype("<i2").tobytes())
    items = stt.load_dataset_dirs(str(tmp_path / "LibriSpeech"), with_durations=True)
    assert [os.path.basename(i[0]) for i in items] == ["19-198-0000.flac", "19-198-0001.flac", "19-198-0003.wav"]
    assert [i[1] for i in items] == ["northanger abbey", "this little work was finished", "a wav one"]
    assert [i[2] for i in items] == [1.5, 0.5, 1.0]
    assert audiofile.duration_seconds(str(d / "19-198.trans.txt")) == 0          # unrecognised: like the mutagen miss
    train, test = stt.split_acoustic_dataset(list(items), [], True, 0.67)
    assert [i[2] for i in train] == [0.5, 1.0] and [i[2] for i in test] == [1.5]
    train, test = stt.split_acoustic_dataset(list(items), [], False, None)
    assert sorted(i[2] for i in train) == [0.5, 1.0, 1.5] and test == []


def test_repo_config_ini_parses(pkg):
    hp = pkg.HyperParameterHandler.read_config_file(os.path.join(ROOT, "config.ini"))
    assert hp["num_layers"] == 3 and hp["hidden_size"] == 768 and hp["signal_processing"] == "fbank"


def test_data_parallel_sharding_is_consistent_across_processes(pkg, tmp_path):
    """every rank must shard ONE order (an unseeded shuffle per process makes the shards overlap and leaks
    held-out items into other ranks' training sets), all shards must have the same size (same number of optimizer
    steps per rank), and manifest transcripts go through clean_label."""
    import subprocess
    import sys
    import stt
    from rnn_speech_b200 import dist as rsdist
    items = [["f%03d.wav" % i, "t%d" % i, float(i)] for i in range(65)]
    # the same permutation in two different interpreter processes (hash randomisation, no shared state)
    code = ("import sys; sys.path.insert(0, %r); import stt; "
            "print(','.join(x[0] for x in stt.shuffled([['f%%03d.wav' %% i, '', 0.0] for i in range(65)], 3)))" % ROOT)
    outs = {subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout for _ in range(2)}
    assert len(outs) == 1
    here = ",".join(x[0] for x in stt.shuffled(items, 3)) + "\n"
    assert outs == {here} and stt.shuffled(items, 3) != stt.shuffled(items, 4)
    train, test = stt.split_acoustic_dataset(list(items), [], False, 0.8)
    assert len(train) == 52 and len(test) == 13 and not {x[0] for x in train} & {x[0] for x in test}
    # equal, disjoint shards of the common order: 65 items on 2 ranks -> 32 + 32 (one item left out this epoch)
    shards = [rsdist.shard(train, r, 3) for r in range(3)]
    assert [len(s) for s in shards] == [17, 17, 17]
    assert len({x[0] for s in shards for x in s}) == 51
    # manifest transcripts are cleaned like the corpus walkers' (upper case / punctuation would end the label early)
    d = tmp_path / "m"
    d.mkdir()
    (d / "manifest.tsv").write_text("a.wav\tHello, World!\n")
    assert stt.load_dataset_dirs(str(d))[0][1] == "hello world"
