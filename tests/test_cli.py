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


def test_repo_config_ini_parses(pkg):
    hp = pkg.HyperParameterHandler.read_config_file(os.path.join(ROOT, "config.ini"))
    assert hp["num_layers"] == 3 and hp["hidden_size"] == 768 and hp["signal_processing"] == "fbank"
