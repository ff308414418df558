# coding=utf-8
"""Hyper-parameter store: same config.ini keys, pickle file and
resume-versus-fork rule as /root/reference/util/hyperparams.py:16-141 (pure host
code, no arithmetic; kept so the stt.py / config.ini surface is a drop-in).
"""
import configparser
import logging
import os
import pickle
import time

ARCH_KEYS = ("num_layers", "hidden_size", "signal_processing", "language")


def read_config_file(config_file):
    """INI -> dict with the reference's keys and defaults (util/hyperparams.py:93-141)."""
    cfg = configparser.ConfigParser()
    if not cfg.read(config_file):
        raise FileNotFoundError(config_file)
    net, gen, trn, log = "acoustic_network_params", "general", "training", "logging"
    d = {}
    for key in ("num_layers", "hidden_size", "batch_size", "mini_batch_size", "grad_clip"):
        d[key] = cfg.getint(net, key)
    for key in ("dropout_input_keep_prob", "dropout_output_keep_prob", "learning_rate", "lr_decay_factor",
                "rnn_state_reset_ratio"):
        d[key] = cfg.getfloat(net, key)
    d["signal_processing"] = cfg.get(net, "signal_processing")
    d["language"] = cfg.get(net, "language")
    d["use_config_file_if_checkpoint_exists"] = cfg.getboolean(gen, "use_config_file_if_checkpoint_exists")
    d["steps_per_checkpoint"] = cfg.getint(gen, "steps_per_checkpoint")
    d["steps_per_evaluation"] = cfg.getint(gen, "steps_per_evaluation")
    d["checkpoint_dir"] = cfg.get(gen, "checkpoint_dir")
    d["training_dataset_dirs"] = cfg.get(trn, "training_dataset_dirs")
    d["training_filelist_cache"] = cfg.get(trn, "training_filelist_cache", fallback=None)
    d["test_dataset_dirs"] = cfg.get(trn, "test_dataset_dirs", fallback=None)
    d["train_frac"] = cfg.getfloat(trn, "train_frac", fallback=None)
    d["max_input_seq_length"] = cfg.getint(trn, "max_input_seq_length")
    d["max_target_seq_length"] = cfg.getint(trn, "max_target_seq_length")
    d["tensorboard_dir"] = cfg.get(trn, "tensorboard_dir", fallback=None)
    if d["tensorboard_dir"] is not None and not os.path.exists(d["tensorboard_dir"]):
        d["tensorboard_dir"] = None
    d["batch_normalization"] = cfg.getboolean(trn, "batch_normalization", fallback=False)
    d["dataset_size_ordering"] = cfg.get(trn, "dataset_size_ordering", fallback="False")
    if d["dataset_size_ordering"] not in ("True", "False", "First_run_only"):
        raise ValueError("dataset_size_ordering must be True, False or First_run_only")
    d["log_file"] = cfg.get(log, "log_file", fallback=None)
    level = cfg.get(log, "log_level", fallback="WARNING")
    d["log_level"] = getattr(logging, level, None)
    if not isinstance(d["log_level"], int):
        raise ValueError("Invalid log level: %s" % level)
    return d


class HyperParameterHandler(object):
    def __init__(self, config_file):
        """Reads the config file; if <checkpoint_dir>/hyperparams.p exists and the
        architecture keys differ, either restores the pickled values or forks a
        fresh checkpoint dir (util/hyperparams.py:17-57)."""
        self.hyper_params = self.read_config_file(config_file)
        if self.hyper_params["log_file"] is not None:
            logging.basicConfig(filename=self.hyper_params["log_file"])
        logging.getLogger().setLevel(self.hyper_params["log_level"])
        logging.info("Using checkpoint %s", self.hyper_params["checkpoint_dir"])
        os.makedirs(self.hyper_params["checkpoint_dir"], exist_ok=True)
        self.file_path = os.path.join(self.hyper_params["checkpoint_dir"], "hyperparams.p")
        if not self.check_exists():
            self.save_params(self.hyper_params)
            logging.info("No hyper params detected at checkpoint... reading config file")
        elif not self.check_changed(self.hyper_params):
            logging.info("No hyper parameter changed detected, using old checkpoint...")
        elif not self.hyper_params["use_config_file_if_checkpoint_exists"]:
            self.hyper_params = self.get_params()
            logging.info("Restoring hyper params from previous checkpoint...")
        else:
            fork = "{0}_hidden_size_{1}_numlayers_{2}_signal_processing_{3}".format(
                int(time.time()), self.hyper_params["hidden_size"], self.hyper_params["num_layers"],
                self.hyper_params["signal_processing"])
            fork = os.path.join(self.hyper_params["checkpoint_dir"], fork)
            os.makedirs(fork)
            self.hyper_params["checkpoint_dir"] = fork
            self.file_path = os.path.join(fork, "hyperparams.p")
            self.save_params(self.hyper_params)

    read_config_file = staticmethod(read_config_file)

    def get_hyper_params(self):
        return self.hyper_params

    def save_params(self, dic):
        with open(self.file_path, "wb") as handle:
            pickle.dump(dic, handle)

    def get_params(self):
        with open(self.file_path, "rb") as handle:
            return pickle.load(handle)

    def check_exists(self):
        return os.path.exists(self.file_path)

    def check_changed(self, new_params):
        if not self.check_exists():
            return False
        old = self.get_params()
        old.setdefault("signal_processing", "mfcc")     # old checkpoints
        old.setdefault("language", "")
        return any(old[k] != new_params[k] for k in ARCH_KEYS)
