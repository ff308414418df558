# coding=utf-8
"""
Main program to use the speech recognizer (B200 acoustic-model path).

Same command line and config.ini surface as the reference's stt.py
(/root/reference/stt.py:360-404): --train_acoustic (alias --train, which the reference's
README/scripts still use), --file, --evaluate, --config, --tb_name, --max_epoch,
--learn_rate, --timeline, --XLA (accepted, no-op).  --train_language / --record /
--generate_text are outside the accelerated path (the reference's language model is a stub,
--record needs audio hardware) and exit with a message.

Datasets: the reference's corpus walkers (util/dataprocessor.py) are out of scope; a dataset
directory here holds a `manifest.tsv` (one `<wav path>\t<transcript>` per line) or LibriSpeech
style `*.trans.txt` files next to .wav files.  `--synthetic N` trains on N seeded synthetic
utterances instead (no files needed).

Multi-GPU: launch with torchrun (one process per GPU); every rank trains on its own shard of
the dataset and the gradients are summed with one NCCL all-reduce per step.
"""
import argparse
import json
import logging
import os
import sys
import time
from random import Random

import numpy as np


def clean_label(_str):
    """util/dataprocessor.py:73-95"""
    _str = _str.strip().lower()
    for ch in ".,?!:":
        _str = _str.replace(ch, "")
    return _str.replace("-", " ").replace("_", " ").replace("  ", " ")


def load_dataset_dirs(dirs, with_durations=False):
    """[audio path, transcript, duration] items from a comma separated dir list: a `manifest.tsv`
    (path<TAB>text) or a LibriSpeech tree (`*.trans.txt` beside `.flac` files -- or `.wav` -- as the reference's
    walker expects, util/dataprocessor.py:263-278).  Durations (header reads, the reference's mutagen pass
    :232-249) are only filled in when the training set is ordered by size."""
    items = []
    for d in [x.strip() for x in dirs.split(",") if x.strip()]:
        manifest = os.path.join(d, "manifest.tsv")
        if os.path.exists(manifest):
            for line in open(manifest, encoding="utf-8"):
                line = line.rstrip("\n")
                if not line:
                    continue
                path, text = line.split("\t", 1)
                # transcripts go through the corpus walkers' clean_label (util/dataprocessor.py:73-95): upper case or
                # punctuation would otherwise end the label at the first unknown character
                items.append([path if os.path.isabs(path) else os.path.join(d, path), clean_label(text), None])
            continue
        for root, _, files in sorted(os.walk(d)):
            for f in sorted(files):
                if f.endswith(".trans.txt"):
                    for line in open(os.path.join(root, f), encoding="utf-8"):
                        key, _, text = line.strip().partition(" ")
                        for ext in (".flac", ".wav"):
                            audio = os.path.join(root, key + ext)
                            if os.path.exists(audio):
                                items.append([audio, clean_label(text), None])
                                break
    if with_durations:
        from rnn_speech_b200.audiofile import duration_seconds
        for item in items:
            item[2] = duration_seconds(item[0])
    return items


SHUFFLE_SEED = 20171103      # shared by every rank: the permutation must be the same in all processes


def shuffled(items, epoch):
    """The same permutation in every process (data parallel: the ranks shard ONE order; an unseeded shuffle per
    process would make the shards overlap and leak held-out items into other ranks' training sets)."""
    out = list(items)
    Random(SHUFFLE_SEED + epoch).shuffle(out)
    return out


def split_acoustic_dataset(train_set, test_set, ordered, train_frac):
    """models/SpeechRecognizer.py:80-95: order by duration or shuffle, then carve the test set out of the training
    set when no test directory is configured."""
    if ordered:
        train_set = sorted(train_set, key=lambda x: x[2] or 0)
    else:
        train_set = shuffled(train_set, 0)
    if not test_set and train_frac is not None:
        num_train = max(1, int(np.floor(train_frac * len(train_set))))
        train_set, test_set = train_set[:num_train], train_set[num_train:]
    return train_set, test_set


def synthetic_dataset(n_items, seconds, sr, seed=0):
    rng = np.random.default_rng(seed)
    words = ["the", "quick", "brown", "fox", "jumps", "over", "lazy", "dog", "it'll", "we've", "coffee", "mississippi"]
    items = []
    for _ in range(n_items):
        sig = (0.1 * rng.standard_normal(int(seconds * sr))).astype(np.float32)
        text = " ".join(rng.choice(words, size=rng.integers(3, 9)))
        items.append([(sig, sr), text, seconds])
    return items


def dist_setup():
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def build_acoustic_training_rnn(rs, hyper_params, prog_params, train_set, test_set, device):
    """stt.py:94-131 of the reference."""
    model = rs.AcousticModel(hyper_params["num_layers"], hyper_params["hidden_size"], hyper_params["batch_size"],
                             hyper_params["max_input_seq_length"], hyper_params["max_target_seq_length"],
                             hyper_params["input_dim"], hyper_params["batch_normalization"],
                             hyper_params["char_map_length"], device=device)
    build = lambda s: model.build_dataset(s, hyper_params["batch_size"], hyper_params["max_input_seq_length"],
                                          hyper_params["max_target_seq_length"], hyper_params["signal_processing"],
                                          hyper_params["char_map"], device=device)
    train_dataset = build(train_set)
    if not test_set:
        model.add_dataset_input(train_dataset)
        has_valid = False
    else:
        model.add_datasets_input(train_dataset, build(test_set))
        has_valid = True
    model.create_training_rnn(hyper_params["dropout_input_keep_prob"], hyper_params["dropout_output_keep_prob"],
                              hyper_params["grad_clip"], hyper_params["learning_rate"],
                              hyper_params["lr_decay_factor"], use_iterator=True)
    model.add_tensorboard(None, hyper_params["tensorboard_dir"], prog_params["tb_name"], prog_params["timeline"])
    model.initialize(None)
    model.restore(None, os.path.join(hyper_params["checkpoint_dir"], "acoustic"))
    if prog_params["learn_rate"] is not None:
        model.set_learning_rate(None, prog_params["learn_rate"])
    return model, build, has_valid


def train_acoustic_rnn(rs, train_set, test_set, hyper_params, prog_params):
    """stt.py:171-235 of the reference: checkpoint / evaluation cadence and the
    'seven non-improving checkpoint windows -> decay the learning rate' rule."""
    rank, world, local = dist_setup()
    import torch
    device = torch.device("cuda", local)
    from rnn_speech_b200 import dist as rsdist
    global_train_set = train_set
    # data parallel: every rank owns a shard of the SAME global order, all shards of equal size (lockstep collectives)
    train_set = rsdist.shard(global_train_set, rank, world)
    ckpt_dir = os.path.join(hyper_params["checkpoint_dir"], "acoustic")
    os.makedirs(ckpt_dir, exist_ok=True)
    model, build, has_valid = build_acoustic_training_rnn(rs, hyper_params, prog_params, train_set, test_set, device)
    timeline = []
    previous_mean_error_rates = []
    current_step = epoch = 0
    while True:
        mean_error_rate = 0
        for _ in range(hyper_params["steps_per_checkpoint"]):
            t0 = time.time()
            _loss, step_err, current_step, dataset_empty = model.run_train_step(
                None, hyper_params["mini_batch_size"], hyper_params["rnn_state_reset_ratio"])
            if prog_params["timeline"]:
                timeline.append({"step": int(current_step), "seconds": time.time() - t0, "loss": float(_loss)})
            mean_error_rate += step_err / hyper_params["steps_per_checkpoint"]
            if dataset_empty is True:
                epoch += 1
                logging.info("End of epoch number : %d", epoch)
                if (prog_params["max_epoch"] is not None) and (epoch > prog_params["max_epoch"]):
                    logging.info("Max number of epochs reached, exiting train step")
                    break
                if hyper_params["dataset_size_ordering"] in ['False', 'First_run_only']:
                    logging.info("Shuffling the training dataset")
                    global_train_set = shuffled(global_train_set, epoch)          # same order on every rank, re-sharded
                    train_set = rsdist.shard(global_train_set, rank, world)
                    model._train_dataset = build(train_set)
                else:
                    logging.info("Reuse the same training dataset")
                model.reset_train_iterator()
        if rank == 0:
            model.save(None, ckpt_dir)
        if (current_step % hyper_params["steps_per_evaluation"] == 0) and has_valid:
            model.run_evaluation(None)
        if mean_error_rate <= min(previous_mean_error_rates, default=sys.maxsize):
            previous_mean_error_rates.clear()
        previous_mean_error_rates.append(mean_error_rate)
        if len(previous_mean_error_rates) >= 7:
            model.learning_rate_decay_op()
            previous_mean_error_rates.clear()
            logging.info("Model is not improving, decaying the learning rate")
            if model.get_learning_rate() < 1e-7:
                logging.info("Learning rate is too low, exiting")
                break
            if rank == 0:
                model.save(None, ckpt_dir)
            logging.info("Overwriting the checkpoint file with the new learning rate")
        if (prog_params["max_epoch"] is not None) and (epoch > prog_params["max_epoch"]):
            logging.info("Max number of epochs reached, exiting training session")
            break
    if prog_params["timeline"] and rank == 0 and hyper_params["tensorboard_dir"]:
        with open(os.path.join(hyper_params["tensorboard_dir"], "timeline-train.json"), "w") as fh:
            json.dump(timeline, fh)


def process_file(rs, audio_processor, hyper_params, file):
    """stt.py:239-264 of the reference."""
    feat_vec, original_len = audio_processor.process_audio_file(file)
    if original_len > hyper_params["max_input_seq_length"]:
        logging.warning("File too long")
        return
    pad = np.zeros((hyper_params["max_input_seq_length"] - len(feat_vec), hyper_params["input_dim"]), dtype=np.float32)
    feat_vec = np.concatenate((feat_vec, pad), 0)
    model = rs.AcousticModel(hyper_params["num_layers"], hyper_params["hidden_size"], 1,
                             hyper_params["max_input_seq_length"], hyper_params["max_target_seq_length"],
                             hyper_params["input_dim"], hyper_params["batch_normalization"],
                             hyper_params["char_map_length"])
    model.create_forward_rnn()
    model.initialize(None)
    model.restore(None, os.path.join(hyper_params["checkpoint_dir"], "acoustic"))
    a, b = feat_vec.shape
    predictions = model.process_input(None, feat_vec.reshape((a, 1, b)), [original_len])
    print(rs.get_labels_str(hyper_params["char_map"], predictions[0]))


def evaluate(rs, hyper_params, test_set=None):
    """stt.py:294-324 of the reference."""
    if test_set is None:
        if hyper_params["test_dataset_dirs"] is None:
            logging.fatal("Setting test_dataset_dirs in config file is mandatory for evaluation mode")
            return
        test_set = load_dataset_dirs(hyper_params["test_dataset_dirs"])
    logging.info("Using %d size of test set", len(test_set))
    if len(test_set) == 0:
        logging.fatal("No files in test set during an evaluation mode")
        return
    model = rs.AcousticModel(hyper_params["num_layers"], hyper_params["hidden_size"], hyper_params["batch_size"],
                             hyper_params["max_input_seq_length"], hyper_params["max_target_seq_length"],
                             hyper_params["input_dim"], hyper_params["batch_normalization"],
                             hyper_params["char_map_length"])
    model.create_forward_rnn()
    model.initialize(None)
    model.restore(None, os.path.join(hyper_params["checkpoint_dir"], "acoustic"))
    wer, cer = model.evaluate_full(None, test_set, hyper_params["max_input_seq_length"],
                                   hyper_params["signal_processing"], hyper_params["char_map"])
    print("Resulting WER : {0:.3g} %".format(wer))
    print("Resulting CER : {0:.3g} %".format(cer))


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config', type=str, default='config.ini',
                        help='Path to configuration file with hyper-parameters.')
    parser.add_argument('--tb_name', type=str, default=None, help='Run name for the log / timeline output')
    parser.add_argument('--max_epoch', type=int, default=None, help='Max epoch to train (no limitation if not provided)')
    parser.add_argument('--learn_rate', type=float, default=None,
                        help='Force learning rate to start from this value (overriding checkpoint value)')
    parser.add_argument('--timeline', dest='timeline', action='store_true',
                        help='Write per-step timing JSON (a tensorboard directory must be provided in config file)')
    parser.add_argument('--XLA', dest='XLA', action='store_true', help='Accepted for compatibility; no effect')
    parser.add_argument('--synthetic', type=int, default=0, help='Use N seeded synthetic utterances as the dataset')
    group = parser.add_mutually_exclusive_group(required=True)
    group.add_argument('--train_acoustic', '--train', dest='train_acoustic', action='store_true',
                       help='Train the acoustic network')
    group.add_argument('--train_language', dest='train_language', action='store_true',
                       help='(reference stub; not part of the accelerated path)')
    group.add_argument('--file', type=str, help='Path to a wav file to process')
    group.add_argument('--record', dest='record', action='store_true', help='(needs audio hardware; not supported)')
    group.add_argument('--evaluate', dest='evaluate', action='store_true', help='Evaluate WER against the test_set')
    group.add_argument('--generate_text', dest='generate_text', action='store_true',
                       help='(reference stub; not part of the accelerated path)')
    args = parser.parse_args(argv)
    return {'config_file': args.config, 'tb_name': args.tb_name, 'max_epoch': args.max_epoch,
            'learn_rate': args.learn_rate, 'timeline': args.timeline, 'train_acoustic': args.train_acoustic,
            'train_language': args.train_language, 'file': args.file, 'record': args.record,
            'evaluate': args.evaluate, 'generate_text': args.generate_text, 'XLA': args.XLA,
            'synthetic': args.synthetic}


def main(argv=None):
    prog_params = parse_args(argv)
    if prog_params['train_language'] or prog_params['generate_text'] or prog_params['record']:
        print("This mode is outside the B200 acoustic-model path (see DESIGN.md, 'Out of scope').")
        return 2
    import rnn_speech_b200 as rs
    hyper_params = rs.HyperParameterHandler(prog_params['config_file']).get_hyper_params()
    audio_processor = rs.AudioProcessor(hyper_params["max_input_seq_length"], hyper_params["signal_processing"])
    hyper_params["input_dim"] = audio_processor.feature_size
    if hyper_params["language"] != "english":
        raise ValueError("Invalid parameter 'language'")
    hyper_params["char_map"] = rs.ENGLISH_CHAR_MAP
    hyper_params["char_map_length"] = len(rs.ENGLISH_CHAR_MAP)

    if prog_params['train_acoustic']:
        if prog_params['synthetic']:
            items = synthetic_dataset(prog_params['synthetic'], 2.0, 16000)
            split = max(1, int(0.9 * len(items)))
            train_set, test_set = items[:split], items[split:]
        else:
            ordered = hyper_params["dataset_size_ordering"] in ['True', 'First_run_only']       # stt.py:34-37
            train_set = load_dataset_dirs(hyper_params["training_dataset_dirs"], with_durations=ordered)
            test_set = load_dataset_dirs(hyper_params["test_dataset_dirs"]) if hyper_params["test_dataset_dirs"] else []
            train_set, test_set = split_acoustic_dataset(train_set, test_set, ordered, hyper_params.get("train_frac"))
        train_acoustic_rnn(rs, train_set, test_set, hyper_params, prog_params)
    elif prog_params['file'] is not None:
        process_file(rs, audio_processor, hyper_params, prog_params['file'])
    elif prog_params['evaluate']:
        evaluate(rs, hyper_params, synthetic_dataset(prog_params['synthetic'], 2.0, 16000, seed=1)
                 if prog_params['synthetic'] else None)
    return 0


if __name__ == "__main__":
    sys.exit(main())
