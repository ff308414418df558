"""rnn-speech_b200: B200-native acoustic-model path of domerin0/rnn-speech.

Import as ``rnn_speech_b200`` (the loader module at the repo root maps the
hyphenated directory name onto an importable package name).

Only the hot path lives here: the CUDA kernels + C ABI (csrc/, build.py ->
librnnspeech_b200.so) and the host-side mirror of the reference's
AudioProcessor / AcousticModel / label-codec interfaces.
"""
from . import _lib                      # noqa: F401  (fails loudly if the .so is missing)
from ._lib import RnnSpeechError, LIB_PATH      # noqa: F401
from .labels import ENGLISH_CHAR_MAP, get_labels_str, get_str_labels, get_str_to_one_hot_encoded  # noqa: F401
from .audioprocessor import AudioProcessor, BatchPrefetcher      # noqa: F401
from .acoustic_model import AcousticModel, OutOfRangeError, levenshtein   # noqa: F401
from .hyperparams import HyperParameterHandler  # noqa: F401

__version__ = "0.1.0"
