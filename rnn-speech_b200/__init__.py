"""rnn-speech_b200: B200-native acoustic-model path of domerin0/rnn-speech.

Import as ``rnn_speech_b200`` (the loader module at the repo root maps the
hyphenated directory name onto an importable package name).

Only the hot path lives here: the CUDA kernels + C ABI (csrc/, build.py ->
librnnspeech_b200.so) and the host-side mirror of the reference's
AudioProcessor / AcousticModel / label-codec interfaces.
"""
import os as _os

# The pipelined schedule spreads a step over ~10 CUDA streams (one per layer, the chunk GEMMs, the weight gradients, the
# bias sums, the input pipeline, the decoder, the read-back).  With the driver's default of 8 hardware queues some of them
# share a queue and wait for each other's heads: measured at cfg-2 (profiles/r02d_sweep17.log) 4 queues cost 2.1 ms per
# step, 16 or 32 are ~0.04 ms faster than 8.  Only effective before the CUDA context exists; an explicit setting wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib                      # noqa: E402,F401  (fails loudly if the .so is missing)
from ._lib import RnnSpeechError, LIB_PATH      # noqa: E402,F401
from .labels import ENGLISH_CHAR_MAP, get_labels_str, get_str_labels, get_str_to_one_hot_encoded  # noqa: E402,F401
from .audioprocessor import AudioProcessor, BatchPrefetcher      # noqa: E402,F401
from .acoustic_model import AcousticModel, OutOfRangeError, levenshtein   # noqa: E402,F401
from .hyperparams import HyperParameterHandler  # noqa: E402,F401

__version__ = "0.1.0"
