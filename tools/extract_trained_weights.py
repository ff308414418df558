#!/usr/bin/env python
"""Extract the shipped 3x1024 acoustic model (SURVEY 8c "golden weights") into a local fixture.

    python tools/extract_trained_weights.py [/root/reference/trained_models/english/acoustic/acousticmodel.ckpt.meta]

The reference's checkpoint .data file is a git-LFS pointer, but acousticmodel.ckpt.meta embeds every variable as a
Const initial value (rnn-speech_b200/tf_checkpoint.py::read_meta_initial_values).  The 12 tensors (101.5 MB fp32) are
written to tests/golden/_local/trained_3x1024.npz -- git-ignored (too large for history, and a trained model is not
ours to commit), but it travels with the gpurun snapshot like the built .so, so that
tests/test_gpu_trained.py can run the realistic-weight parity check on the GPU box, where /root/reference does not
exist.  The test skips when the file is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEFAULT = "/root/reference/trained_models/english/acoustic/acousticmodel.ckpt.meta"
OUT = os.path.join(ROOT, "tests", "golden", "_local", "trained_3x1024.npz")


def main():
    meta = sys.argv[1] if len(sys.argv) > 1 else DEFAULT
    import rnn_speech_b200  # noqa: F401
    from rnn_speech_b200.tf_checkpoint import read_meta_initial_values
    values = read_meta_initial_values(meta)
    keep = {k: np.asarray(v) for k, v in values.items()
            if k.split("/")[-1] in ("input_w", "input_b", "output_w", "output_b", "kernel", "bias", "global_step", "learning_rate")}
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez(OUT, **keep)
    for k in sorted(keep):
        print("%-60s %-16s %s" % (k, keep[k].shape, keep[k].dtype))
    print("wrote %s (%.1f MB)" % (OUT, os.path.getsize(OUT) / 1e6))


if __name__ == "__main__":
    main()
