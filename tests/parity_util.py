"""Shared helpers of the GPU parity tests (test infrastructure).

grad_report(): per-parameter-tensor comparison of a flat gradient buffer with the float64 oracle's:
max-norm error (of the tensor's largest entry), relative L2 error and cosine -- a max-norm bound alone says
little about the many small entries.
"""
import json
import os

import numpy as np

from oracle import model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def grad_report(got_flat, want_flat, L, H, F, C):
    got = model.unflatten(np.asarray(got_flat, dtype=np.float64), L, H, F, C)
    want = model.unflatten(np.asarray(want_flat, dtype=np.float64), L, H, F, C)
    rows = []
    for name, _ in model.param_shapes(L, H, F, C):
        g, w = got[name].ravel(), want[name].ravel()
        nw = float(np.linalg.norm(w))
        rows.append({
            "tensor": name,
            "max_err_of_max": float(np.abs(g - w).max() / max(np.abs(w).max(), 1e-300)),
            "rel_l2": float(np.linalg.norm(g - w) / max(nw, 1e-300)),
            "cosine": float(np.dot(g, w) / max(np.linalg.norm(g) * nw, 1e-300)),
            "norm": nw,
        })
    return rows


def format_report(rows, title):
    out = ["%s" % title, "  %-48s %12s %12s %14s %12s" % ("tensor", "max/max", "rel-L2", "1-cosine", "|g|")]
    for r in rows:
        out.append("  %-48s %12.3e %12.3e %14.3e %12.4e" % (r["tensor"], r["max_err_of_max"], r["rel_l2"], 1.0 - r["cosine"],
                                                           r["norm"]))
    return "\n".join(out)


def keep_artifact(name, obj):
    """Leave a JSON under gpurun_out/ (merged back from the GPU box) when the directory can be written."""
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "w") as fh:
            json.dump(obj, fh, indent=1)
    except OSError:
        pass


def tie_report(logits_gpu, logits_oracle, lens, margin):
    """(frames compared, near-tie frames, argmax mismatches on frames whose oracle top-2 margin exceeds `margin`)."""
    from oracle import ctc
    lens = np.asarray(lens)
    mg = ctc.top2_margin(logits_oracle, lens)
    valid = np.arange(logits_oracle.shape[0])[:, None] < lens[None, :]
    safe = (mg > margin) & valid
    same = logits_gpu.argmax(-1) == logits_oracle.argmax(-1)
    return int(valid.sum()), int((valid & ~safe).sum()), int((~same & safe).sum())
