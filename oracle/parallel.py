"""Oracle (test infrastructure, NOT product code): per-utterance stages of the CPU baseline spread over host
processes -- the feature extraction (the reference runs it in tf.data's parallel map,
/root/reference/models/AcousticModel.py:819-822) and the CTC loss/gradient (TF's CTCLoss op shards the batch
items over its thread pool).  Used only by bench.py's CPU legs so that the baseline really uses every host core;
the arithmetic is oracle/features.py and oracle/ctc.py, unchanged.
"""
import multiprocessing as mp
import os

import numpy as np

_POOL = None
_POOL_N = 0


def _init_worker():
    # one BLAS thread per worker: the workers are the parallelism
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def pool(n=None):
    global _POOL, _POOL_N
    n = int(n or host_threads())
    if _POOL is None or _POOL_N != n:
        if _POOL is not None:
            _POOL.terminate()
        _POOL = mp.get_context("spawn").Pool(n, initializer=_init_worker)
        _POOL_N = n
    return _POOL


def close():
    global _POOL
    if _POOL is not None:
        _POOL.terminate()
        _POOL = None


def _fbank_one(args):
    from oracle import features
    sig, sr, tmax = args
    f, n = features.fbank(sig, sr, tmax)
    return f.astype(np.float32), int(n)


def _ctc_one(args):
    from oracle import ctc
    logits_b, lab, L, blank = args
    if L == 0 or len(lab) > L:
        return 0.0, None
    logp = ctc.log_softmax(np.asarray(logits_b[:L], dtype=np.float64))
    l, g, _, _ = ctc.ctc_item(logp, np.asarray(lab, dtype=np.int64), blank)
    return float(l), g.astype(np.float32)


def fbank_batch(sigs, sr, tmax, n=None):
    return pool(n).map(_fbank_one, [(s, sr, tmax) for s in sigs])


def ctc_batch(logits, labels_list, seq_len, n=None):
    """Same result as oracle.ctc.ctc_loss_and_grad (float32 gradient), items spread over the pool."""
    logits = np.asarray(logits)
    T, B, C = logits.shape
    out = pool(n).map(_ctc_one, [(np.ascontiguousarray(logits[:, b, :]), labels_list[b], int(seq_len[b]), C - 1)
                                 for b in range(B)])
    loss = np.zeros(B)
    grad = np.zeros((T, B, C), np.float32)
    for b, (l, g) in enumerate(out):
        loss[b] = l
        if g is not None:
            grad[:g.shape[0], b, :] = g
    return loss, grad
