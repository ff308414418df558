"""CPU: the beam-search restatement (oracle/ctc.py::beam_search_decode, TF's CTCBeamSearchDecoder) against
exhaustive enumeration on lattices small enough to enumerate, and its documented properties."""
import itertools

import numpy as np
import pytest

from oracle import ctc


def _exhaustive(logits):
    """log P(label sequence) summed over every alignment, by enumeration."""
    T, C = logits.shape
    lp = logits - np.log(np.exp(logits).sum(-1, keepdims=True))
    tot = {}
    for path in itertools.product(range(C), repeat=T):
        s = sum(lp[t, k] for t, k in enumerate(path))
        seq, prev = [], -1
        for k in path:
            if k != prev and k != C - 1:
                seq.append(k)
            prev = k
        tot[tuple(seq)] = np.logaddexp(tot.get(tuple(seq), -np.inf), s)
    return tot


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_full_width_beam_finds_the_most_probable_labelling(seed):
    rng = np.random.default_rng(seed)
    T, C = 6, 4
    logits = (rng.standard_normal((T, 1, C)) * 2).astype(np.float32)
    out, score = ctc.beam_search_decode(logits, [T], beam_width=10000, merge_repeated=False)
    tot = _exhaustive(logits[:, 0].astype(np.float64))
    best = max(tot, key=tot.get)
    assert tuple(out[0]) == best
    assert abs(float(score[0]) - tot[best]) < 1e-4


def test_merge_repeated_collapses_output_labels_and_width_one_is_a_path():
    # frames: a, blank, a  -> labelling "aa"; TF's merge_repeated=True reports "a" (LabelSeq quirk)
    C = 3
    logits = np.full((3, 1, C), -5.0, np.float32)
    logits[0, 0, 0] = logits[2, 0, 0] = 5.0
    logits[1, 0, C - 1] = 5.0
    merged, _ = ctc.beam_search_decode(logits, [3], beam_width=100, merge_repeated=True)
    plain, _ = ctc.beam_search_decode(logits, [3], beam_width=100, merge_repeated=False)
    assert list(plain[0]) == [0, 0] and list(merged[0]) == [0]
    # scores normalised or not: same path
    a, _ = ctc.beam_search_decode(logits, [3], normalize=False)
    assert list(a[0]) == [0]


def test_peaky_outputs_agree_with_greedy_and_empty_sequences_decode_to_nothing():
    rng = np.random.default_rng(5)
    T, B, C = 40, 4, 20
    logits = (rng.standard_normal((T, B, C)) * 8).astype(np.float32)     # one class dominates every frame
    lens = np.array([40, 25, 0, 1])
    beam, _ = ctc.beam_search_decode(logits, lens, beam_width=100, merge_repeated=False)
    greedy = ctc.greedy_decode(logits, lens)
    for b in range(B):
        assert list(beam[b]) == list(greedy[b])
    assert len(beam[2]) == 0
