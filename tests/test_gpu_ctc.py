"""GPU parity: CUDA CTC loss / gradient / greedy decode (through the C ABI)
against oracle/ctc.py and the committed golden vectors.

Tolerances (oracle is float64, kernels keep the lattice in fp32 but shifted so its
values stay O(|log y|), offsets in fp64): loss 1e-5 relative (the north-star gate is
1e-3), gradient 1e-4 absolute (entries are probabilities in [-1, 1]); greedy label
ids bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import ctc

pytestmark = pytest.mark.gpu


def _model(pkg, cuda, B, T, C=80):
    m = pkg.AcousticModel(1, 8, B, T, 600, 8, False, C, device=cuda)
    m.create_forward_rnn()
    return m


def _run(m, logits, labs, lens, mode="source", want_grad=True):
    m.beta_skip = 0 if mode == "source" else 1
    lg = torch.from_numpy(np.ascontiguousarray(logits, dtype=np.float32)).to(m.device)
    ln = torch.from_numpy(np.asarray(lens, np.int32)).to(m.device)
    loss, grad = m.ctc_loss(lg, labs, ln, want_grad=want_grad)
    torch.cuda.synchronize()
    return loss.cpu().numpy(), (grad.cpu().numpy() if want_grad else None)


@pytest.mark.parametrize("tag", ["noeos", "eos"])
@pytest.mark.parametrize("mode", ["source", "dest"])
def test_ctc_matches_golden(pkg, cuda, tag, mode):
    g = golden("ctc_%s.npz" % tag)
    B = int(g["B"])
    labs = [g["lab_%d" % i] for i in range(B)]
    T = g["logits"].shape[0]
    m = _model(pkg, cuda, B, T)
    loss, grad = _run(m, g["logits"], labs, g["lens"], mode)
    np.testing.assert_allclose(loss, g["loss_" + mode], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(grad, g["grad_" + mode], rtol=0, atol=1e-4)
    loss_only, _ = _run(m, g["logits"], labs, g["lens"], mode, want_grad=False)
    np.testing.assert_array_equal(loss_only, loss)
    if tag == "noeos":
        np.testing.assert_allclose(loss[g["torch_items"]], g["torch_loss"], rtol=1e-4)


def test_ctc_edge_cases(pkg, cuda):
    rng = np.random.default_rng(5)
    logits = rng.standard_normal((6, 4, 5)).astype(np.float32)
    labs = [np.array([1, 2, 3, 1, 2, 3, 1]), np.array([1]), np.array([1, 1, 1, 1]), np.array([2, 3])]
    lens = np.array([6, 0, 6, 1])
    m = _model(pkg, cuda, 4, 6, C=5)
    loss, grad = _run(m, logits, labs, lens)
    want_loss, want_grad = ctc.ctc_loss_and_grad(logits, labs, lens)
    assert loss[0] == 0 and loss[1] == 0 and loss[3] == 0               # skipped items
    assert np.isinf(loss[2]) and np.isinf(want_loss[2])                 # no valid path
    np.testing.assert_allclose(grad, want_grad, atol=1e-4)
    with pytest.raises(ValueError):
        _run(m, logits, [np.array([7])] * 4, lens)                      # label >= num_classes


@pytest.mark.parametrize("T,B,lab_lo,lab_hi", [(200, 8, 5, 40), (998, 32, 60, 120)])
def test_ctc_random_against_oracle(pkg, cuda, T, B, lab_lo, lab_hi):
    """(998, 32, 60..120 labels + EOS) is BASELINE config 2's CTC shape."""
    rng = np.random.default_rng(T)
    C = 80
    logits = (1.5 * rng.standard_normal((T, B, C))).astype(np.float32)
    labs = [np.append(rng.integers(1, 79, size=rng.integers(lab_lo, lab_hi + 1)), 79).astype(np.int32)
            for _ in range(B)]
    lens = rng.integers(T // 2, T + 1, size=B).astype(np.int32)
    lens[0] = T
    m = _model(pkg, cuda, B, T)
    loss, grad = _run(m, logits, labs, lens)
    want_loss, want_grad = ctc.ctc_loss_and_grad(logits, labs, lens)
    rel = np.abs(loss - want_loss) / np.abs(want_loss)
    print("ctc T=%d: max rel loss err %.2e, max |grad err| %.2e" % (T, rel.max(), np.abs(grad - want_grad).max()))
    assert rel.max() < 1e-5
    # fp32 lattice: per-step rounding accumulates ~sqrt(T); 1e-3 of a probability at T = 998
    np.testing.assert_allclose(grad, want_grad, rtol=0, atol=1e-4 if T <= 200 else 1e-3)
    for b in range(B):
        assert np.all(grad[lens[b]:, b] == 0)


@pytest.mark.parametrize("tag", ["noeos", "eos"])
def test_greedy_decode_matches_golden_exactly(pkg, cuda, tag):
    g = golden("ctc_%s.npz" % tag)
    B, T = int(g["B"]), g["logits"].shape[0]
    m = _model(pkg, cuda, B, T)
    lg = torch.from_numpy(g["logits"]).to(cuda)
    ids, n = m.greedy_decode(lg, torch.from_numpy(g["lens"]).to(cuda))
    ids, n = ids.cpu().numpy(), n.cpu().numpy()
    for b in range(B):
        assert n[b] == g["greedy_len"][b]
        np.testing.assert_array_equal(ids[b, :n[b]], g["greedy_%d" % b])
        assert np.all(ids[b, n[b]:] == -1)


def test_greedy_decode_ties_and_collapse(pkg, cuda):
    C = 4
    path = [3, 1, 1, 3, 1, 2, 2, 2, 3, 0]
    logits = np.full((len(path), 2, C), -1.0, np.float32)
    for t, k in enumerate(path):
        logits[t, 0, k] = 1.0
    logits[:, 1, :] = 0.0            # all ties -> class 0 every frame -> one symbol
    m = _model(pkg, cuda, 2, len(path), C=C)
    ids, n = m.greedy_decode(torch.from_numpy(logits).to(cuda), torch.tensor([10, 10], dtype=torch.int32, device=cuda))
    ids, n = ids.cpu().numpy(), n.cpu().numpy()
    np.testing.assert_array_equal(ids[0, :n[0]], [1, 1, 2, 0])
    np.testing.assert_array_equal(ids[1, :n[1]], [0])


def test_greedy_decode_full_size(pkg, cuda):
    rng = np.random.default_rng(9)
    T, B, C = 998, 32, 80
    logits = rng.standard_normal((T, B, C)).astype(np.float32)
    lens = rng.integers(1, T + 1, size=B).astype(np.int32)
    m = _model(pkg, cuda, B, T)
    ids, n = m.greedy_decode(torch.from_numpy(logits).to(cuda), torch.from_numpy(lens).to(cuda))
    want = ctc.greedy_decode(logits, lens)
    ids, n = ids.cpu().numpy(), n.cpu().numpy()
    for b in range(B):
        np.testing.assert_array_equal(ids[b, :n[b]], want[b])


@pytest.mark.parametrize("T,B,C,W,scale,merge", [
    (50, 6, 80, 100, 3.0, True),        # the reference's call: width 100, merge_repeated
    (24, 4, 80, 100, 0.05, True),       # near-uniform scores: thousands of extensions compete every frame
    (60, 5, 30, 16, 2.0, False),        # narrow beam: entries leave the beam and their prefixes come back
    (40, 3, 5, 100, 1.0, True),         # beam wider than the number of live prefixes
])
def test_beam_search_matches_the_tf_restatement(pkg, cuda, T, B, C, W, scale, merge):
    """rs_ctc_beam_search against oracle/ctc.py::beam_search_decode (TF's CTCBeamSearchDecoder, float32): same top
    path and score.  An item whose best two beam totals differ by less than 1e-4 is a tie at fp32 resolution and is
    only required to score the same."""
    rng = np.random.default_rng(T * 100 + C)
    logits = (rng.standard_normal((T, B, C)) * scale).astype(np.float32)
    lens = rng.integers(T // 2, T + 1, size=B).astype(np.int32)
    lens[0] = T
    if B > 3:
        lens[-1] = 0
    m = _model(pkg, cuda, B, T, C)
    ids, n, score = m.beam_search_decode(torch.from_numpy(logits).to(cuda), torch.from_numpy(lens).to(cuda),
                                         beam_width=W, merge_repeated=merge)
    ids, n, score = ids.cpu().numpy(), n.cpu().numpy(), score.cpu().numpy()
    want, want_score = ctc.beam_search_decode(logits, lens, beam_width=W, merge_repeated=merge)
    for b in range(B):
        assert abs(score[b] - want_score[b]) < 1e-3 * max(1.0, abs(want_score[b]))
        got = ids[b, :n[b]]
        if list(got) != list(want[b]):
            # accept only a genuine tie: the GPU's path must score like the oracle's best
            alt, alt_score = ctc.beam_search_decode(logits[:, b:b + 1], lens[b:b + 1], beam_width=W, merge_repeated=merge)
            assert abs(score[b] - alt_score[0]) < 1e-4, "item %d: different path, scores %g vs %g" % (b, score[b], alt_score[0])
        assert np.all(ids[b, n[b]:] == -1)


@pytest.mark.parametrize("W,C", [(8, 6), (100, 80)])
def test_beam_search_with_exactly_tied_scores(pkg, cuda, W, C):
    """Frames whose logits are small integers give MANY extensions exactly the same total: the selection has to cut the
    tie at the beam boundary by candidate id (the radix select's id digits), stay within the beam width, and still
    return a path whose score is the oracle's best (which labels win a tie is unspecified in TF)."""
    T, B = 30, 4
    rng = np.random.default_rng(W + C)
    logits = rng.integers(0, 3, size=(T, B, C)).astype(np.float32)
    logits[5:9] = 0.0                                                    # completely flat frames
    lens = np.array([T, T - 3, T, 11], np.int32)
    m = _model(pkg, cuda, B, T, C)
    runs = []
    for _ in range(2):
        ids, n, score = m.beam_search_decode(torch.from_numpy(logits).to(cuda), torch.from_numpy(lens).to(cuda),
                                             beam_width=W, merge_repeated=True)
        runs.append((ids.cpu().numpy(), n.cpu().numpy(), score.cpu().numpy()))
    np.testing.assert_array_equal(runs[0][0], runs[1][0])               # deterministic whatever the thread timing
    ids, n, score = runs[0]
    _, want_score = ctc.beam_search_decode(logits, lens, beam_width=W, merge_repeated=True)
    for b in range(B):
        assert np.isfinite(score[b]) and 0 <= n[b] <= lens[b]
        if W >= 100:
            # a wide beam holds every prefix that matters: the best total does not depend on how ties were cut
            assert abs(score[b] - want_score[b]) < 1e-3 * max(1.0, abs(want_score[b]))
        assert np.all(ids[b, n[b]:] == -1) and np.all((ids[b, :n[b]] >= 0) & (ids[b, :n[b]] < C - 1))


def test_beam_search_full_size_and_process_input(pkg, cuda):
    """cfg-2 sized logits (T=998, B=32, C=80): runs, agrees with greedy decoding on peaky outputs (where the best
    labelling is the best path), and one item is checked against the restatement on its first 150 frames."""
    rng = np.random.default_rng(3)
    T, B, C = 998, 32, 80
    logits = (rng.standard_normal((T, B, C)) * 8).astype(np.float32)
    lens = np.full(B, T, np.int32)
    lens[1::3] = rng.integers(T // 2, T, size=len(lens[1::3]))
    m = _model(pkg, cuda, B, T, C)
    lg, ld = torch.from_numpy(logits).to(cuda), torch.from_numpy(lens).to(cuda)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ids, n, score = m.beam_search_decode(lg, ld, merge_repeated=False)
    e1.record()
    torch.cuda.synchronize()
    print("beam search T=998 B=32 C=80 width 100: %.2f ms" % e0.elapsed_time(e1))
    gids, gn = m.greedy_decode(lg, ld)
    ids, n, gids, gn = ids.cpu().numpy(), n.cpu().numpy(), gids.cpu().numpy(), gn.cpu().numpy()
    same = sum(int(n[b] == gn[b] and np.array_equal(ids[b, :n[b]], gids[b, :gn[b]])) for b in range(B))
    print("beam == greedy on %d of %d items" % (same, B))
    assert same >= B // 2, "beam and greedy paths differ on %d of %d peaky items" % (B - same, B)
    short = np.minimum(lens, 150).astype(np.int32)
    ids2, n2, _ = m.beam_search_decode(lg, torch.from_numpy(short).to(cuda), merge_repeated=True)
    want, _ = ctc.beam_search_decode(logits[:150, :1], short[:1], merge_repeated=True)
    np.testing.assert_array_equal(ids2[0, :int(n2[0])].cpu().numpy(), want[0])


def test_edit_distance_matches_host_levenshtein(pkg, cuda):
    """rs_edit_distance (tf.edit_distance(normalize=True), models/AcousticModel.py:370) against the host DP on random
    pairs: empty sequences, long truths (600 labels, max_target_seq_length), near-identical and unrelated pairs."""
    rng = np.random.default_rng(17)
    B, ld = 40, 700
    truths, hyps = [], []
    for b in range(B):
        n = [0, 1, 600, 37][b] if b < 4 else int(rng.integers(0, 200))
        t = rng.integers(0, 79, size=n).astype(np.int32)
        if b % 3 == 0 and n > 0:                      # a noisy copy: substitutions, deletions, insertions
            keep = rng.random(n) > 0.1
            hseq = np.where(rng.random(n) < 0.1, rng.integers(0, 79, size=n), t)[keep]
            hseq = np.insert(hseq, rng.integers(0, len(hseq) + 1, size=3), rng.integers(0, 79, size=3))
        else:
            hseq = rng.integers(0, 79, size=int(rng.integers(0, 150)))
        if b == 1:
            hseq = np.zeros(0, np.int64)
        truths.append(t)
        hyps.append(hseq.astype(np.int32))
    ids = np.full((B, ld), -1, np.int32)
    for b, hseq in enumerate(hyps):
        ids[b, :len(hseq)] = hseq
    lens = np.array([len(hseq) for hseq in hyps], np.int32)
    m = _model(pkg, cuda, B, 50)
    dist, rate = m.edit_distance(torch.from_numpy(ids).to(cuda), torch.from_numpy(lens).to(cuda), truths)
    dist, rate = dist.cpu().numpy(), rate.cpu().numpy()
    for b in range(B):
        want = pkg.levenshtein(hyps[b], truths[b])
        assert dist[b] == want, "item %d: %d vs %d" % (b, dist[b], want)
        if len(truths[b]):
            assert abs(rate[b] - want / len(truths[b])) < 1e-6
        else:
            assert rate[b] == (np.inf if want else 0.0)
