"""GPU parity: the LSTM stack forward / backward (through the C ABI) against
oracle/model.py (float64) and the golden vectors of BASELINE config 1.

Tolerances (kernels accumulate in fp32): logits 2e-4 absolute, final state 1e-4,
CTC loss 1e-4 relative (north-star gate 1e-3), gradients 1e-3 of the largest
gradient entry.  Greedy label ids must be identical wherever the oracle's top-2
logit margin exceeds MARGIN (frames closer than that are ties at fp32 resolution
and are counted and reported, not compared).
"""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import ctc, model

pytestmark = pytest.mark.gpu
MARGIN = 1e-3


def _build(pkg, cuda, L, H, F, C, B, T, flat, training, ki=1.0, ko=1.0):
    m = pkg.AcousticModel(L, H, B, T, 600, F, False, C, device=cuda)
    if training:
        m.create_training_rnn(ki, ko, 1, 3e-4, 0.33)
    else:
        m.create_forward_rnn()
    m.load_flat_params(flat)
    return m


def _dev(a, cuda, dtype):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(cuda)


def _assert_labels_match(logits_gpu, logits_oracle, lens):
    margin = ctc.top2_margin(logits_oracle, lens)
    safe = margin > MARGIN
    same = logits_gpu.argmax(-1) == logits_oracle.argmax(-1)
    valid = np.arange(logits_oracle.shape[0])[:, None] < np.asarray(lens)[None, :]
    assert np.all(same[safe & valid]), "argmax differs on a frame with margin > %g" % MARGIN
    n_tie = int((~safe & valid).sum())
    return n_tie


def _lev(a, b):
    """Plain host Levenshtein distance (test-side checker)."""
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def _assert_ids_match(ids, n, logits_oracle, lens, tag):
    """Greedy label ids against the oracle's, for EVERY utterance: identical when no frame of the utterance is a near
    tie (oracle top-2 margin <= MARGIN); otherwise one flipped frame can change the collapsed sequence by at most two
    edits (a b a -> a a a), so the Levenshtein distance is bounded by 2 x (near-tie frames of that utterance)."""
    ref = ctc.greedy_decode(logits_oracle, lens)
    margin = ctc.top2_margin(logits_oracle, lens)
    ties = (margin <= MARGIN).sum(0)
    exact = 0
    for b in range(len(ref)):
        got = [int(v) for v in ids[b, :int(n[b])]]
        d = _lev(got, [int(v) for v in ref[b]])
        assert d <= 2 * int(ties[b]), "%s: utterance %d differs by %d edits with %d near-tie frames" % (tag, b, d, ties[b])
        exact += d == 0
    print("%s: %d of %d utterances decode to exactly the oracle's ids (%d near-tie frames of %d, all within the bound)"
          % (tag, exact, len(ref), int(ties.sum()), int(np.sum(lens))))
    return exact


def test_cfg1_golden_forward_backward(pkg, cuda):
    g = golden("model_cfg1.npz")
    L, H, F, C, T, B = [int(v) for v in g["dims"]]
    m = _build(pkg, cuda, L, H, F, C, B, T, g["flat_params"], training=True)
    x, lens = _dev(g["x"], cuda, np.float32), _dev(g["lens"], cuda, np.int32)
    labs = [g["lab_%d" % i] for i in range(B)]
    logits = m.forward(x, lens, training=True)
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], atol=2e-4)
    np.testing.assert_allclose(m.rnn_state[0, 0].cpu().numpy(), g["state_c"], atol=1e-4)
    np.testing.assert_allclose(m.rnn_state[0, 1].cpu().numpy(), g["state_h"], atol=1e-4)
    loss, grad = m.ctc_loss(logits, labs, lens)
    np.testing.assert_allclose(loss.cpu().numpy(), g["loss"], rtol=1e-4)
    m.grads.zero_()
    m.backward(x, lens, grad)
    got = m.grads.cpu().numpy()
    scale = np.abs(g["flat_grads"]).max()
    assert np.abs(got - g["flat_grads"]).max() < 1e-3 * scale
    # greedy labels identical to the oracle's (bit-exact ids)
    ids, n = m.greedy_decode(logits, lens)
    want = ctc.greedy_decode(g["logits"], g["lens"])
    ties = _assert_labels_match(logits.cpu().numpy(), g["logits"], g["lens"])
    print("cfg-1 golden: %d near-tie frames of %d" % (ties, int(np.sum(g["lens"]))))
    _assert_ids_match(ids.cpu().numpy(), n.cpu().numpy(), g["logits"], g["lens"], "cfg-1 golden")


@pytest.mark.parametrize("L,H,F,C,B,T,ki,ko", [
    (2, 50, 20, 50, 3, 17, 1.0, 1.0),       # odd sizes (reference's own test model: L2 H50 C50)
    (2, 64, 120, 80, 5, 33, 0.8, 0.5),      # dropout on both sides of every cell
    (3, 128, 120, 80, 40, 12, 1.0, 1.0),    # batch > 32 in one launch (two batch chunks; tile size 64)
    (2, 256, 120, 80, 21, 45, 0.8, 0.5),    # two chains per CTA, the second with 5 of 16 rows; self-validating exchange
    (1, 512, 40, 30, 32, 40, 1.0, 0.7),     # the same kernel at another hidden size
])
def test_random_models_with_state_and_dropout(pkg, cuda, monkeypatch, L, H, F, C, B, T, ki, ko):
    import sys
    # (batches of 33..64 rows are one launch of the recurrent kernels when the batch-tile size allows it)
    monkeypatch.setattr(sys.modules[pkg.__name__ + ".acoustic_model"], "TC_MAX_BATCH", 64)
    rng = np.random.default_rng(L * 1000 + H)
    p = model.init_params(L, H, F, C, seed=1, dtype=np.float64)
    for k in p:
        if p[k].ndim == 1:
            p[k] = rng.standard_normal(p[k].shape) * 0.1
    flat = model.flatten(p, L, H, F, C)
    x = rng.standard_normal((T, B, F))
    lens = rng.integers(0, T + 1, size=B).astype(np.int32)
    lens[0] = T
    state = [(rng.standard_normal((B, H)) * .3, rng.standard_normal((B, H)) * .3) for _ in range(L)]
    m = _build(pkg, cuda, L, H, F, C, B, T, flat, training=True, ki=ki, ko=ko)
    st = np.stack([np.stack([c, h]) for c, h in state])                # [L,2,B,H]
    m.rnn_state.copy_(_dev(st, cuda, np.float32))
    xd, ld = _dev(x, cuda, np.float32), _dev(lens, cuda, np.int32)
    logits = m.forward(xd, ld, training=True)
    keep_in, keep_out, seed, _ = m._last_fwd
    want, new_state, cache = model.forward(p, x, lens, L, H, state=state, keep_in=keep_in, keep_out=keep_out,
                                           seed=seed)
    np.testing.assert_allclose(logits.cpu().numpy(), want, atol=3e-4)
    for l in range(L):
        np.testing.assert_allclose(m.rnn_state[l, 0].cpu().numpy(), new_state[l][0], atol=2e-4)
        np.testing.assert_allclose(m.rnn_state[l, 1].cpu().numpy(), new_state[l][1], atol=2e-4)
    dl = rng.standard_normal(want.shape) * (np.arange(T)[:, None, None] < lens[None, :, None])
    m.grads.zero_()
    m.backward(xd, ld, _dev(dl, cuda, np.float32))
    gw = model.flatten(model.backward(p, cache, dl, L, H), L, H, F, C)
    got = m.grads.cpu().numpy()
    assert np.abs(got - gw).max() < 1e-3 * np.abs(gw).max()
    # accumulate semantics: a second forward/backward doubles the buffer (same seed & state)
    m.rnn_state.copy_(_dev(st, cuda, np.float32))
    m._dropout_calls -= 1
    m.forward(xd, ld, training=True)
    m.backward(xd, ld, _dev(dl, cuda, np.float32))
    assert np.abs(m.grads.cpu().numpy() - 2 * gw).max() < 2e-3 * np.abs(gw).max()


def test_inference_path_equals_training_path_without_dropout(pkg, cuda):
    L, H, F, C, B, T = 2, 64, 20, 30, 4, 21
    rng = np.random.default_rng(2)
    flat = model.flatten(model.init_params(L, H, F, C, seed=4), L, H, F, C)
    x = _dev(rng.standard_normal((T, B, F)), cuda, np.float32)
    lens = _dev(np.array([21, 3, 0, 17]), cuda, np.int32)
    a = _build(pkg, cuda, L, H, F, C, B, T, flat, training=True)
    b = _build(pkg, cuda, L, H, F, C, B, T, flat, training=False)
    la = a.forward(x, lens, training=True)
    lb = b.forward(x, lens, training=False)
    assert torch.equal(la, lb)
    # padded frames give the output bias; zero-length rows keep their state
    bias = a.param_views()["Output_layer/output_b"]
    assert torch.equal(la[5, 1], bias) and torch.equal(la[0, 2], bias)
    assert float(b.rnn_state[:, :, 2].abs().max()) == 0.0
    # process_input does not carry the state over (models/AcousticModel.py:716)
    before = b.rnn_state.clone()
    pred = b.process_input(None, x.cpu().numpy(), lens.cpu().numpy())
    assert torch.equal(before, b.rnn_state) and pred.dtype == np.int32 and pred.shape[0] == B


def test_cfg2_shape_forward_against_oracle(pkg, cuda):
    """BASELINE config 2 (3x768, B=32, T=998, F=120): logits, CTC loss (1e-3 gate)
    and greedy labels against the float64 oracle on the same synthetic input."""
    L, H, F, C, B, T = 3, 768, 120, 80, 32, 998
    rng = np.random.default_rng(0)
    p = model.init_params(L, H, F, C, seed=0)
    flat = model.flatten(p, L, H, F, C)
    x = rng.standard_normal((T, B, F)).astype(np.float32)
    lens = np.full(B, T, np.int32)
    lens[1::4] = rng.integers(T // 2, T, size=len(lens[1::4]))
    labs = [np.append(rng.integers(1, 79, size=rng.integers(60, 121)), 79).astype(np.int32) for _ in range(B)]
    m = _build(pkg, cuda, L, H, F, C, B, 1000, flat, training=False)
    logits = m.forward(_dev(x, cuda, np.float32), _dev(lens, cuda, np.int32), training=False)
    want, _, _ = model.forward(p, x, lens, L, H, keep_cache=False)
    got = logits.cpu().numpy()
    err = np.abs(got - want).max()
    loss, _ = m.ctc_loss(logits, labs, _dev(lens, cuda, np.int32), want_grad=False)
    want_loss, _ = ctc.ctc_loss_and_grad(want, labs, lens, want_grad=False)
    rel = np.abs(loss.cpu().numpy() - want_loss) / np.abs(want_loss)
    ties = _assert_labels_match(got, want, lens)
    print("cfg-2 forward: max |logit err| %.2e, max rel CTC loss err %.2e, %d near-tie frames of %d"
          % (err, rel.max(), ties, int(lens.sum())))
    assert err < 1e-3
    assert rel.max() < 1e-3
    # near ties (oracle top-2 margin <= MARGIN) are counted and bounded, not waved through: Xavier weights give flat
    # posteriors, a tenth of the frames at most may be that close; and every utterance WITHOUT a near-tie frame must
    # decode to exactly the oracle's label ids
    assert ties <= 0.1 * int(lens.sum())
    ids, n = m.greedy_decode(logits, _dev(lens, cuda, np.int32))
    _assert_ids_match(ids.cpu().numpy(), n.cpu().numpy(), want, lens, "cfg-2 forward")
    # the decode kernel itself is exact: the host decoder on the GPU's own logits gives the same ids
    own = ctc.greedy_decode(logits.cpu().numpy(), lens)
    for b in range(B):
        np.testing.assert_array_equal(ids[b, :int(n[b])].cpu().numpy(), own[b])


def cfg2_backward_parity(pkg, cuda, keep_in, keep_out, tag):
    """One forward + backward at BASELINE config 2's shape (3x768, B=32, T=998, ragged lengths) through the C ABI,
    dlogits from the ORACLE's CTC on the oracle's logits, every gradient tensor against oracle.model.backward
    (float64; models/AcousticModel.py:386-401).  Returns the per-tensor report."""
    from parity_util import format_report, grad_report, keep_artifact, tie_report
    L, H, F, C, B, T = 3, 768, 120, 80, 32, 998
    rng = np.random.default_rng(7)
    p = model.init_params(L, H, F, C, seed=0, dtype=np.float64)
    flat = model.flatten(p, L, H, F, C)
    x = rng.standard_normal((T, B, F))
    lens = np.full(B, T, np.int32)
    lens[1::4] = rng.integers(T // 2, T, size=len(lens[1::4]))
    lens[2] = 301
    labs = [np.append(rng.integers(1, 79, size=rng.integers(60, 121)), 79).astype(np.int32) for _ in range(B)]
    m = _build(pkg, cuda, L, H, F, C, B, 1000, flat, training=True, ki=keep_in, ko=keep_out)
    assert m.uses_tensor_cores
    xd, ld = _dev(x, cuda, np.float32), _dev(lens, cuda, np.int32)
    logits = m.forward(xd, ld, training=True)
    ki, ko, seed, _ = m._last_fwd
    want, _, cache = model.forward(p, x.astype(np.float32), lens, L, H, keep_in=ki, keep_out=ko, seed=seed)
    got = logits.cpu().numpy()
    frames, ties, mism = tie_report(got, want, lens, MARGIN)
    want_loss, dl = ctc.ctc_loss_and_grad(want, labs, lens)
    m.grads.zero_()
    m.backward(xd, ld, _dev(dl, cuda, np.float32))
    gw = model.flatten(model.backward(p, cache, dl, L, H), L, H, F, C)
    rows = grad_report(m.grads.cpu().numpy(), gw, L, H, F, C)
    title = ("cfg-2 shape backward vs float64 oracle [%s, keep %.1f/%.1f]: max |logit err| %.2e, %d frames, %d near ties, "
             "%d argmax mismatches" % (tag, keep_in, keep_out, np.abs(got - want).max(), frames, ties, mism))
    print(format_report(rows, title))
    keep_artifact("r02_grad_parity_%s.json" % tag, {"title": title, "rows": rows})
    return rows, np.abs(got - want).max(), mism


@pytest.mark.parametrize("keep_in,keep_out,tag", [(1.0, 1.0, "nodrop"), (0.8, 0.5, "dropout")])
def test_cfg2_shape_backward_against_oracle(pkg, cuda, keep_in, keep_out, tag):
    """VERDICT r01 item 1 / SURVEY a15: the gradients the headline number leans on, at the benchmarked shape.
    Gate: relative L2 error <= 1e-3 and cosine >= 1 - 1e-6 for EVERY parameter tensor."""
    rows, err, mism = cfg2_backward_parity(pkg, cuda, keep_in, keep_out, tag)
    assert err < 1e-3 and mism == 0
    for r in rows:
        assert r["rel_l2"] <= 1e-3, "%s: relative L2 error %.3e" % (r["tensor"], r["rel_l2"])
        assert r["cosine"] >= 1.0 - 1e-6, "%s: cosine %.9f" % (r["tensor"], r["cosine"])


@pytest.mark.parametrize("L,H,B,T,chunk", [(3, 256, 8, 150, 32), (2, 128, 16, 97, 20)])
def test_time_chunked_wavefront_is_bitwise_the_single_launch_schedule(pkg, cuda, monkeypatch, L, H, B, T, chunk):
    """The pipelined schedule (time chunks, layers as a wavefront on several streams) only reorders launches:
    logits and carried state must be bit-identical to one launch per layer (RS_TC_CHUNK=0), with dropout, ragged
    lengths and a carried-in state.  Weight gradients are summed chunk after chunk instead of in one GEMM, which
    regroups the fp32 additions: equal to 1e-5 of the largest entry."""
    F, C = 40, 30
    rng = np.random.default_rng(5)
    flat = model.flatten(model.init_params(L, H, F, C, seed=6), L, H, F, C)
    x = _dev(rng.standard_normal((T, B, F)), cuda, np.float32)
    lens_np = rng.integers(T // 3, T + 1, size=B).astype(np.int32)
    lens_np[0] = T
    lens = _dev(lens_np, cuda, np.int32)
    st0 = _dev(0.1 * rng.standard_normal((L, 2, B, H)), cuda, np.float32)
    dl = _dev(rng.standard_normal((T, B, C)) * (np.arange(T)[:, None, None] < lens_np[None, :, None]), cuda, np.float32)
    out = []
    for ch in (0, chunk):
        monkeypatch.setenv("RS_TC_CHUNK", str(ch))
        m = _build(pkg, cuda, L, H, F, C, B, T, flat, training=True, ki=0.8, ko=0.5)
        assert m.uses_tensor_cores
        m.rnn_state.copy_(st0)
        logits = m.forward(x, lens, training=True)
        m.grads.zero_()
        m.backward(x, lens, dl)
        torch.cuda.synchronize()
        out.append((logits.clone(), m.rnn_state.clone(), m.grads.clone()))
    assert torch.equal(out[0][0], out[1][0]), "logits differ"
    assert torch.equal(out[0][1], out[1][1]), "carried state differs"
    gdiff = float((out[0][2] - out[1][2]).abs().max() / out[0][2].abs().max())
    assert gdiff < 1e-5, "gradients differ by %g of the largest entry" % gdiff
    assert bool(torch.isfinite(out[1][2]).all())


def test_cfg4_shape_wavefront_against_oracle(pkg, cuda, monkeypatch):
    """BASELINE config 4's model (5x1024 LSTM, 120-dim input, per-GPU batch 16, ragged utterance lengths as a
    duration-bucketed batch gives) at a frame count the float64 oracle finishes in seconds, through the pipelined
    schedule (time chunks of 16 steps: 64 CTAs per layer, two layers in flight, one K-block of the forward weights in
    shared memory because 1024 columns exceed tensor memory): logits, CTC loss, greedy labels and gradients."""
    L, H, F, C, B, T = 5, 1024, 120, 80, 16, 56
    monkeypatch.setenv("RS_TC_CHUNK", "16")
    rng = np.random.default_rng(44)
    p = model.init_params(L, H, F, C, seed=3, dtype=np.float64)
    flat = model.flatten(p, L, H, F, C)
    x = rng.standard_normal((T, B, F))
    lens = np.sort(rng.integers(T // 2, T + 1, size=B))[::-1].astype(np.int32)      # sorted by duration, longest first
    lens[0] = T
    labs = [np.append(rng.integers(1, 79, size=rng.integers(3, 9)), 79).astype(np.int32) for _ in range(B)]
    m = _build(pkg, cuda, L, H, F, C, B, T, flat, training=True)
    assert m.uses_tensor_cores
    xd, ld = _dev(x, cuda, np.float32), _dev(lens, cuda, np.int32)
    logits = m.forward(xd, ld, training=True)
    want, _, cache = model.forward(p, x, lens, L, H)
    got = logits.cpu().numpy()
    assert np.abs(got - want).max() < 3e-4
    _assert_labels_match(got, want, lens)
    loss, grad = m.ctc_loss(logits, labs, ld)
    want_loss, dl = ctc.ctc_loss_and_grad(want, labs, lens)
    assert (np.abs(loss.cpu().numpy() - want_loss) / np.abs(want_loss)).max() < 1e-4
    m.grads.zero_()
    m.backward(xd, ld, _dev(dl, cuda, np.float32))
    gw = model.flatten(model.backward(p, cache, dl, L, H), L, H, F, C)
    err = np.abs(m.grads.cpu().numpy() - gw).max() / np.abs(gw).max()
    print("cfg-4 shape: gradient error %.2e of the largest entry" % err)
    assert err < 2e-3


def test_batch_tiles_training_against_oracle(pkg, cuda):
    """A mini-batch above the batch-tile size (32 utterances: 32 + 32 + 32 + 4 here) runs as batch tiles that share
    parameters, gradients and workspace: logits, carried state and the accumulated gradient against the oracle."""
    L, H, F, C, B, T = 2, 128, 40, 30, 100, 19
    rng = np.random.default_rng(9)
    p = model.init_params(L, H, F, C, seed=2, dtype=np.float64)
    flat = model.flatten(p, L, H, F, C)
    x = rng.standard_normal((T, B, F))
    lens = rng.integers(1, T + 1, size=B).astype(np.int32)
    lens[0] = T
    state = [(rng.standard_normal((B, H)) * .3, rng.standard_normal((B, H)) * .3) for _ in range(L)]
    m = _build(pkg, cuda, L, H, F, C, B, T, flat, training=True)
    assert m.uses_tensor_cores and m._tiles is not None and len(m._tiles) == 4
    m.rnn_state.copy_(_dev(np.stack([np.stack([c, h]) for c, h in state]), cuda, np.float32))
    xd, ld = _dev(x, cuda, np.float32), _dev(lens, cuda, np.int32)
    logits = m.forward(xd, ld, training=True)
    want, new_state, cache = model.forward(p, x, lens, L, H, state=state)
    np.testing.assert_allclose(logits.cpu().numpy(), want, atol=3e-4)
    for l in range(L):
        np.testing.assert_allclose(m.rnn_state[l, 0].cpu().numpy(), new_state[l][0], atol=2e-4)
        np.testing.assert_allclose(m.rnn_state[l, 1].cpu().numpy(), new_state[l][1], atol=2e-4)
    dl = rng.standard_normal(want.shape) * (np.arange(T)[:, None, None] < lens[None, :, None])
    m.grads.zero_()
    m.backward(xd, ld, _dev(dl, cuda, np.float32))
    gw = model.flatten(model.backward(p, cache, dl, L, H), L, H, F, C)
    assert np.abs(m.grads.cpu().numpy() - gw).max() < 2e-3 * np.abs(gw).max()


def test_cfg5_shape_inference_greedy_labels(pkg, cuda):
    """BASELINE config 5 (forward only, batch 256, 3x768) at a frame count the float64 oracle finishes in seconds:
    eight batch tiles through the tensor-core path; logits and greedy label ids (process_input) against the oracle."""
    L, H, F, C, B, T = 3, 768, 120, 80, 256, 20
    rng = np.random.default_rng(55)
    p = model.init_params(L, H, F, C, seed=8, dtype=np.float64)
    flat = model.flatten(p, L, H, F, C)
    x = rng.standard_normal((T, B, F)).astype(np.float32)
    lens = rng.integers(T // 2, T + 1, size=B).astype(np.int32)
    m = _build(pkg, cuda, L, H, F, C, B, T, flat, training=False)
    assert m.uses_tensor_cores and len(m._tiles) == 8
    logits = m.forward(_dev(x, cuda, np.float32), _dev(lens, cuda, np.int32), training=False, keep_state=False)
    want, _, _ = model.forward(p, x.astype(np.float64), lens, L, H, keep_cache=False)
    got = logits.cpu().numpy()
    assert np.abs(got - want).max() < 3e-4
    _assert_labels_match(got, want, lens)
    m.decoder = "greedy"
    pred = m.process_input(None, x, lens)
    ref = ctc.greedy_decode(want, lens)
    margin_ok = (ctc.top2_margin(want, lens) > MARGIN) | (np.arange(T)[:, None] >= lens[None, :])
    for b in range(B):
        if margin_ok[:, b].all():
            row = pred[b]
            np.testing.assert_array_equal(row[row != C], np.asarray(ref[b]))


def test_stale_tile_stress(pkg, cuda):
    """The recurrent kernels exchange h_t / dgates_t between CTAs through TMA stores, a counter and TMA loads.  A
    consumer that fetches a tile before the producer's store is visible reads what the PREVIOUS call left there, so
    two inputs alternate: every repetition must reproduce the first run of its input bit for bit (logits, carried
    state and gradients; cfg-2 shape, dropout on, the pipelined schedule)."""
    L, H, F, C, B, T = 3, 768, 120, 80, 32, 998
    rng = np.random.default_rng(21)
    flat = model.flatten(model.init_params(L, H, F, C, seed=0), L, H, F, C)
    x = torch.from_numpy(rng.standard_normal((T, B, F)).astype(np.float32)).to(cuda)
    lens_np = np.full(B, T, np.int32)
    lens_np[1::4] = rng.integers(T // 2, T, size=len(lens_np[1::4]))
    lens = torch.from_numpy(lens_np).to(cuda)
    dl = torch.from_numpy((rng.standard_normal((T, B, C)) * (np.arange(T)[:, None, None] < lens_np[None, :, None]))
                          .astype(np.float32)).to(cuda)
    xs = [x, torch.flip(x, dims=[0]) * 0.7]
    dls = [dl, torch.flip(dl, dims=[1]) * 1.3]
    m = _build(pkg, cuda, L, H, F, C, B, 1000, flat, training=True, ki=0.8, ko=0.5)
    refs = [None, None]
    for it in range(16):
        k = it & 1
        m.rnn_state.zero_()
        m._dropout_calls = 0
        logits = m.forward(xs[k], lens, training=True)
        m.grads.zero_()
        m.backward(xs[k], lens, dls[k])
        torch.cuda.synchronize()
        got = (logits.clone(), m.rnn_state.clone(), m.grads.clone())
        if refs[k] is None:
            refs[k] = got
            continue
        for a, b, what in zip(got, refs[k], ("logits", "state", "gradients")):
            assert torch.equal(a, b), "repetition %d: %s differ from the first run of the same input" % (it, what)


def test_exchange_fault_injection(pkg, cuda, monkeypatch):
    """The recurrent kernels publish h_t / dgates_t with plain stores and an unordered hint (a relaxed add on a counter);
    a consumer whose TMA fetch overtakes a store sees the fill pattern (a bf16 NaN) in its tile, the tensor core turns it
    into NaN accumulator columns, the epilogue votes, scans, and has the tile fetched and multiplied again.  Here the debug
    instantiation of both kernels sends the hint 3 us BEFORE half of every publish (RS_TS_FAULT=1): retries must happen
    (more MMA batches than steps) and logits, carried state and gradients must not change by a bit."""
    L, H, F, C, B, T = 1, 256, 40, 30, 27, 60
    rng = np.random.default_rng(3)
    flat = model.flatten(model.init_params(L, H, F, C, seed=4), L, H, F, C)
    x = _dev(rng.standard_normal((T, B, F)), cuda, np.float32)
    lens_np = rng.integers(T // 2, T + 1, size=B).astype(np.int32)
    lens_np[0] = T
    lens = _dev(lens_np, cuda, np.int32)
    dl = _dev(rng.standard_normal((T, B, C)) * (np.arange(T)[:, None, None] < lens_np[None, :, None]), cuda, np.float32)
    monkeypatch.setenv("RS_TC_CHUNK", "0")              # one launch per layer: the debug instantiations are single-launch
    m = _build(pkg, cuda, L, H, F, C, B, T, flat, training=True, ki=0.8, ko=0.5)
    assert m.uses_tensor_cores

    def run():
        m.rnn_state.zero_()
        m._dropout_calls = 0
        logits = m.forward(x, lens, training=True)
        m.grads.zero_()
        m.backward(x, lens, dl)
        torch.cuda.synchronize()
        return logits.clone(), m.rnn_state.clone(), m.grads.clone()

    ref = run()
    dbg_f = torch.zeros((2 * T, 16), dtype=torch.int64, device=cuda)
    dbg_b = torch.zeros((2 * T, 16), dtype=torch.int64, device=cuda)
    pkg._lib.call("rs_am_set_debug_timeline", m._handle, dbg_f.data_ptr(), dbg_b.data_ptr())
    try:
        clean = run()
        for a, b, what in zip(clean, ref, ("logits", "state", "gradients")):
            assert torch.equal(a, b), "debug instantiation: %s differ" % what
        batches_f, batches_b = int(dbg_f[T - 1, 7]), int(dbg_b[0, 2])
        assert batches_f == T and batches_b == T - 1, "no retries expected without the fault: %d, %d" % (batches_f, batches_b)
        monkeypatch.setenv("RS_TS_FAULT", "1")
        dbg_f.zero_(); dbg_b.zero_()
        got = run()
        batches_f, batches_b = int(dbg_f[T - 1, 7]), int(dbg_b[0, 2])
        print("fault injection: %d forward MMA batches for %d steps (chain 0 of CTA 0), %d backward for %d" % (batches_f, T, batches_b, T - 1))
        assert batches_f > T and batches_b > T - 1, "the injected fault caused no retries: the test tests nothing"
        for a, b, what in zip(got, ref, ("logits", "state", "gradients")):
            assert torch.equal(a, b), "with the injected fault: %s differ" % what
    finally:
        pkg._lib.call("rs_am_set_debug_timeline", m._handle, 0, 0)


@pytest.mark.parametrize("L,H,F,C,B,T,ki", [
    (2, 128, 40, 30, 12, 23, 0.8),     # tensor-core path, input dropout after the normalisation
    (1, 50, 20, 50, 5, 17, 1.0),       # FFMA path (H not a multiple of 64)
])
def test_batch_normalization_against_oracle(pkg, cuda, L, H, F, C, B, T, ki):
    """normalization=True (models/AcousticModel.py:253-259, config.ini batch_normalization): moments over the batch
    axis per (t, h), eps 1e-3, no scale / offset, padded frames included -- logits and every gradient against
    oracle/model.py."""
    rng = np.random.default_rng(H)
    p = model.init_params(L, H, F, C, seed=5, dtype=np.float64)
    p["input_b"] = rng.standard_normal(H) * 0.2
    flat = model.flatten(p, L, H, F, C)
    x = rng.standard_normal((T, B, F)) * (1.0 + np.arange(B)[None, :, None] * 0.3)      # different scales per item
    lens = rng.integers(T // 2, T + 1, size=B).astype(np.int32)
    lens[0] = T
    x = x * (np.arange(T)[:, None, None] < lens[None, :, None])                          # padded frames are zeros
    m = pkg.AcousticModel(L, H, B, T, 600, F, True, C, device=cuda)
    m.create_training_rnn(ki, 1.0, 1, 3e-4, 0.33)
    m.load_flat_params(flat)
    xd, ld = _dev(x, cuda, np.float32), _dev(lens, cuda, np.int32)
    logits = m.forward(xd, ld, training=True)
    keep_in, keep_out, seed, _ = m._last_fwd
    want, _, cache = model.forward(p, x, lens, L, H, keep_in=keep_in, keep_out=keep_out, seed=seed, normalization=True)
    np.testing.assert_allclose(logits.cpu().numpy(), want, atol=5e-4)
    dl = rng.standard_normal(want.shape) * (np.arange(T)[:, None, None] < lens[None, :, None])
    m.grads.zero_()
    m.backward(xd, ld, _dev(dl, cuda, np.float32))
    gw = model.flatten(model.backward(p, cache, dl, L, H), L, H, F, C)
    got = m.grads.cpu().numpy()
    assert np.abs(got - gw).max() < 2e-3 * np.abs(gw).max()
    # the input-dense gradients are the ones that pass through the normalisation
    n_in = F * H + H
    assert np.abs(got[:n_in] - gw[:n_in]).max() < 2e-3 * np.abs(gw[:n_in]).max()


def test_infer_signals_matches_oracle_pipeline(pkg, cuda):
    """The batched inference path of bench.py --config cfg5 / stt.py --evaluate (AcousticModel.infer_signals: host PCM
    -> one staged copy -> feature kernels -> forward per batch tile -> greedy decode -> ids on the host), 70 clips =
    tiles of 32 + 32 + 6, against the oracle's features + forward + greedy decode."""
    from oracle import features
    L, H, F, C, B, sr = 2, 128, 120, 80, 70, 16000
    rng = np.random.default_rng(70)
    p = model.init_params(L, H, F, C, seed=11, dtype=np.float64)
    flat = model.flatten(p, L, H, F, C)
    sigs = [(0.1 * rng.standard_normal(int(sr * s))).astype(np.float32) for s in rng.uniform(0.4, 1.0, size=B)]
    Tmax = 100
    ap = pkg.AudioProcessor(Tmax, "fbank", device=cuda)
    m = _build(pkg, cuda, L, H, F, C, B, Tmax, flat, training=False)
    assert m.uses_tensor_cores and len(m._tiles) == 3
    ids, n = m.infer_signals(ap, sigs, sr)
    feats = [features.fbank(s, sr, Tmax) for s in sigs]
    lens = np.array([min(k, Tmax) for _, k in feats])
    T = int(max(ap.num_frames(len(s), sr) for s in sigs))
    assert ids.shape == (B, min(T, Tmax)) and n.shape == (B,)
    x = np.zeros((ids.shape[1], B, F))
    for b, (f, _) in enumerate(feats):
        x[:len(f), b] = f
    want, _, _ = model.forward(p, x, lens, L, H, keep_cache=False)
    for b in range(B):
        assert np.all(ids[b, n[b]:] == -1)
    exact = _assert_ids_match(ids, n, want, lens, "infer_signals")
    assert exact >= B // 4
    # the pipelined form (bench.py --config cfg5's throughput figure): the batch staged and copied by the prefetcher's
    # thread and stream, the rest as above -- the same ids, bit for bit
    pre = pkg.BatchPrefetcher(ap)
    t1 = pre.submit(sigs, sr, defer_features=True)
    t2 = pre.submit(sigs[::-1], sr, defer_features=True)
    ids1, n1 = m.infer_ticket(ap, t1, sr)
    ids2, n2 = m.infer_ticket(ap, t2, sr)
    pre.close()
    assert np.array_equal(ids1, ids) and np.array_equal(n1, n)
    assert np.array_equal(n2, n[::-1])
    for b in range(B):
        assert np.array_equal(ids2[b, :n2[b]], ids[B - 1 - b, :n[B - 1 - b]])
