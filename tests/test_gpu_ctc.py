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
