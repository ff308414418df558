"""Oracle cross-checks for CTC (CPU): torch.nn.functional.ctc_loss is an
independent implementation, valid only when no label equals the blank id."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import ctc


def _torch_ctc(logits, labs, lens, blank):
    lt = torch.tensor(np.asarray(logits, np.float64), requires_grad=True)
    loss = torch.nn.functional.ctc_loss(
        torch.log_softmax(lt, -1), torch.tensor(np.concatenate(labs).astype(np.int64)),
        torch.tensor(np.asarray(lens, np.int64)), torch.tensor([len(l) for l in labs]),
        blank=blank, reduction="none")
    loss.sum().backward()
    return loss.detach().numpy(), lt.grad.numpy()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_loss_and_grad_match_torch_without_blank_labels(seed):
    rng = np.random.default_rng(seed)
    T, B, C = 25, 4, 12
    logits = rng.standard_normal((T, B, C)) * 2
    labs = [rng.integers(0, C - 1, size=n) for n in (3, 7, 1, 10)]
    labs[1][2] = labs[1][1]
    lens = np.array([25, 20, 2, 25])
    for mode in ("source", "dest"):
        loss, grad = ctc.ctc_loss_and_grad(logits, labs, lens, beta_skip=mode)
        tl, tg = _torch_ctc(logits, labs, lens, C - 1)
        np.testing.assert_allclose(loss, tl, rtol=1e-12, atol=1e-10)
        np.testing.assert_allclose(grad, tg, rtol=0, atol=1e-10)


def test_skipped_items_and_no_path():
    rng = np.random.default_rng(5)
    logits = rng.standard_normal((6, 3, 5))
    labs = [np.array([1, 2, 3, 1, 2, 3, 1]), np.array([1]), np.array([1, 1, 1, 1])]
    lens = np.array([6, 0, 6])
    loss, grad = ctc.ctc_loss_and_grad(logits, labs, lens)
    assert loss[0] == 0 and loss[1] == 0 and np.all(grad[:, :2] == 0)      # longer than input / empty
    assert np.isinf(loss[2])                                                # repeats need 7 frames
    np.testing.assert_allclose(grad[:, 2], np.exp(ctc.log_softmax(logits[:, 2])))


def test_eos_equal_blank_quirk_changes_loss():
    g = golden("ctc_eos.npz")
    assert not np.allclose(g["loss_source"][:3], g["loss_dest"][:3])
    g2 = golden("ctc_noeos.npz")
    np.testing.assert_allclose(g2["loss_source"], g2["loss_dest"], rtol=1e-12)
    np.testing.assert_allclose(g2["loss_source"][g2["torch_items"]], g2["torch_loss"], rtol=1e-10)


@pytest.mark.parametrize("tag", ["noeos", "eos"])
def test_oracle_reproduces_golden(tag):
    g = golden("ctc_%s.npz" % tag)
    labs = [g["lab_%d" % i] for i in range(int(g["B"]))]
    for mode in ("source", "dest"):
        loss, grad = ctc.ctc_loss_and_grad(g["logits"], labs, g["lens"], beta_skip=mode)
        np.testing.assert_allclose(loss, g["loss_" + mode], rtol=1e-12)
        np.testing.assert_allclose(grad, g["grad_" + mode], atol=1e-12)
    dec = ctc.greedy_decode(g["logits"], g["lens"])
    for i, d in enumerate(dec):
        np.testing.assert_array_equal(d, g["greedy_%d" % i])


def test_greedy_decode_rules():
    C = 4
    path = [3, 1, 1, 3, 1, 2, 2, 2, 3, 0]
    logits = np.full((len(path), 1, C), -1.0)
    for t, k in enumerate(path):
        logits[t, 0, k] = 1.0
    np.testing.assert_array_equal(ctc.greedy_decode(logits, [len(path)])[0], [1, 1, 2, 0])
    np.testing.assert_array_equal(ctc.greedy_decode(logits, [3])[0], [1])
    np.testing.assert_array_equal(ctc.greedy_decode(np.zeros((5, 1, C)), [5])[0], [0])   # ties -> first index


def test_sparse_from_dense_drops_zero_and_fills_empty():
    rows = ctc.sparse_from_dense([[5, 0, 7, 0], [0, 0, 0, 0]], 80, batch_size=3)
    np.testing.assert_array_equal(rows[0], [5, 7])
    np.testing.assert_array_equal(rows[1], [79])
    np.testing.assert_array_equal(rows[2], [79])
