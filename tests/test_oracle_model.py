"""Oracle cross-checks for the LSTM stack (CPU): the hand-written backward in
oracle/model.py against torch-CPU autograd of an independent restatement of the
TF cell; the update rule against a second formulation."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import ctc, model, optim


def _torch_forward(p, x, lens, L, H, state, ki, ko, seed):
    ki, ko = float(np.float32(ki)), float(np.float32(ko))      # float32 keep probabilities, as the oracle takes them
    T, B, F = x.shape
    tp = {k: torch.tensor(v, requires_grad=True) for k, v in p.items()}
    cur = (torch.tensor(x).reshape(T * B, F) @ tp["input_w"] + tp["input_b"]).reshape(T, B, H)
    valid = torch.tensor(np.arange(T)[:, None] < np.asarray(lens)[None, :])
    for l in range(L):
        K, b = tp["kernel_%d" % l], tp["bias_%d" % l]
        mi = model.dropout_mask(seed, 2 * l, T, B, H, ki)
        mo = model.dropout_mask(seed, 2 * l + 1, T, B, H, ko)
        xin = cur if mi is None else cur * torch.tensor(mi) / ki
        c, h = torch.tensor(state[l][0]), torch.tensor(state[l][1])
        outs = []
        for t in range(T):
            g = torch.cat([xin[t], h], 1) @ K + b
            i, j, f, o = g.split(H, 1)
            cn = c * torch.sigmoid(f + 1.0) + torch.sigmoid(i) * torch.tanh(j)
            hn = torch.tanh(cn) * torch.sigmoid(o)
            v = valid[t][:, None]
            outs.append(torch.where(v, hn, torch.zeros_like(hn)))
            c, h = torch.where(v, cn, c), torch.where(v, hn, h)
        out = torch.stack(outs)
        cur = out if mo is None else out * torch.tensor(mo) / ko
    C = tp["output_w"].shape[1]
    return (cur.reshape(T * B, H) @ tp["output_w"] + tp["output_b"]).reshape(T, B, C), tp, (c, h)


@pytest.mark.parametrize("ki,ko", [(1.0, 1.0), (0.8, 0.5)])
def test_forward_backward_match_autograd(ki, ko):
    L, H, F, C, T, B = 2, 16, 12, 10, 9, 3
    rng = np.random.default_rng(0)
    p = model.init_params(L, H, F, C, seed=3, dtype=np.float64)
    for k in p:
        if p[k].ndim == 1:
            p[k] = rng.standard_normal(p[k].shape) * 0.1
    x = rng.standard_normal((T, B, F))
    lens = np.array([9, 5, 0])
    state = [(rng.standard_normal((B, H)) * .3, rng.standard_normal((B, H)) * .3) for _ in range(L)]
    logits, new_state, cache = model.forward(p, x, lens, L, H, state=state, keep_in=ki, keep_out=ko, seed=7)
    dl = rng.standard_normal(logits.shape)
    g = model.backward(p, cache, dl, L, H)
    tl, tp, (c, h) = _torch_forward(p, x, lens, L, H, state, ki, ko, 7)
    np.testing.assert_allclose(logits, tl.detach().numpy(), atol=1e-12)
    np.testing.assert_allclose(new_state[-1][1], h.detach().numpy(), atol=1e-12)
    # frozen rows: item 2 has length 0 -> state unchanged, logits = output bias
    np.testing.assert_allclose(new_state[0][0][2], state[0][0][2])
    np.testing.assert_allclose(logits[:, 2], np.broadcast_to(p["output_b"], (T, C)), atol=1e-12)
    (tl * torch.tensor(dl)).sum().backward()
    for k in p:
        np.testing.assert_allclose(g[k], tp[k].grad.numpy(), atol=1e-11, err_msg=k)


def test_flat_layout_roundtrip_and_count():
    L, H, F, C = 3, 768, 120, 80
    assert model.param_count(L, H, F, C) == 14319440
    assert model.param_count(5, 1024, 120, 80) == 42169424
    p = model.init_params(1, 8, 4, 5, seed=1)
    flat = model.flatten(p, 1, 8, 4, 5)
    q = model.unflatten(flat, 1, 8, 4, 5)
    for k in p:
        np.testing.assert_array_equal(p[k], q[k])
    assert np.all(p["bias_0"] == 0) and abs(p["kernel_0"]).max() <= np.sqrt(6.0 / (16 + 32))


def test_model_golden_reproduced():
    g = golden("model_cfg1.npz")
    L, H, F, C, T, B = [int(v) for v in g["dims"]]
    p = model.unflatten(g["flat_params"], L, H, F, C)
    logits, state, cache = model.forward(p, g["x"], g["lens"], L, H)
    np.testing.assert_allclose(logits, g["logits"], atol=1e-12)
    labs = [g["lab_%d" % i] for i in range(B)]
    loss, dlogits = ctc.ctc_loss_and_grad(logits, labs, g["lens"])
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-12)
    grads = model.backward(p, cache, dlogits, L, H)
    np.testing.assert_allclose(model.flatten(grads, L, H, F, C), g["flat_grads"], atol=1e-10)


def test_dropout_mask_statistics_and_determinism():
    m = model.dropout_mask(42, 3, 50, 4, 64, 0.8)
    assert m.shape == (50, 4, 64) and abs(m.mean() - 0.8) < 0.02
    np.testing.assert_array_equal(m, model.dropout_mask(42, 3, 50, 4, 64, 0.8))
    assert (m != model.dropout_mask(42, 2, 50, 4, 64, 0.8)).any()
    assert model.dropout_mask(1, 0, 2, 2, 2, 1.0) is None


def test_clip_and_adam_against_torch():
    rng = np.random.default_rng(0)
    n = 1000
    # |g| bounded away from 0 so that the eps placement (TF: outside the corrected sqrt) is negligible
    theta, g = rng.standard_normal(n), rng.choice([-1.0, 1.0], n) * rng.uniform(20.0, 60.0, n)
    m, v = np.zeros(n), np.zeros(n)
    tt = torch.tensor(theta.copy(), requires_grad=True)
    opt = torch.optim.Adam([tt], lr=3e-4, betas=(0.9, 0.999), eps=1e-8)
    th = theta.copy()
    for step in range(1, 4):
        th, m, v, norm = optim.clip_adam_step(th, g, m, v, step, 3e-4, 1.0)
        assert abs(norm - np.linalg.norm(g)) < 1e-9
        tt.grad = torch.tensor(g * (1.0 / max(np.linalg.norm(g), 1.0)))
        opt.step()
    # torch puts eps inside the bias-corrected denominator, TF outside: equal to O(eps)
    np.testing.assert_allclose(th, tt.detach().numpy(), atol=5e-8)



@pytest.mark.parametrize("ki,ko", [(1.0, 1.0), (0.8, 0.5)])
def test_torch_cpu_leg_equals_numpy_oracle(ki, ko):
    """oracle/model_torch.py (the MKL leg of bench.py's CPU baseline) is the same arithmetic as oracle/model.py."""
    from oracle import model_torch
    L, H, F, C, T, B = 2, 16, 12, 10, 11, 4
    rng = np.random.default_rng(3)
    p = model.init_params(L, H, F, C, seed=7, dtype=np.float64)
    for k in p:
        if p[k].ndim == 1:
            p[k] = rng.standard_normal(p[k].shape) * 0.1
    x = rng.standard_normal((T, B, F))
    lens = np.array([11, 7, 0, 10])
    logits, _, cache = model.forward(p, x, lens, L, H, keep_in=ki, keep_out=ko, seed=9)
    tl, tcache = model_torch.forward(p, x, lens, L, H, keep_in=ki, keep_out=ko, seed=9, dtype=torch.float64)
    np.testing.assert_allclose(tl.numpy(), logits, rtol=1e-11, atol=1e-12)
    dl = rng.standard_normal(logits.shape) * (np.arange(T)[:, None, None] < lens[None, :, None])
    want = model.backward(p, cache, dl, L, H)
    got = model_torch.backward(tcache, dl, L, H)
    for k in want:
        np.testing.assert_allclose(got[k], want[k], rtol=1e-10, atol=1e-11)
