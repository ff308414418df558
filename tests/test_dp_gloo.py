"""Data-parallel host logic on CPU (gloo, world_size 2): sharding utterances over ranks and
summing the flat gradient buffer reproduces the reference's gradient accumulation over
mini-batches (models/AcousticModel.py:386-401) -- checked with the oracle's gradients."""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

L, H, F, C, T, B = 1, 16, 12, 10, 14, 4


def _dist_module():
    import rnn_speech_b200
    from rnn_speech_b200 import dist as mod
    return mod


def _problem():
    from oracle import model
    rng = np.random.default_rng(0)
    p = model.init_params(L, H, F, C, seed=2, dtype=np.float64)
    x = rng.standard_normal((T, B, F))
    lens = np.array([14, 9, 14, 11])
    labs = [np.append(rng.integers(1, C - 1, size=3), C - 1) for _ in range(B)]
    return p, x, lens, labs


def _grads(p, x, lens, labs):
    from oracle import ctc, model
    logits, _, cache = model.forward(p, x, lens, L, H)
    loss, dlogits = ctc.ctc_loss_and_grad(logits, labs, lens)
    return model.flatten(model.backward(p, cache, dlogits, L, H), L, H, F, C), loss


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rsd = _dist_module()
    p, x, lens, labs = _problem()
    mine = rsd.shard(list(range(B)))                      # utterance ids of this rank
    assert rsd.world() == (rank, world) and mine == list(range(B))[rank::world]
    g, loss = _grads(p, x[:, mine], lens[mine], [labs[i] for i in mine])
    flat = torch.from_numpy(g.copy())
    rsd.allreduce_sum_(flat)
    acc = torch.tensor([float(np.mean(loss / lens[mine])), 0.0, 1.0], dtype=torch.float64)
    rsd.allreduce_sum_(acc)
    if rank == 0:
        np.save(out, np.concatenate([flat.numpy(), acc.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_minibatch_accumulation(tmp_path):
    out = str(tmp_path / "g.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    p, x, lens, labs = _problem()
    want = np.zeros_like(got[:-3])
    mean_losses = []
    for r in range(2):                                    # the reference: two accumulated mini-batches
        ids = list(range(B))[r::2]
        g, loss = _grads(p, x[:, ids], lens[ids], [labs[i] for i in ids])
        want += g
        mean_losses.append(np.mean(loss / lens[ids]))
    np.testing.assert_allclose(got[:-3], want, rtol=1e-12, atol=1e-12)
    # and it equals the gradient of the summed loss over the whole batch of 4
    g_all, _ = _grads(p, x, lens, labs)
    np.testing.assert_allclose(got[:-3], g_all, rtol=1e-9, atol=1e-10)
    assert abs(got[-3] / got[-1] - np.mean(mean_losses)) < 1e-12 and got[-1] == 2.0


def test_single_process_is_a_no_op():
    rsd = _dist_module()
    t = torch.arange(4.0)
    assert rsd.world() == (0, 1) and torch.equal(rsd.allreduce_sum_(t.clone()), t)
    assert rsd.shard([1, 2, 3]) == [1, 2, 3]
