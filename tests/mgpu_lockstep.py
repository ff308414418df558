"""Multi-GPU check (run under torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_lockstep.py

1. the library's own NCCL communicator (rs_comm_init / rs_allreduce_sum) sums a buffer over the ranks;
2. data-parallel training stays in lockstep when the ranks' shards hold DIFFERENT numbers of mini-batches (ADVICE r01:
   a rank whose dataset runs dry mid-step must still join the collectives): every rank ends the epoch in the same
   call, with the same global_step and bit-identical parameters;
3. N ranks x 1 mini-batch equal one process accumulating N mini-batches (the reference's mini_batch_size = N,
   models/AcousticModel.py:386-401) -- the summed gradients are compared.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rnn_speech_b200 as rs          # noqa: E402
from rnn_speech_b200 import dist as rsdist        # noqa: E402


def synth(n_items, seed, seconds=1.0, sr=16000):
    rng = np.random.default_rng(seed)
    items = []
    for _ in range(n_items):
        sig = (0.1 * rng.standard_normal(int(seconds * sr))).astype(np.float32)
        items.append([(sig, sr), "the quick fox"])
    return items


def build(dev, items, B):
    m = rs.AcousticModel(2, 128, B, 100, 50, 120, False, 80, device=dev, seed=0)
    ds = m.build_dataset(items, B, 100, 50, "fbank", rs.ENGLISH_CHAR_MAP, device=dev)
    m.add_dataset_input(ds)
    m.create_training_rnn(1.0, 1.0, 1, 1e-3, 0.33, use_iterator=True)
    m.initialize(None)
    return m


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # 1. in-library all-reduce
    t = torch.full((1000003,), float(rank + 1), dtype=torch.float32, device=dev)
    rsdist.allreduce_sum_(t)
    torch.cuda.synchronize()
    want = world * (world + 1) / 2.0
    assert float(t.min()) == want and float(t.max()) == want, (float(t.min()), want)
    assert rsdist._COMM["handle"] is not None, "the library's communicator was not used"
    # 2. unequal shards: rank r owns 2 + (r == 0) mini-batches of 4 utterances
    B = 4
    items = synth(B * (3 if rank == 0 else 2), seed=100 + rank)
    m = build(dev, items, B)
    steps, empties = 0, 0
    for _ in range(6):
        loss, err, step, empty = m.run_train_step(None, 1, 1.0, compute_error_rate=False)
        steps += 1
        if empty:
            empties = steps
            break
    flags = torch.tensor([float(steps), float(m.global_step)], device=dev)
    gathered = [torch.zeros_like(flags) for _ in range(world)]
    dist.all_gather(gathered, flags)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks fell out of lockstep: %s" % gathered
    ref = m.params.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, m.params), "parameters differ between ranks"
    # every rank saw the end of the epoch in the same call (the third: rank 1 ran dry there)
    assert empties == 3 and m.global_step == 3, (empties, m.global_step)
    # 3. N ranks x 1 mini-batch == 1 process x N accumulated mini-batches (each mini-batch from a zero RNN state)
    all_items = [synth(B, seed=500 + r) for r in range(world)]

    def accumulate(model, item_sets):
        model.start_batch(None, True)
        for its in item_sets:
            ds = model.build_dataset(its, B, 100, 50, "fbank", rs.ENGLISH_CHAR_MAP, device=dev)
            for x_d, len_d, dense in ds:
                model.rnn_state.zero_()
                model.step_on_batch(x_d, len_d, model.sparse_labels_from_dense(dense, True), True, False)

    md = build(dev, all_items[rank], B)
    accumulate(md, [all_items[rank]])
    md.end_batch(None, True, rnn_state_reset_ratio=1.0)          # all-reduce over the ranks, clip, Adam
    dp_grads = md.grads.clone()                                   # the all-reduced (summed) gradient
    if rank == 0:
        single = build(dev, all_items[0], B)
        accumulate(single, all_items)
        # the update without any collective: this process pretends to be alone
        import rnn_speech_b200.acoustic_model as am_mod
        saved = am_mod.allreduce_sum_
        am_mod.allreduce_sum_ = lambda t: t
        try:
            single.end_batch(None, True, rnn_state_reset_ratio=1.0)
        finally:
            am_mod.allreduce_sum_ = saved
        # (gradients, not parameters: Adam's first step is lr * sign(g), which turns a last-bit difference of a
        #  near-zero entry -- the two schedules add the chunks in different orders -- into 2 * lr)
        diff = float((single.grads - dp_grads).abs().max() / single.grads.abs().max())
        print("dp%d vs mini_batch_size=%d: gradient difference %.3e of the largest entry" % (world, world, diff))
        assert diff < 1e-5
    dist.barrier()
    if rank == 0:
        print("mgpu_lockstep ok: world %d, in-library NCCL all-reduce, unequal shards in lockstep (3 steps, epoch end agreed)" % world)
    rsdist.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
