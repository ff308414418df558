"""Diagnostics run on the GPU box: prints error magnitudes of each kernel family
against the oracle (more detail than the pass/fail of pytest)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rnn_speech_b200 as rs   # noqa: E402
from oracle import ctc, features, model   # noqa: E402

dev = torch.device("cuda:0")


def ctc_diag():
    for (T, B, lo, hi, full) in ((40, 5, 3, 12, False), (200, 8, 5, 40, False), (200, 8, 5, 40, True), (998, 32, 60, 120, True)):
        rng = np.random.default_rng(T + B)
        C = 80
        logits = (1.5 * rng.standard_normal((T, B, C))).astype(np.float32)
        labs = [np.append(rng.integers(1, 79, size=rng.integers(lo, hi + 1)), 79).astype(np.int32) for _ in range(B)]
        lens = np.full(B, T, np.int32) if full else rng.integers(T // 2, T + 1, size=B).astype(np.int32)
        m = rs.AcousticModel(1, 8, B, T, 600, 8, False, C, device=dev)
        m.create_forward_rnn()
        lg = torch.from_numpy(logits).to(dev)
        ln = torch.from_numpy(lens).to(dev)
        loss, grad = m.ctc_loss(lg, labs, ln)
        torch.cuda.synchronize()
        wl, wg = ctc.ctc_loss_and_grad(logits, labs, lens)
        loss, grad = loss.cpu().numpy(), grad.cpu().numpy()
        print("CTC T=%d B=%d full=%s: loss rel err per item %s" % (T, B, full, np.array2string(np.abs(loss - wl) / np.abs(wl), precision=2)))
        e = np.abs(grad - wg)
        print("    grad max err per item", np.array2string(e.max(axis=(0, 2)), precision=2), "rowsum gpu/oracle",
              float(np.abs(grad.sum(-1)).max()), float(np.abs(wg.sum(-1)).max()))
        t, b, k = np.unravel_index(e.argmax(), e.shape)
        print("    worst at t=%d b=%d k=%d: gpu %.6f oracle %.6f (len %d)" % (t, b, k, grad[t, b, k], wg[t, b, k], lens[b]))


def fbank_diag():
    rng = np.random.default_rng(0)
    sigs = [(0.1 * rng.standard_normal(160000)).astype(np.float32) for _ in range(32)]
    ap = rs.AudioProcessor(1000, "fbank", device=dev)
    feats, nframes = ap.process_batch(sigs, 16000, time_major=True)
    f = feats.cpu().numpy()
    print("fbank nframes", nframes.cpu().numpy()[:4])
    for b in (0, 7, 31):
        want, _ = features.fbank(sigs[b], 16000, 1000)
        e = np.abs(f[:998, b] - want)
        print("fbank b=%d: max err static %.3e delta %.3e ddelta %.3e; mean|static mean| %.2e" %
              (b, e[:, :40].max(), e[:, 40:80].max(), e[:, 80:].max(), np.abs(f[:998, b, :40].mean(0)).max()))
    print("fbank tail zero:", bool(np.all(f[998:] == 0)))
    again, _ = ap.process_batch(sigs, 16000, time_major=True)
    print("fbank deterministic:", bool(torch.equal(again, feats)))


def tc_diag():
    for (N, K) in ((32, 64), (32, 256), (64, 128)):
        rng = np.random.default_rng(N + K)
        A = rng.standard_normal((128, K)).astype(np.float32)
        B = rng.standard_normal((N, K)).astype(np.float32)
        want = A.astype(np.float64) @ B.astype(np.float64).T
        Ad, Bd = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
        for split in (0, 1):
            D = torch.full((128, N), float("nan"), dtype=torch.float32, device=dev)
            rs._lib.call("rs_tc_selftest", Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, split,
                         torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got = D.cpu().numpy()
            e = np.abs(got - want)
            print("tc N=%d K=%d split=%d: max rel err %.3e; nan %d; row-err profile %s" % (
                N, K, split, e.max() / np.abs(want).max(), int(np.isnan(got).sum()),
                np.array2string(e.max(axis=1)[::16], precision=2)))
            if e.max() / np.abs(want).max() > 0.05:
                print("   got[0,:8]", got[0, :8], "\n   want[0,:8]", want[0, :8])
                print("   got[1,:8]", got[1, :8], "\n   want[1,:8]", want[1, :8])
                print("   got[64,:8]", got[64, :8], "\n   want[64,:8]", want[64, :8])


if __name__ == "__main__":
    which = sys.argv[1:] or ["ctc", "fbank"]
    if "ctc" in which:
        ctc_diag()
    if "fbank" in which:
        fbank_diag()
    if "tc" in which:
        tc_diag()
