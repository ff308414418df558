"""BASELINE config 4 on one GPU: 5x1024 LSTM, 120-dim fbank, per-GPU batch 16 of variable-length utterances (2-20 s at
16 kHz, sorted by duration into batches as the reference's dataset_size_ordering does), full training step through the
public API with HOST PCM.  Not a bench.py line (bench.py measures config 2); numbers go to DESIGN.md.

    python tools/bench_cfg4.py [--batches 6]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import rnn_speech_b200 as rs
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=6)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    L, H, F, C, B, sr, Tmax = 5, 1024, 120, 80, 16, 16000, 2000
    rng = np.random.default_rng(0)
    secs = np.sort(rng.uniform(2.0, 20.0, size=B * a.batches))
    batches = []
    for i in range(a.batches):
        sigs = [(0.1 * rng.standard_normal(int(s * sr))).astype(np.float32) for s in secs[i * B:(i + 1) * B]]
        labs = [np.append(rng.integers(1, 79, size=max(4, int(len(x) / sr * 12))), 79).astype(np.int32) for x in sigs]
        batches.append((sigs, labs))
    proc = rs.AudioProcessor(Tmax, "fbank", device=dev)
    m = rs.AcousticModel(L, H, B, Tmax, 600, F, False, C, device=dev, seed=0)
    m.create_training_rnn(0.8, 0.5, 1, 3e-4, 0.33)
    m.initialize(None)
    pre = rs.BatchPrefetcher(proc)

    def epoch():
        n_utt, frames = 0, 0
        ticket = pre.submit(batches[0][0], sr)
        for i, (sigs, labs) in enumerate(batches):
            feats, nframes = ticket.result()
            if i + 1 < len(batches):
                ticket = pre.submit(batches[i + 1][0], sr)
            T = int(min(int(nframes.max()), Tmax))
            m.start_batch(None, True)
            m.step_on_batch(feats[:T].contiguous(), torch.clamp(nframes, max=Tmax), labs, compute_gradients=True,
                            compute_error_rate=False)
            loss = m.end_batch(None, True, rnn_state_reset_ratio=1.0)[0]
            n_utt += len(sigs)
            frames += int(nframes.sum())
        return n_utt, frames, float(loss)

    epoch()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_utt, frames, loss = epoch()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"workload": "cfg4: 5x1024 LSTM, fbank-120, batch 16 of 2-20 s utterances sorted by duration, %d batches" % a.batches,
                      "utt_per_sec": n_utt / dt, "frames_per_sec": frames / dt, "ms_per_batch": 1e3 * dt / a.batches,
                      "audio_seconds_per_sec": float(secs.sum()) / dt, "tensor_cores": bool(m.uses_tensor_cores),
                      "last_loss": loss}))


if __name__ == "__main__":
    main()
