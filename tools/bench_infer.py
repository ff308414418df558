"""BASELINE config 5: inference only (the stt.py --file path) -- batch 256 of 5 s clips at 16 kHz (T = 498 frames),
3x768 LSTM: features + forward + CTC greedy decode through the public API with HOST PCM.  Prints clips/sec and the
p50 / p90 latency of one batch.  Not a bench.py line (bench.py measures config 2); numbers go to DESIGN.md.

    python tools/bench_infer.py [--batch 256] [--seconds 5] [--iters 20]
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
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    L, H, F, C, sr = 3, 768, 120, 80, 16000
    n = int(a.seconds * sr)
    rng = np.random.default_rng(0)
    sigs = [(0.1 * rng.standard_normal(n)).astype(np.float32) for _ in range(a.batch)]
    proc = rs.AudioProcessor(500, "fbank", device=dev)
    m = rs.AcousticModel(L, H, a.batch, 500, 600, F, False, C, device=dev, seed=0)
    m.create_forward_rnn()
    m.initialize(None)

    def once():
        feats, nframes = proc.process_batch(sigs, sr, time_major=True)
        lens = torch.clamp(nframes, max=500)
        logits = m.forward(feats, lens, training=False, keep_state=False)
        ids, out_len = m.greedy_decode(logits, lens)
        return ids.cpu(), out_len.cpu()          # the decoded ids reach the host: end of the request

    for _ in range(3):
        once()
    lat = []
    for _ in range(a.iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        once()
        lat.append(time.perf_counter() - t0)
    lat = np.array(lat) * 1e3
    print(json.dumps({"workload": "cfg5: inference, batch %d x %.0f s @16 kHz, 3x768 LSTM, fbank-120, greedy decode" % (a.batch, a.seconds),
                      "clips_per_sec": a.batch / (np.median(lat) / 1e3), "p50_ms": float(np.median(lat)),
                      "p90_ms": float(np.percentile(lat, 90)), "tensor_cores": bool(m.uses_tensor_cores),
                      "batch_tiles": len(m._tiles) if m._tiles else 1}))


if __name__ == "__main__":
    main()
