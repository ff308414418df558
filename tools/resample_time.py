"""Timing of the audio front door kernels (csrc/resample.cu) on one B200: 32 utterances of 10 s, 16 kHz int16 mono
-> float32 at 22 050 Hz.  CUDA events on the launching stream; algorithmic bytes = 2 n read + 4 n_out written."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rnn_speech_b200 as rs

lib = rs._lib
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
B, n = 32, 160000
out = {}
for name, sr_in, fmt, dtype in (("s16_16k_to_22k", 16000, lib.PCM_S16, np.int16), ("f32_16k_to_22k", 16000, lib.PCM_F32, np.float32),
                                ("s16_48k_to_22k", 48000, lib.PCM_S16, np.int16)):
    host = (3000 * rng.standard_normal(B * n)).astype(dtype) if dtype == np.int16 else \
        (0.1 * rng.standard_normal(B * n)).astype(np.float32)
    src = torch.from_numpy(host).to(dev)
    n_out = int(lib.raw("rs_resample_num_samples")(n, sr_in, 22050))
    in_off = torch.from_numpy(np.arange(B + 1, dtype=np.int64) * n).to(dev)
    out_off = torch.from_numpy(np.arange(B + 1, dtype=np.int64) * n_out).to(dev)
    dst = torch.empty((B * n_out,), dtype=torch.float32, device=dev)
    ws = torch.empty((int(lib.raw("rs_resample_workspace_bytes")(B, n_out)),), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: lib.call("rs_resample_forward", src.data_ptr(), fmt, 1, in_off.data_ptr(), B, n_out, sr_in, 22050,
                           dst.data_ptr(), out_off.data_ptr(), ws.data_ptr(), ws.numel(), st)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nbytes = B * (n * host.itemsize + 4 * n_out)
    out[name] = {"ms_per_batch": round(ms, 4), "algorithmic_GBps": round(nbytes / ms / 1e6, 1),
                 "utterances": B, "seconds_each": n / sr_in, "bytes": nbytes}
print(json.dumps(out))
