import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rnn_speech_b200 as rs
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
B, n = 32, 160000
pcm = torch.from_numpy((0.1 * rng.standard_normal(B * n)).astype(np.float32)).to(dev)
off = torch.from_numpy(np.arange(B + 1, dtype=np.int64) * n).to(dev)
ap = rs.AudioProcessor(1000, 'fbank', device=dev)
f, nf = ap.features_device(pcm, off, B, n, 16000)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ap.features_device(pcm, off, B, n, 16000, out=f, nframes=nf)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print('fbank 32 x 10 s: %.3f ms per batch = %.0f GB/s of algorithmic bytes' % (ms, 32 * 1119040 / ms / 1e6))
