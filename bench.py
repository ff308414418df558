#!/usr/bin/env python
"""Headline benchmark: utterances/sec of one full acoustic-model training step
(features -> LSTM stack forward -> CTC loss+grad -> backward -> [NCCL all-reduce]
-> clip + Adam) on BASELINE.json config 2: 3x768 LSTM, 120-dim fbank, per-GPU
batch 32 of 10 s synthetic 16 kHz audio (T = 998 frames), 80 labels.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...     # CPU restatement of the reference path

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM,
`e2e` through the public API with pinned HOST PCM (H2D inside the timed region,
loss read back every step).  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(L=3, H=768, F=120, C=80, B=32, seconds=10.0, sr=16000, Tmax=1000, lab_lo=60, lab_hi=120,
           keep_in=0.8, keep_out=0.5, lr=3e-4, clip=1)
WORKLOAD = ("cfg2: 3x768 LSTM, fbank-120, per-GPU batch 32 x 10 s @16 kHz (T=998), C=80, labels 60-120+EOS, "
            "dropout keep 0.8/0.5, grad clip 1, Adam")


def synth_batch(rng, B, n, lab_lo, lab_hi):
    sigs = [(0.1 * rng.standard_normal(n)).astype(np.float32) for _ in range(B)]
    labs = [np.append(rng.integers(1, 79, size=rng.integers(lab_lo, lab_hi + 1)), 79).astype(np.int32)
            for _ in range(B)]
    return sigs, labs


def train_flops_per_utt(c, T):
    per_frame = 2 * c["F"] * c["H"] + c["L"] * 16 * c["H"] ** 2 + 2 * c["H"] * c["C"]
    return 3.0 * per_frame * T


# --------------------------------------------------------------------------- CPU arm
def cpu_step(sigs, labs, params, c, dtype=np.float32):
    """One training step of the restated reference CPU path (oracle/): features,
    forward, CTC, backward, clip + Adam.  Returns seconds."""
    from oracle import ctc, features, model, optim
    L, H, F, C = c["L"], c["H"], c["F"], c["C"]
    t0 = time.perf_counter()
    feats, lens = [], []
    for s in sigs:
        f, n = features.fbank(s, c["sr"], c["Tmax"])
        feats.append(f)
        lens.append(min(n, c["Tmax"]))
    T = max(lens)
    x = np.zeros((T, len(sigs), F), dtype)
    for b, f in enumerate(feats):
        x[:len(f), b] = f
    lens = np.array(lens)
    logits, _, cache = model.forward(params, x, lens, L, H, keep_in=c["keep_in"], keep_out=c["keep_out"], seed=1,
                                     dtype=dtype)
    loss, dlogits = ctc.ctc_loss_and_grad(logits, labs, lens)
    grads = model.backward(params, cache, dlogits, L, H, dtype=dtype)
    flat_g = model.flatten(grads, L, H, F, C)
    flat_p = model.flatten(params, L, H, F, C)
    optim.clip_adam_step(flat_p, flat_g, np.zeros_like(flat_g), np.zeros_like(flat_g), 1, c["lr"], c["clip"])
    return time.perf_counter() - t0, float(np.mean(loss / lens))


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, timed on
    the host cores.  TensorFlow-1 / librosa cannot be installed (no network, no
    Py3.12 wheels), so this is the oracle port (DESIGN.md 'Reference arm')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import model
    c = dict(CFG)
    total = args.steps + args.warmup
    b_ref = int(max(2, min(c["B"], 96 // max(total, 1))))
    rng = np.random.default_rng(0)
    params = model.init_params(c["L"], c["H"], c["F"], c["C"], seed=0)
    sigs, labs = synth_batch(rng, b_ref, int(c["seconds"] * c["sr"]), c["lab_lo"], c["lab_hi"])
    for _ in range(args.warmup):
        cpu_step(sigs, labs, params, c)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_step(sigs, labs, params, c)
        t += dt
    value = b_ref * args.steps / t
    cores = cpu_threads()
    line = {
        "impl": "reference", "metric": "utterances/sec", "value": value, "unit": "utt/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_batch": b_ref},
        "cpu_baseline": {"value": value, "unit": "utt/s", "cores": cores, "kind": "port",
                         "sample": "%d utterances of 10 s per step (of the batch of 32), full training step, "
                                   "numpy/OpenBLAS fp32 on %s" % (b_ref, cpu_model_name())},
        "e2e": {"value": value, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import rnn_speech_b200 as rs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    c = dict(CFG)
    n = int(c["seconds"] * c["sr"])
    rng = np.random.default_rng(1234 + rank)
    sigs, labs = synth_batch(rng, c["B"], n, c["lab_lo"], c["lab_hi"])

    ap = rs.AudioProcessor(c["Tmax"], "fbank", device=dev)
    m = rs.AcousticModel(c["L"], c["H"], c["B"], c["Tmax"], 600, c["F"], False, c["C"], device=dev, seed=0)
    m.create_training_rnn(c["keep_in"], c["keep_out"], c["clip"], c["lr"], 0.33)
    m.initialize(None)
    m.enable_timing()
    launch_count = rs._lib.raw("rs_launch_count")

    # device-resident inputs for `value`
    offsets = np.arange(c["B"] + 1, dtype=np.int64) * n
    pcm_host = torch.from_numpy(np.concatenate(sigs)).pin_memory()
    pcm_d = pcm_host.to(dev)
    off_d = torch.from_numpy(offsets).to(dev)
    # `value`: PCM resident in HBM.  The feature kernels of step s+1 run on a low-priority side stream while step s
    # trains (two feature buffers alternate; events order producer and consumer), as the input pipeline does in e2e.
    feats = [torch.empty((c["Tmax"], c["B"], c["F"]), dtype=torch.float32, device=dev) for _ in range(2)]
    nfr = [torch.empty((c["B"],), dtype=torch.int32, device=dev) for _ in range(2)]
    side = torch.cuda.Stream(device=dev, priority=0)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    state = {"slot": 0}

    def issue_features(slot):
        with torch.cuda.stream(side):
            side.wait_event(consumed[slot])
            ap.features_device(pcm_d, off_d, c["B"], n, c["sr"], time_major=True, out=feats[slot], nframes=nfr[slot])
            ready[slot].record(side)

    for ev in consumed:
        ev.record()
    issue_features(0)

    def step_resident():
        slot = state["slot"]
        issue_features(1 - slot)                      # next step's features, concurrent with this step
        torch.cuda.current_stream().wait_event(ready[slot])
        m.start_batch(None, True)
        m.step_on_batch(feats[slot], nfr[slot], labs, compute_gradients=True, compute_error_rate=False)
        m.apply_gradients()
        consumed[slot].record()
        state["slot"] = 1 - slot

    prefetch = rs.BatchPrefetcher(ap)
    pending = [prefetch.submit(sigs, c["sr"], time_major=True)]

    def step_e2e():
        # public API with HOST buffers.  Every step stages, copies (pinned, H2D) and featurises ONE mini-batch --
        # the next one, on the prefetcher's side stream, as the reference's tf.data pipeline prefetches -- trains on
        # the one submitted a step earlier, and reads the mean loss back.
        f, nf = pending[0].result()
        pending[0] = prefetch.submit(sigs, c["sr"], time_major=True)
        m.start_batch(None, True)
        m.step_on_batch(f, nf, labs, compute_gradients=True, compute_error_rate=False)
        mean_loss, _, _ = m.end_batch(None, True, rnn_state_reset_ratio=1.0)
        return float(mean_loss)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launch_count() - l0

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = timed(step_resident, args.steps)
    rec_f, rec_b = m.recurrent_ms()
    trace = m.recurrent_trace()
    chunk = int(os.environ.get("RS_TC_CHUNK", "128"))
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(max(3, args.warmup)):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    utts = c["B"] * world * args.steps
    value = utts / (ms / 1e3)
    e2e = utts / (ms_e2e / 1e3)
    if rank != 0:
        return
    T = 998
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
    # dominant kernels: the recurrent kernels (forward + backward).  With the pipelined schedule one layer is
    # several launches of <= `chunk` steps each: `achieved` = algorithmic flops of one launch (2*B*H*4H per step) over
    # the average launch duration (CUDA events on the launching streams); launch_ms lists the per-layer sums.
    durs = [b - a for d in (0, 1) for l in trace[d] for (a, b) in l]          # every recurrent launch of the step, ms
    n_launches = max(1, len(durs))
    steps_per_launch = T * 2.0 * c["L"] / n_launches                        # = the chunk length, averaged
    rec_flops_launch = 2.0 * c["B"] * c["H"] * 4 * c["H"] * steps_per_launch   # h_{t-1} @ Wh per launch
    rec_ms = float(np.mean(durs)) if durs else float(np.mean(rec_f + rec_b))
    achieved = rec_flops_launch / (rec_ms / 1e3) / 1e12
    tc = bool(m.uses_tensor_cores)
    n_launch = [len(x) for x in trace[0]] + [len(x) for x in trace[1]]

    def busy(intervals):
        """length of the union of (start, stop) intervals"""
        tot, end = 0.0, -1e30
        for a, b in sorted(intervals):
            if b > end:
                tot += b - max(a, end)
                end = b
        return tot
    rec_busy = busy([iv for l in trace[0] for iv in l]) + busy([iv for l in trace[1] for iv in l])
    roofline = {"kernel": ("rec_ts_fwd_kernel / rec_ts_bwd_kernel (persistent tcgen05 recurrent kernels, weights resident "
                           "in tensor memory; %d launches per layer per direction, <= %d steps each, layers overlapped as a "
                           "wavefront)" % (max(n_launch), chunk)) if tc else "lstm_rec_fwd_kernel / lstm_rec_bwd_kernel (fp32 FFMA)",
                "bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": achieved / tensor_peak,
                # dram__bytes_read+write per launch of 128 steps from profiles/r01d_ncu_final.txt (fwd 88.7 MB, bwd 102.7 MB)
                "traffic": 95.7e6 if tc else None, "peak_source": peak_src, "avg_launch_ms": rec_ms,
                "steps_per_launch": steps_per_launch,
                "launch_ms": {"fwd": rec_f, "bwd": rec_b}, "launches_per_layer": max(n_launch),
                "sum_launch_ms": sum(rec_f) + sum(rec_b),
                "share_of_step": rec_busy / (ms / args.steps),
                "note": "algorithmic flops = 2*B*H*4H per step (151 MFLOP at cfg-2, 19.3 GFLOP per 128-step launch; the "
                        "bf16x3 forward issues 3x that on the tensor pipe).  The recurrence is a chain of T dependent steps with a grid-wide "
                        "exchange of h per step: latency-bound, not tensor-bound (DESIGN.md 'Recurrent step budget'); "
                        "launch_ms sums a layer's chunk launches, which run concurrently with other layers' (sum_launch_ms "
                        "exceeds the step); share_of_step = time during which at least one recurrent launch is running"}
    line = {
        "metric": "utterances/sec", "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16x3" if tc else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": c["B"] * world, "parallelism": "dp%d" % world,
                   "arithmetic": ("bf16 tensor cores with a 3-term hi/lo split (fp32-grade products, fp32 accumulate) "
                                  "forward and in all batched GEMMs; plain bf16 in the backward dh recurrence; fp64 "
                                  "feature extraction; fp32 CTC / Adam") if tc else "fp32",
                   "schedule": "time chunks of %d steps, layers as a wavefront (2 recurrent launches in flight), chunk GEMMs "
                               "and weight-gradient GEMMs on the remaining SMs" % chunk,
                   "l2": "per-step working set (activations 2.1 GB + 57 MB params x4) exceeds the 126 MB L2; no flush needed",
                   "train_tflop_per_step": train_flops_per_utt(c, T) * c["B"] * world / 1e12},
        "e2e": {"value": e2e, "unit": "utt/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(pcm_host.numel() * 4 + offsets.nbytes + sum(l.nbytes for l in labs)
                                          + 4 * (c["B"] + 1)),
                "d2h_bytes_per_step": 12,
                "note": "AudioProcessor.process_batch (pinned staging + H2D + feature kernels) of the next mini-batch runs "
                        "on BatchPrefetcher's side stream while AcousticModel.step_on_batch / end_batch train on the "
                        "current one; one mini-batch is staged, copied and featurised per timed step"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import model
        params = model.init_params(c["L"], c["H"], c["F"], c["C"], seed=0)
        nb = args.cpu_sample
        dt, _ = cpu_step(sigs[:nb], labs[:nb], params, c)
        line["cpu_baseline"] = {"value": nb / dt, "unit": "utt/s", "cores": cpu_threads(), "kind": "port",
                                "sample": "1 full training step on %d of the 32 utterances (10 s each), "
                                          "numpy/OpenBLAS fp32, %s, os.cpu_count=%s"
                                          % (nb, cpu_model_name(), os.cpu_count())}
    print(json.dumps(line))


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=8, help="utterances in the cpu_baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
        _shutdown()


if __name__ == "__main__":
    main()
