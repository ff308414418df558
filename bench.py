#!/usr/bin/env python
"""Benchmarks of the acoustic-model hot path on B200 (one JSON line per run, rank 0).

    python bench.py --gpus 1 --steps 10 --warmup 3                 # headline: BASELINE config 2
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                            # CPU restatement of the reference path
    python bench.py --config cfg4 | cfg5                            # BASELINE configs 4 and 5 (same JSON schema)

cfg2 (default): utterances/sec of one full training step (features -> LSTM stack forward -> CTC loss+grad ->
    backward -> [NCCL all-reduce] -> clip + Adam): 3x768 LSTM, 120-dim fbank, per-GPU batch 32 of 10 s synthetic
    16 kHz audio (T = 998 frames), 80 labels.
cfg4: the same step on a 5x1024 LSTM, per-GPU batch 16 of 2-20 s utterances sorted by duration into batches.
cfg5: inference only (features + forward + greedy decode), batch 256 of 5 s clips; clips/sec and batch latency.

`value` is timed with inputs resident in HBM, `e2e` through the public API with pinned HOST PCM (H2D inside the
timed region, result read back every step).  The training configs time the same K steps twice: first as they run in
production (`value`), then with CUDA events around every recurrent launch and chunk GEMM (the `roofline` block: the
events cost 0.25 ms per step, `roofline.instrumented_ms_per_step`).  See DESIGN.md "Measurement".
"""
import os
import sys


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _is_reference_arm(argv):
    for i, a in enumerate(argv):
        if a == "--impl" and i + 1 < len(argv) and argv[i + 1] == "reference":
            return True
        if a == "--impl=reference":
            return True
    return False


# The CPU arm must use every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, and the BLAS libraries
# read these variables when they are first loaded -- so they are set here, before numpy / torch are imported.
if _is_reference_arm(sys.argv):
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_k] = str(_host_threads())

# (the package sets this at import too; here it is certain to precede the CUDA context: see rnn-speech_b200/__init__.py)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import argparse          # noqa: E402
import json              # noqa: E402
import subprocess        # noqa: E402
import threading         # noqa: E402
import time              # noqa: E402

import numpy as np       # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    "cfg2": dict(L=3, H=768, F=120, C=80, B=32, seconds=10.0, sr=16000, Tmax=1000, lab_lo=60, lab_hi=120,
                 keep_in=0.8, keep_out=0.5, lr=3e-4, clip=1, train=True, metric="utterances/sec", unit="utt/s",
                 workload=("cfg2: 3x768 LSTM, fbank-120, per-GPU batch 32 x 10 s @16 kHz (T=998), C=80, labels 60-120+EOS, "
                           "dropout keep 0.8/0.5, grad clip 1, Adam")),
    "cfg4": dict(L=5, H=1024, F=120, C=80, B=16, sec_lo=2.0, sec_hi=20.0, sr=16000, Tmax=2000, nbatches=8,
                 keep_in=0.8, keep_out=0.5, lr=3e-4, clip=1, train=True, metric="utterances/sec", unit="utt/s",
                 workload=("cfg4: 5x1024 LSTM, fbank-120, per-GPU batch 16 of 2-20 s utterances @16 kHz sorted by duration into "
                           "8 batches (T=198..1998, padded to the batch maximum), C=80, ~12 labels/s + EOS, dropout keep 0.8/0.5, "
                           "grad clip 1, Adam; a step = one batch, the 8 batches cycle")),
    "cfg5": dict(L=3, H=768, F=120, C=80, B=256, seconds=5.0, sr=16000, Tmax=500, train=False,
                 metric="clips/sec", unit="clips/s",
                 workload=("cfg5: inference, batch 256 x 5 s @16 kHz (T=498), 3x768 LSTM, fbank-120, features + forward + "
                           "CTC greedy decode; a step = one batch")),
}


def synth_batch(rng, B, n, lab_lo, lab_hi):
    sigs = [(0.1 * rng.standard_normal(n)).astype(np.float32) for _ in range(B)]
    labs = [np.append(rng.integers(1, 79, size=rng.integers(lab_lo, lab_hi + 1)), 79).astype(np.int32)
            for _ in range(B)]
    return sigs, labs


def synth_cfg4(rng, c):
    """nbatches batches of B utterances, durations uniform in [sec_lo, sec_hi], sorted by duration
    (the reference's dataset_size_ordering, models/SpeechRecognizer.py:80-81)."""
    secs = np.sort(rng.uniform(c["sec_lo"], c["sec_hi"], size=c["B"] * c["nbatches"]))
    batches = []
    for i in range(c["nbatches"]):
        sigs = [(0.1 * rng.standard_normal(int(s * c["sr"]))).astype(np.float32) for s in secs[i * c["B"]:(i + 1) * c["B"]]]
        labs = [np.append(rng.integers(1, 79, size=max(4, int(len(x) / c["sr"] * 12))), 79).astype(np.int32) for x in sigs]
        batches.append((sigs, labs))
    return batches, float(secs.sum())


def fwd_flops_per_frame(c):
    return 2 * c["F"] * c["H"] + c["L"] * 16 * c["H"] ** 2 + 2 * c["H"] * c["C"]


# --------------------------------------------------------------------------- CPU arm
def cpu_step(sigs, labs, params, c, leg="numpy", train=True):
    """One step of the restated reference CPU path (oracle/): features (one process per utterance, as tf.data's
    parallel map), the LSTM stack (leg = "numpy": OpenBLAS, "torch": MKL; every host thread), CTC loss + gradient
    (one process per item, as TF's CTCLoss shards the batch), clip + Adam.  Returns (seconds, display loss)."""
    from oracle import ctc, model, model_torch, optim, parallel
    L, H, F, C = c["L"], c["H"], c["F"], c["C"]
    t0 = time.perf_counter()
    out = parallel.fbank_batch(sigs, c["sr"], c["Tmax"])
    lens = np.array([min(n, c["Tmax"]) for _, n in out])
    T = int(lens.max())
    x = np.zeros((T, len(sigs), F), np.float32)
    for b, (f, _) in enumerate(out):
        x[:len(f), b] = f
    keep_in, keep_out = (c["keep_in"], c["keep_out"]) if train else (1.0, 1.0)
    if leg == "torch":
        tl, cache = model_torch.forward(params, x, lens, L, H, keep_in=keep_in, keep_out=keep_out, seed=1)
        logits = tl.numpy()
    else:
        logits, _, cache = model.forward(params, x, lens, L, H, keep_in=keep_in, keep_out=keep_out, seed=1,
                                         dtype=np.float32, keep_cache=train)
    if not train:
        ctc.greedy_decode(logits, lens)
        return time.perf_counter() - t0, 0.0
    loss, dlogits = parallel.ctc_batch(logits, labs, lens)
    if leg == "torch":
        grads = model_torch.backward(cache, dlogits, L, H)
    else:
        grads = model.backward(params, cache, dlogits, L, H, dtype=np.float32)
    flat_g = model.flatten(grads, L, H, F, C)
    flat_p = model.flatten(params, L, H, F, C)
    optim.clip_adam_step(flat_p, flat_g, np.zeros_like(flat_g), np.zeros_like(flat_g), 1, c["lr"], c["clip"])
    return time.perf_counter() - t0, float(np.mean(loss / np.maximum(lens, 1)))


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else _host_threads()
    except Exception:
        return _host_threads()


def cpu_pick_leg(sigs, labs, params, c, train):
    """Time one step of each model leg (numpy/OpenBLAS and torch/MKL) and keep the faster."""
    import torch
    torch.set_num_threads(_host_threads())
    times = {}
    for leg in ("numpy", "torch"):
        times[leg], _ = cpu_step(sigs, labs, params, c, leg=leg, train=train)
    return min(times, key=times.get), times


def cpu_workload(c, name, rng):
    """(sigs, labs) of ONE step of the config on the CPU arm: the whole per-GPU batch."""
    if name == "cfg4":
        batches, _ = synth_cfg4(rng, c)
        return batches[len(batches) // 2]           # a median-duration batch
    return synth_batch(rng, c["B"] if name != "cfg5" else 32, int(c["seconds"] * c["sr"]), c.get("lab_lo", 60), c.get("lab_hi", 120))


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, timed on the host cores, every step the
    SAME workload as the GPU arm's step (the full per-GPU batch).  TensorFlow-1 / librosa cannot be installed (no
    network, no Py3.12 wheels), so this is the oracle port (DESIGN.md 'Reference arm'); under torchrun only rank 0
    works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import model, parallel
    c = dict(CONFIGS[args.config])
    rng = np.random.default_rng(0)
    params = model.init_params(c["L"], c["H"], c["F"], c["C"], seed=0)
    sigs, labs = cpu_workload(c, args.config, rng)
    parallel.pool()                                   # worker start-up is not part of a step
    leg, leg_times = cpu_pick_leg(sigs, labs, params, c, c["train"])        # also the first warm-up
    for _ in range(max(0, args.warmup - 2)):
        cpu_step(sigs, labs, params, c, leg=leg, train=c["train"])
    t = 0.0
    for _ in range(args.steps):
        dt, _ = cpu_step(sigs, labs, params, c, leg=leg, train=c["train"])
        t += dt
    parallel.close()
    units = len(sigs)
    value = units * args.steps / t
    cores = _host_threads()
    line = {
        "impl": "reference", "metric": c["metric"], "value": value, "unit": c["unit"], "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": c["workload"], "sample_batch": units, "same_config": units == c["B"],
                   "model_leg": leg, "model_leg_step_seconds": leg_times},
        "cpu_baseline": {"value": value, "unit": c["unit"], "cores": cores, "kind": "port",
                         "sample": "%d utterances per step (the per-GPU batch is %d), %s; features and CTC one process per "
                                   "utterance (%d workers), LSTM stack %s fp32 with %d BLAS threads, on %s"
                                   % (units, c["B"], "full training step" if c["train"] else "features + forward + greedy decode",
                                      cores, "torch-CPU/MKL" if leg == "torch" else "numpy/OpenBLAS", cpu_blas_threads(),
                                      cpu_model_name())},
        "e2e": {"value": value, "unit": c["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


# --------------------------------------------------------------------------- GPU arm: shared pieces
class Harness(object):
    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), \
            "bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm"
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass

    def barrier(self):
        torch = self.torch
        torch.cuda.synchronize()
        if self.world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def timed(self, fn, steps, launch_count=None):
        """K calls of fn bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = launch_count() if launch_count else 0
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, (launch_count() - l0) if launch_count else 0

    def tensor_peak(self):
        p = self.peaks.get("bf16_tflops_sustained")
        return (p, "measured (MEASURED_PEAKS.json bf16_tflops_sustained)") if p else (1400.0, "fallback 1.4 PFLOP/s sustained")

    def hbm_peak(self):
        p = self.peaks.get("hbm_gbs")
        return (p, "measured (MEASURED_PEAKS.json hbm_gbs)") if p else (6500.0, "fallback 6.5 TB/s")


def busy(intervals):
    """length of the union of (start, stop) intervals"""
    tot, end = 0.0, -1e30
    for a, b in sorted(intervals):
        if b > end:
            tot += b - max(a, end)
            end = b
    return tot


def recurrent_roofline(h, m, c, T, step_ms, directions, traffic):
    """roofline of the dominant kernels (the persistent recurrent kernels) from the library's own CUDA-event trace of
    the last step: achieved = algorithmic flops of one launch (2*B*H*4H per step) / average launch duration."""
    rec_f, rec_b = m.recurrent_ms()
    trace = m.recurrent_trace()
    chunk = int(os.environ.get("RS_TC_CHUNK", "128"))
    tensor_peak, peak_src = h.tensor_peak()
    import rnn_speech_b200.acoustic_model as _am
    B = c["B"] if c["B"] <= _am.TC_MAX_BATCH else _am.TC_MAX_BATCH          # rows per recurrent launch (batch tiles above that)
    durs = [b - a for d in directions for l in trace[d] for (a, b) in l]
    n_launches = max(1, len(durs))
    steps_per_launch = T * float(len(directions)) * c["L"] / n_launches
    rec_flops_launch = 2.0 * B * c["H"] * 4 * c["H"] * steps_per_launch
    rec_ms = max(float(np.mean(durs)) if durs else float(np.mean(rec_f + rec_b)), 1e-9)
    achieved = rec_flops_launch / (rec_ms / 1e3) / 1e12
    n_launch = [len(x) for d in directions for x in trace[d]]
    rec_busy = sum(busy([iv for l in trace[d] for iv in l]) for d in directions)
    # a mini-batch above the batch-tile size runs tile after tile: the trace is the last tile's
    n_tiles = len(m._tiles) if getattr(m, "_tiles", None) else 1
    rec_busy *= n_tiles
    tc = bool(m.uses_tensor_cores)
    # batches of 17..32 run as two 16-row chains per CTA with the validated exchange (lstm_rec_ts.cu dispatch)
    # (forward also takes batches of at most 16 rows, as one chain)
    two = 16 < B <= 32 and os.environ.get("RS_TS_XCHG", "1") != "0"
    fwd3 = two or (B <= 16 and os.environ.get("RS_TS_XCHG", "1") != "0" and os.environ.get("RS_TS_XCHG16", "1") != "0")
    names = " / ".join([("rec_ts_fwd3_kernel" if fwd3 else "rec_ts_fwd_kernel"),
                        ("rec_ts_bwd4_kernel" if two and os.environ.get("RS_TS_CHAINS_BWD", "1") != "0" else "rec_ts_bwd_kernel")][d]
                       for d in directions)
    return {"kernel": ("%s (persistent tcgen05 recurrent kernels, weights resident in tensor memory%s; %d launches per layer per "
                       "direction, <= %d steps each, layers overlapped as a wavefront)"
                       % (names, ", two 16-row chains per CTA, h_t / dgates_t exchanged through L2 with plain stores + a relaxed "
                                 "hint and validated by the tensor core (NaN fill pattern)" if two else "", max(n_launch), chunk))
            if tc else "lstm_rec_fwd_kernel / lstm_rec_bwd_kernel (fp32 FFMA)",
            "bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
            "traffic": traffic if tc else None, "peak_source": peak_src, "avg_launch_ms": rec_ms,
            "steps_per_launch": steps_per_launch,
            "launch_ms": {"fwd": rec_f, "bwd": rec_b}, "launches_per_layer": max(n_launch),
            "sum_launch_ms": sum(rec_f) + (sum(rec_b) if 1 in directions else 0.0),
            "share_of_step": rec_busy / step_ms, "batch_tiles": n_tiles,
            "note": "algorithmic flops = 2*B*H*4H per recurrent step and layer (the bf16x3 products issue 3x that on the tensor "
                    "pipe).  The recurrence is a chain of T dependent steps with a grid-wide exchange of h per step: "
                    "latency-bound, not tensor-bound (DESIGN.md 'Recurrent step budget'); launch_ms sums a layer's chunk "
                    "launches, which run concurrently with other layers' (sum_launch_ms exceeds the step); share_of_step = "
                    "time during which at least one recurrent launch is running"}


def arithmetic_note(tc):
    return ("bf16 tensor cores with a 3-term hi/lo split (fp32-grade products, fp32 accumulate) in every GEMM and in both "
            "recurrences (forward h @ Wh and backward dgates @ Wh^T); fp64 feature extraction; fp32 CTC / Adam") if tc else "fp32"


def event_ms(torch, fn, stream, reps):
    """median duration of fn() alone on `stream` (CUDA events on that stream)."""
    out = []
    with torch.cuda.stream(stream):
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            out.append(e0.elapsed_time(e1))
    return float(np.median(out))


# --------------------------------------------------------------------------- cfg2 / cfg4: training step
def run_train(args, name):
    h = Harness(args)
    torch = h.torch
    import rnn_speech_b200 as rs
    c = dict(CONFIGS[name])
    dev, rank, world = h.dev, h.rank, h.world
    rng = np.random.default_rng(1234 + rank)
    if name == "cfg2":
        n = int(c["seconds"] * c["sr"])
        batches = [synth_batch(rng, c["B"], n, c["lab_lo"], c["lab_hi"])]
        audio_seconds = c["seconds"] * c["B"]
    else:
        batches, total_sec = synth_cfg4(rng, c)
        audio_seconds = total_sec / len(batches)
    NB = len(batches)

    ap = rs.AudioProcessor(c["Tmax"], "fbank", device=dev)
    m = rs.AcousticModel(c["L"], c["H"], c["B"], c["Tmax"], 600, c["F"], False, c["C"], device=dev, seed=0)
    m.create_training_rnn(c["keep_in"], c["keep_out"], c["clip"], c["lr"], 0.33)
    m.initialize(None)
    # The CUDA events around every recurrent launch and chunk GEMM (rs_am_enable_timing: what the roofline block is made
    # of) cost 0.25 ms per step (profiles/r02d_sweep18.log: 13.73 -> 13.48 ms; they take the launch-to-launch overlap from
    # the streams that carry the schedule).  So `value`, `e2e` and `with_error_rate` are timed WITHOUT them, and the roofline
    # comes from a second timed region of the same K steps WITH them, right behind (roofline.measured_over).
    # RS_BENCH_TIMED_EVENTS=1: the events in every region, as before.
    events_everywhere = bool(os.environ.get("RS_BENCH_TIMED_EVENTS"))
    if events_everywhere:
        m.enable_timing()
    launch_count = rs._lib.raw("rs_launch_count")

    # per batch: device-resident PCM, offsets, frame counts known on the host
    res = []
    for sigs, labs in batches:
        lens = [len(s) for s in sigs]
        offsets = np.zeros(len(sigs) + 1, np.int64)
        np.cumsum(lens, out=offsets[1:])
        T = min(c["Tmax"], max(ap.num_frames(x, c["sr"]) for x in lens))
        res.append({"pcm": torch.from_numpy(np.concatenate(sigs)).to(dev), "off": torch.from_numpy(offsets).to(dev),
                    "max_n": max(lens), "T": T, "labs": labs, "sigs": sigs,
                    "h2d": int(4 * sum(lens) + offsets.nbytes + sum(l.nbytes for l in labs) + 4 * (len(sigs) + 1))})
    # `value`: PCM resident in HBM.  The feature kernels of step s+1 run on a low-priority side stream while step s
    # trains (two feature buffers alternate; events order producer and consumer), as the input pipeline does in e2e.
    feats = [torch.empty((c["Tmax"], c["B"], c["F"]), dtype=torch.float32, device=dev) for _ in range(2)]
    nfr = [torch.empty((c["B"],), dtype=torch.int32, device=dev) for _ in range(2)]
    side = torch.cuda.Stream(device=dev, priority=0)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    state = {"slot": 0, "i": 0, "err": False}

    def issue_features(slot, r):
        with torch.cuda.stream(side):
            side.wait_event(consumed[slot])
            ap.features_device(r["pcm"], r["off"], c["B"], r["max_n"], c["sr"], time_major=True, out=feats[slot], nframes=nfr[slot])
            ready[slot].record(side)

    for ev in consumed:
        ev.record()
    issue_features(0, res[0])

    diag_h2d = int(os.environ.get("RS_DIAG_H2D", "0"))       # diagnostic: a 20 MB pinned H2D copy per resident step
    if diag_h2d:
        dg_host = torch.empty(5 << 20, dtype=torch.float32).pin_memory()
        dg_dev = torch.empty(5 << 20, dtype=torch.float32, device=dev)
        dg_stream = torch.cuda.Stream(device=dev, priority=0)
        dg_ev = torch.cuda.Event()

    def step_resident():
        slot, i = state["slot"], state["i"]
        r = res[i % NB]
        if diag_h2d:
            if diag_h2d == 2:                                  # placed: behind the previous step's optimizer
                dg_ev.record()
                dg_stream.wait_event(dg_ev)
            with torch.cuda.stream(dg_stream):
                dg_dev.copy_(dg_host, non_blocking=True)
        issue_features(1 - slot, res[(i + 1) % NB])       # next step's features, concurrent with this step
        torch.cuda.current_stream().wait_event(ready[slot])
        m.start_batch(None, True)
        x = feats[slot] if r["T"] == c["Tmax"] or NB == 1 else feats[slot][:r["T"]]
        m.step_on_batch(x, nfr[slot], r["labs"], compute_gradients=True, compute_error_rate=state["err"])
        if state["err"]:
            m.end_batch(None, True, rnn_state_reset_ratio=1.0)
        else:
            m.apply_gradients()
        consumed[slot].record()
        state["slot"], state["i"] = 1 - slot, i + 1

    prefetch = rs.BatchPrefetcher(ap)
    gate_features = os.environ.get("RS_PREFETCH_GATE", "1") != "0"
    pending = [prefetch.submit(res[0]["sigs"], c["sr"], time_major=True)]
    e2e_i = [0]
    diag_e2e, diag_keep = os.environ.get("RS_DIAG_E2E", ""), []

    def step_e2e():
        # public API with HOST buffers.  Every step stages, copies (pinned, ONE H2D) and featurises ONE mini-batch --
        # the next one, on the prefetcher's side stream, as the reference's tf.data pipeline prefetches -- trains on
        # the one submitted a step earlier, and reads the mean loss back.
        i = e2e_i[0]
        r = res[i % NB]
        if diag_e2e:
            # diagnostic: no input pipeline at all (the features of the first ticket again and again); "c": no read either
            if not diag_keep:
                diag_keep.extend(pending[0].result())
            m.start_batch(None, True)
            m.step_on_batch(diag_keep[0], diag_keep[1], r["labs"], compute_gradients=True, compute_error_rate=False)
            if diag_e2e == "c":
                m.apply_gradients()
                return 0.0
            return float(m.end_batch(None, True, rnn_state_reset_ratio=1.0 if diag_e2e != "d" else 1e-9)[0])
        f, nf = pending[0].result()
        # (the worker stages and copies at once; the feature kernels are enqueued in front of this step's forward pass,
        #  beside its head: RS_PREFETCH_GATE=0 leaves them to the worker thread, i.e. to wherever the device happens to be)
        nxt = pending[0] = prefetch.submit(res[(i + 1) % NB]["sigs"], c["sr"], time_major=True, defer_features=gate_features,
                                            copy_after_current=gate_features and os.environ.get("RS_PREFETCH_COPY_GATE", "0") != "0")
        m.start_batch(None, True)
        x = f if r["T"] == c["Tmax"] or NB == 1 else f[:r["T"]]
        if gate_features:
            nxt.launch_features()
        m.step_on_batch(x, nf, r["labs"], compute_gradients=True, compute_error_rate=False)
        mean_loss, _, _ = m.end_batch(None, True, rnn_state_reset_ratio=1.0)
        e2e_i[0] = i + 1
        return float(mean_loss)

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(h.local)
    if rank == 0:
        sampler.start()
    ms, launches = h.timed(step_resident, args.steps, launch_count)
    clocks = sampler.stop() if rank == 0 else None
    # the product's default training step: with the reference's per-mini-batch prediction + error rate
    # (models/AcousticModel.py:641 fetches acc_error_rate_op in every run_step), decoder overlapped with backward
    state["err"] = True
    for _ in range(2):
        step_resident()
    ms_err, _ = h.timed(step_resident, args.steps)
    state["err"] = False
    for _ in range(max(3, args.warmup)):
        step_e2e()
    ms_e2e, _ = h.timed(step_e2e, args.steps)

    # ---- the roofline's region: the same K resident steps with CUDA events around every recurrent launch
    m.enable_timing()
    for _ in range(2):
        step_resident()
    ms_instr, _ = h.timed(step_resident, args.steps)
    T_last = res[(state["i"] - 1) % NB]["T"]
    # dram__bytes_read + write per recurrent launch from profiles/r02d_ncu_kernels.txt: backward 110.3 MB per 128-step
    # launch, forward 25.7 MB per 38-step launch = 65 MB per 96 steps; weighted by 24 / 33 launches per step
    roofline = recurrent_roofline(h, m, c, T_last, ms_instr / args.steps, (0, 1), 84.0e6 if name == "cfg2" else None) \
        if rank == 0 else None
    if roofline is not None:
        roofline["instrumented_ms_per_step"] = ms_instr / args.steps
        roofline["measured_over"] = ("a second timed region of the same %d resident steps, run right behind the first, with CUDA "
                                     "events on the launching streams around every recurrent launch; `value` is the region "
                                     "without them (the events themselves cost %.2f ms per step)"
                                     % (args.steps, (ms_instr - ms) / args.steps)) if not events_everywhere else \
            "the timed region of `value` (RS_BENCH_TIMED_EVENTS=1)"

    if os.environ.get("RS_BENCH_E2E_PHASES"):
        # diagnostic: the phases of the END-TO-END step and the gap between two steps (stderr)
        seq = []
        host_t = []
        for _ in range(16):
            m._phase_events = []
            t_h = time.perf_counter()
            step_e2e()
            host_t.append((time.perf_counter() - t_h) * 1e3)
            seq.append(m._phase_events)
        torch.cuda.synchronize()
        m._phase_events = None
        ph = {}
        for ev in seq[1:]:
            for (na, ea), (_, eb) in zip(ev[:-1], ev[1:]):
                ph.setdefault(na, []).append(ea.elapsed_time(eb))
        gaps = [a[-1][1].elapsed_time(b[0][1]) for a, b in zip(seq[1:-1], seq[2:])]
        sys.stderr.write("e2e phases (ms): %s; end -> next forward %.3f; step (forward mark to forward mark) %.3f\n" % (
            ", ".join("%s %.3f" % (k, float(np.median(v))) for k, v in ph.items()), float(np.median(gaps)),
            float(np.median([a[0][1].elapsed_time(b[0][1]) for a, b in zip(seq[1:-1], seq[2:])]))))
        sys.stderr.write("e2e per step (ms, forward mark to forward mark): %s\n" % " ".join(
            "%.2f" % a[0][1].elapsed_time(b[0][1]) for a, b in zip(seq[:-1], seq[1:])))
        sys.stderr.write("e2e backward per step (ms): %s\n" % " ".join(
            "%.2f" % dict((na, ea.elapsed_time(eb)) for (na, ea), (_, eb) in zip(ev[:-1], ev[1:]))["backward"] for ev in seq))
        sys.stderr.write("e2e host time of the call chain per step (ms): %s\n" % " ".join("%.2f" % t for t in host_t))
    # ---- per-family rooflines: phases of a step from CUDA events on the launching stream (every rank runs the steps:
    # the all-reduce is a collective), features alone on their stream
    fam_ms = {}
    for _ in range(3):
        m._phase_events = []
        step_resident()
        torch.cuda.synchronize()
        ev = m._phase_events
        for (na, ea), (_, eb) in zip(ev[:-1], ev[1:]):
            fam_ms.setdefault(na, []).append(ea.elapsed_time(eb))
    m._phase_events = None
    fam_ms = {k: float(np.median(v)) for k, v in fam_ms.items()}
    r0 = res[(state["i"]) % NB]
    fb_ms = event_ms(torch, lambda: ap.features_device(r0["pcm"], r0["off"], c["B"], r0["max_n"], c["sr"], time_major=True,
                                                       out=feats[0], nframes=nfr[0]), side, 5)
    h.barrier()

    utts = c["B"] * world * args.steps
    value = utts / (ms / 1e3)
    if rank != 0:
        return
    tc = bool(m.uses_tensor_cores)
    hbm_peak, hbm_src = h.hbm_peak()
    tensor_peak, _ = h.tensor_peak()
    Tm = float(np.mean([r["T"] for r in res]))
    n_samples = float(np.mean([r["pcm"].numel() for r in res]))
    P = m.n_params
    frames = float(np.mean([sum(min(c["Tmax"], ap.num_frames(len(s), c["sr"])) for s in r["sigs"]) for r in res]))

    def fam(ms_, bytes_=None, flops=None):
        d = {"ms_per_step": ms_}
        if bytes_ is not None:
            gbs = bytes_ / (ms_ / 1e3) / 1e9
            d.update({"bound": "hbm", "algorithmic_bytes": bytes_, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak})
        if flops is not None:
            tf = flops / (ms_ / 1e3) / 1e12
            d.update({"bound": "tensor", "algorithmic_flops": flops, "achieved": tf, "peak": tensor_peak, "unit": "TFLOP/s", "frac": tf / tensor_peak})
        return d
    fwd_fl = fwd_flops_per_frame(c) * Tm * c["B"]           # padded frames are computed too (dynamic_rnn semantics)
    roofline["families"] = {
        "fbank (fbank_logmel2_kernel + fbank_delta_kernel)": fam(fb_ms, 4.0 * n_samples + 4.0 * frames * c["F"]),
        "ctc (ctc_lse / ctc_lattice1 / ctc_grad kernels)": fam(fam_ms.get("ctc", float("nan")), 2.0 * Tm * c["B"] * c["C"] * 4),
        "clip_adam (sumsq_kernel + clip_adam_kernel)": fam(fam_ms.get("clip_adam", float("nan")), 32.0 * P),
        "lstm_stack_forward (gemm_tc + rec_ts_fwd)": fam(fam_ms.get("forward", float("nan")), flops=fwd_fl),
        "lstm_stack_backward (gemm_tc + rec_ts_bwd)": fam(fam_ms.get("backward", float("nan")), flops=2.0 * fwd_fl),
        "allreduce": {"ms_per_step": fam_ms.get("allreduce", 0.0), "bytes": 4.0 * P if world > 1 else 0.0},
        "note": "ms_per_step: CUDA events on the launching stream around each phase of a step (median of 3 steps after the "
                "timed region); fbank timed alone on its side stream (it overlaps the previous step otherwise); hbm peak = " + hbm_src,
    }
    line = {
        "metric": c["metric"], "value": value, "unit": c["unit"], "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16x3" if tc else "f32", "data": "synthetic",
        "config": {"workload": c["workload"], "global_batch": c["B"] * world, "parallelism": "dp%d" % world,
                   "arithmetic": arithmetic_note(tc),
                   "schedule": "time chunks (forward 96 steps: the layers of a wave side by side, chunk GEMMs as bursts between "
                               "waves, layer 0's input GEMMs of later chunks beside the first waves; backward %s, the dx "
                               "GEMMs between layers ahead of the weight-gradient GEMMs on the remaining SMs, the fill "
                               "patterns of the exchanged planes written at the head of the pass); weight planes packed once "
                               "per optimizer step, beside the input dense"
                               % ("%s steps: 2 recurrent launches in flight" % os.environ.get("RS_TC_CHUNK", "128")
                                  if 2 * (c["H"] // 16) + 32 <= 148 else "256 steps: 1 recurrent launch in flight (two of "
                                  "%d CTAs would leave the GEMMs fewer than 32 SMs)" % (c["H"] // 16)),
                   "l2": "per-step working set (activations > 2 GB, parameters x4) exceeds the 126 MB L2; no flush needed",
                   "train_tflop_per_step": 3.0 * fwd_fl * world / 1e12,
                   "audio_seconds_per_step": audio_seconds * world},
        "e2e": {"value": utts / (ms_e2e / 1e3), "unit": c["unit"], "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(np.mean([r["h2d"] for r in res])),
                "d2h_bytes_per_step": 16,
                "note": "one mini-batch is staged (pinned buffer), copied (ONE H2D) and featurised per timed step, for the step "
                        "after this one: BatchPrefetcher stages and copies on its thread and stream, the feature kernels are "
                        "enqueued in front of the forward pass; AcousticModel.step_on_batch / end_batch train on the current "
                        "mini-batch, and end_batch reads the mean loss back every step (16 bytes, copied behind the CTC "
                        "kernel: the call returns while the backward pass and the optimizer are still running)"},
        "with_error_rate": {"value": utts / (ms_err / 1e3), "unit": c["unit"], "ms_per_step": ms_err / args.steps,
                            "note": "the same step with the reference's per-mini-batch prediction (beam search, width 100) + "
                                    "edit distance (run_train_step's default), decoder on a side stream under the backward pass, "
                                    "mean loss / error rate read back every step"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(args, c, name, batches)
    print(json.dumps(line))


def cpu_baseline_sample(args, c, name, batches):
    """The oracle port timed on this box's host cores: one step of a bounded sample of the same workload."""
    from oracle import model, parallel
    params = model.init_params(c["L"], c["H"], c["F"], c["C"], seed=0)
    sigs, labs = batches[len(batches) // 2]
    nb = min(args.cpu_sample, len(sigs))
    sigs, labs = sigs[:nb], labs[:nb]
    parallel.pool()
    leg, times = cpu_pick_leg(sigs, labs, params, c, c["train"])
    parallel.close()
    return {"value": nb / times[leg], "unit": c["unit"], "cores": _host_threads(), "kind": "port",
            "sample": "1 %s on %d of the %d utterances of a batch, LSTM stack %s (numpy/OpenBLAS %.1f s, torch/MKL %.1f s), "
                      "features + CTC one process per utterance, %s, %d host threads"
                      % ("full training step" if c["train"] else "inference step", nb, c["B"],
                         "torch-CPU/MKL" if leg == "torch" else "numpy/OpenBLAS", times["numpy"], times["torch"],
                         cpu_model_name(), _host_threads())}


# --------------------------------------------------------------------------- cfg5: inference
def run_infer(args, name):
    h = Harness(args)
    torch = h.torch
    import rnn_speech_b200 as rs
    c = dict(CONFIGS[name])
    dev, rank, world = h.dev, h.rank, h.world
    rng = np.random.default_rng(1234 + rank)
    n = int(c["seconds"] * c["sr"])
    sigs = [(0.1 * rng.standard_normal(n)).astype(np.float32) for _ in range(c["B"])]
    ap = rs.AudioProcessor(c["Tmax"], "fbank", device=dev)
    m = rs.AcousticModel(c["L"], c["H"], c["B"], c["Tmax"], 600, c["F"], False, c["C"], device=dev, seed=0)
    m.create_forward_rnn()
    m.initialize(None)
    events_everywhere = bool(os.environ.get("RS_BENCH_TIMED_EVENTS"))    # (see run_train: the roofline's events get their
    if events_everywhere:                                                 #  own timed region)
        m.enable_timing()
    launch_count = rs._lib.raw("rs_launch_count")
    T = min(c["Tmax"], ap.num_frames(n, c["sr"]))
    pcm_d = torch.from_numpy(np.concatenate(sigs)).to(dev)
    off_d = torch.from_numpy(np.arange(c["B"] + 1, dtype=np.int64) * n).to(dev)

    def step_resident():
        m.infer_pcm_device(ap, pcm_d, off_d, c["B"], n, c["sr"])

    lat = []

    def step_e2e():
        t0 = time.perf_counter()
        ids, lens = m.infer_signals(ap, sigs, c["sr"])            # host PCM in, decoded ids on the host out
        lat.append(time.perf_counter() - t0)
        return ids, lens

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(h.local)
    if rank == 0:
        sampler.start()
    ms, launches = h.timed(step_resident, args.steps, launch_count)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(max(3, args.warmup)):
        step_e2e()
    del lat[:]
    ms_serial, _ = h.timed(step_e2e, args.steps)             # one batch at a time: the latency figures
    lat_ms = np.array(lat) * 1e3
    # throughput: the same call chain with the NEXT batch staged and copied (pinned buffer, ONE H2D) by the prefetcher's
    # thread and stream while this batch's kernels run; ids and lengths are read back every batch
    prefetch = rs.BatchPrefetcher(ap)
    pend = [prefetch.submit(sigs, c["sr"], defer_features=True)]

    def step_pipelined():
        tk = pend[0]
        pend[0] = prefetch.submit(sigs, c["sr"], defer_features=True)
        return m.infer_ticket(ap, tk, c["sr"])

    for _ in range(max(3, args.warmup)):
        step_pipelined()
    ms_e2e, _ = h.timed(step_pipelined, args.steps)
    pend[0].staged()
    prefetch.close()
    # ---- the roofline's region: the same K resident steps with CUDA events around every recurrent launch
    m.enable_timing()
    for _ in range(2):
        step_resident()
    ms_instr, _ = h.timed(step_resident, args.steps)
    roofline = recurrent_roofline(h, m, c, T, ms_instr / args.steps, (0,), None) if rank == 0 else None
    if roofline is not None:
        roofline["instrumented_ms_per_step"] = ms_instr / args.steps
        roofline["measured_over"] = ("a second timed region of the same %d resident steps with CUDA events on the launching "
                                     "streams around every recurrent launch; `value` is the region without them"
                                     % args.steps) if not events_everywhere else "the timed region of `value` (RS_BENCH_TIMED_EVENTS=1)"
    clips = c["B"] * world * args.steps
    if rank != 0:
        return
    tc = bool(m.uses_tensor_cores)
    ids, lens = step_e2e()
    line = {
        "metric": c["metric"], "value": clips / (ms / 1e3), "unit": c["unit"], "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16x3" if tc else "f32", "data": "synthetic",
        "config": {"workload": c["workload"], "global_batch": c["B"] * world, "parallelism": "dp%d" % world,
                   "arithmetic": arithmetic_note(tc), "batch_tiles": len(m._tiles) if m._tiles else 1,
                   "l2": "per-step activations (> 1 GB) exceed the 126 MB L2; no flush needed",
                   "fwd_tflop_per_step": fwd_flops_per_frame(c) * T * c["B"] * world / 1e12},
        "e2e": {"value": clips / (ms_e2e / 1e3), "unit": c["unit"], "ms_per_step": ms_e2e / args.steps,
                "one_batch_at_a_time": {"value": clips / (ms_serial / 1e3), "ms_per_step": ms_serial / args.steps},
                "latency_ms": {"p50": float(np.percentile(lat_ms, 50)), "p95": float(np.percentile(lat_ms, 95)),
                               "min": float(lat_ms.min()), "max": float(lat_ms.max()), "batches": int(len(lat_ms))},
                "h2d_bytes_per_step": int(4 * n * c["B"] + 8 * (c["B"] + 1)),
                "d2h_bytes_per_step": int(ids.nbytes + lens.nbytes),
                "note": "host PCM -> pinned staging -> ONE H2D copy -> feature kernels -> forward (batch tiles of 32) -> greedy "
                        "decode -> decoded ids and lengths copied to the host.  value: AcousticModel.infer_ticket, the next "
                        "batch staged and copied by BatchPrefetcher while this one computes; latency_ms and "
                        "one_batch_at_a_time: AcousticModel.infer_signals, wall clock per batch of 256 clips, nothing overlapped"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(args, c, name, [(sigs, [None] * len(sigs))])
    print(json.dumps(line))


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            try:
                from rnn_speech_b200 import dist as rsdist
                rsdist.close()
            except Exception:
                pass
            dist.destroy_process_group()
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=32, help="utterances in the cpu_baseline sample")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 16 if args.config == "cfg4" else 10
    if args.warmup is None:
        args.warmup = 8 if args.config == "cfg4" else 3
    if args.impl == "reference":
        run_reference(args)
    else:
        if CONFIGS[args.config]["train"]:
            run_train(args, args.config)
        else:
            run_infer(args, args.config)
        _shutdown()


if __name__ == "__main__":
    main()
