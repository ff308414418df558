"""Diagnostics run on the GPU box: prints error magnitudes of each kernel family
against the oracle (more detail than the pass/fail of pytest) plus timing traces.
Test tooling (lives under tests/ because it uses the oracle as its checker); not
collected by pytest, run as `python tests/gpu_diag.py <mode>`."""
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
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            m.ctc_loss(lg, labs, ln)
        e1.record()
        torch.cuda.synchronize()
        print("CTC T=%d B=%d: %.3f ms per loss+grad call" % (T, B, e0.elapsed_time(e1) / 5))
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
            rs._lib.diag_call("rs_tc_selftest", Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, split,
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


def gemm_diag():
    import time
    for (M, N, K) in ((300, 200, 120), (1024, 3072, 768), (31936, 3072, 768)):
        rng = np.random.default_rng(M + N + K)
        A = rng.standard_normal((M, K)).astype(np.float32)
        B = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        bias = rng.standard_normal(N).astype(np.float32)
        Ad, Bd, bd = (torch.from_numpy(a).to(dev) for a in (A, B, bias))
        want = (Ad.double() @ Bd.double().T + bd.double()).cpu().numpy()
        scratch = torch.empty(2 * (M * K + N * K) * 2 + 64, dtype=torch.uint8, device=dev)
        for products in (1, 3):
            C = torch.full((M, N), float("nan"), dtype=torch.float32, device=dev)
            args = (Ad.data_ptr(), Bd.data_ptr(), bd.data_ptr(), C.data_ptr(), M, N, K, products, scratch.data_ptr(),
                    scratch.numel(), torch.cuda.current_stream().cuda_stream)
            rs._lib.diag_call("rs_gemm_tc_test", *args)
            torch.cuda.synchronize()
            got = C.cpu().numpy()
            e = np.abs(got - want)
            t0 = time.perf_counter()
            for _ in range(5):
                rs._lib.diag_call("rs_gemm_tc_test", *args)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
            print("gemm %dx%dx%d products=%d: rel err %.3e nan %d; %.3f ms incl. split (%.1f TFLOP/s useful)" % (
                M, N, K, products, e.max() / np.abs(want).max(), int(np.isnan(got).sum()), dt * 1e3,
                2.0 * M * N * K / dt / 1e12))
            if not (e.max() / np.abs(want).max() < 0.05):
                bad = np.argwhere(~(e < 0.05 * np.abs(want).max()))
                print("   first bad entries", bad[:5].tolist(), "of", len(bad))


def gemm_bench():
    """The tcgen05 GEMM alone (operands split once, CUDA events around `reps` launches) on the shapes of the
    pipelined schedule: tile width, grid size and CTA lifetime."""
    import ctypes
    shapes = [("dK chunk", 768, 3072, 4096, 1), ("dK full", 768, 3072, 31936, 1), ("dx chunk", 4096, 768, 3072, 0),
              ("gx chunk", 3072, 4096, 768, 0), ("gx full", 3072, 31936, 768, 0)]
    for name, M, N, K, acc in shapes:
        g = torch.Generator(device=dev); g.manual_seed(1)
        A = torch.randn((M, K), device=dev, generator=g)
        B = torch.randn((N, K), device=dev, generator=g)
        C = torch.zeros((M, N), device=dev)
        scratch = torch.empty(2 * (M * K + N * K) * 2 + 64, dtype=torch.uint8, device=dev)
        for products in (3, 1):
            row = []
            for bn in (128, 256):
                for ctas, tpc in ((0, 0), (52, 0), (36, 0), (0, 1)):
                    ms = ctypes.c_float()
                    rs._lib.diag_call("rs_gemm_tc_bench", A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, products, bn, ctas,
                                 tpc, acc, 10, scratch.data_ptr(), scratch.numel(), ctypes.byref(ms),
                                 torch.cuda.current_stream().cuda_stream)
                    n_sm = 148 if (ctas == 0) else ctas
                    tf = 2.0 * M * N * K * (3 if products == 3 else 1) / (ms.value * 1e-3) / 1e12
                    row.append("bn%d/%s: %.3f ms %4.0f TF (%.1f/SM)" % (bn, ("tpc1" if tpc else "g%d" % n_sm), ms.value, tf, tf / n_sm))
            print("%-9s %dx%dx%d products=%d\n    %s\n    %s" % (name, M, N, K, products, " | ".join(row[:4]), " | ".join(row[4:])))


def model_diag(which):
    def golden(name):
        return np.load(os.path.join(ROOT, "tests", "golden", name))
    if "cfg1" in which:
        g = golden("model_cfg1.npz")
        L, H, F, C, T, B = [int(v) for v in g["dims"]]
        m = rs.AcousticModel(L, H, B, T, 600, F, False, C, device=dev)
        m.create_training_rnn(1.0, 1.0, 1, 3e-4, 0.33)
        m.load_flat_params(g["flat_params"])
        print("cfg1: tensor cores:", m.uses_tensor_cores)
        x = torch.from_numpy(g["x"]).to(dev)
        lens = torch.from_numpy(g["lens"]).to(dev)
        logits = m.forward(x, lens, training=True)
        torch.cuda.synchronize()
        got = logits.cpu().numpy()
        print("cfg1 fwd: max |logit err| %.3e (nan %d); state c err %.3e h err %.3e" % (
            np.abs(got - g["logits"]).max(), int(np.isnan(got).sum()),
            np.abs(m.rnn_state[0, 0].cpu().numpy() - g["state_c"]).max(),
            np.abs(m.rnn_state[0, 1].cpu().numpy() - g["state_h"]).max()))
        labs = [g["lab_%d" % i] for i in range(B)]
        loss, grad = m.ctc_loss(logits, labs, lens)
        print("cfg1 loss rel err", np.abs(loss.cpu().numpy() - g["loss"]) / g["loss"])
        m.grads.zero_()
        m.backward(x, lens, grad)
        torch.cuda.synchronize()
        gg = m.grads.cpu().numpy()
        views = m.grad_views()
        want = g["flat_grads"]
        print("cfg1 bwd: max |grad err| / max|grad| = %.3e (nan %d)" % (np.abs(gg - want).max() / np.abs(want).max(), int(np.isnan(gg).sum())))
        off = 0
        for name, v in views.items():
            n = v.numel()
            e = np.abs(gg[off:off + n] - want[off:off + n]).max() / (np.abs(want[off:off + n]).max() + 1e-30)
            print("    %-55s rel err %.3e" % (name, e))
            off += n
    if "cfg2" in which:
        L, H, F, C, B, T = 3, 768, 120, 80, 32, 998
        rng = np.random.default_rng(0)
        p = model.init_params(L, H, F, C, seed=0)
        flat = model.flatten(p, L, H, F, C)
        x = rng.standard_normal((T, B, F)).astype(np.float32)
        lens = np.full(B, T, np.int32)
        lens[1::4] = rng.integers(T // 2, T, size=len(lens[1::4]))
        m = rs.AcousticModel(L, H, B, 1000, 600, F, False, C, device=dev)
        m.create_training_rnn(1.0, 1.0, 1, 3e-4, 0.33)
        m.load_flat_params(flat)
        m.enable_timing()
        dbg_f = torch.zeros((2 * T, 16), dtype=torch.int64, device=dev)
        dbg_b = torch.zeros((2 * T, 16), dtype=torch.int64, device=dev)
        rs._lib.call("rs_am_set_debug_timeline", m._handle, dbg_f.data_ptr(), dbg_b.data_ptr())
        print("cfg2: tensor cores:", m.uses_tensor_cores)
        xd, ld = torch.from_numpy(x).to(dev), torch.from_numpy(lens).to(dev)
        for it in range(2):
            m.rnn_state.zero_()
            logits = m.forward(xd, ld, training=True, keep_state=False)
            torch.cuda.synchronize()
        want, _, _ = model.forward(p, x, lens, L, H, keep_cache=False)
        got = logits.cpu().numpy()
        margin = ctc.top2_margin(want, lens)
        valid = np.arange(T)[:, None] < lens[None, :]
        diff = (got.argmax(-1) != want.argmax(-1)) & valid
        print("cfg2 fwd: max |logit err| %.3e (nan %d); argmax mismatches %d of %d frames; largest margin among mismatches %.2e" % (
            np.abs(got - want).max(), int(np.isnan(got).sum()), int(diff.sum()), int(valid.sum()),
            float(margin[diff].max()) if diff.any() else 0.0))
        labs = [np.append(rng.integers(1, 79, size=rng.integers(60, 121)), 79).astype(np.int32) for _ in range(B)]
        loss, grad = m.ctc_loss(logits, labs, ld)
        wl, _ = ctc.ctc_loss_and_grad(want, labs, lens, want_grad=False)
        print("cfg2 loss max rel err %.3e" % (np.abs(loss.cpu().numpy() - wl) / wl).max())
        m.grads.zero_()
        m.backward(xd, ld, grad)
        torch.cuda.synchronize()
        print("cfg2 grads finite:", bool(torch.isfinite(m.grads).all()), "norm %.4e" % float(m.grads.double().norm()))
        print("cfg2 recurrent kernel ms fwd/bwd:", m.recurrent_ms())
        names = ["P0 barrier seen", "P1 TMA issued", "M0 first stage landed", "M1 MMAs issued", "E0 accum ready",
                 "E1 math+stores done", "E2 fences+cta bar done"]
        for tag, d in (("fwd", dbg_f.cpu().numpy()), ("bwd", dbg_b.cpu().numpy())):
            steps = list(range(400, 406)) if tag == "fwd" else list(range(405, 399, -1))
            t0 = d[steps[0], 0]
            print("timeline %s (ns relative to step %d's P0; columns: %s)" % (tag, steps[0], "; ".join(names)))
            for sidx in steps:
                print("   step %d:" % sidx, " ".join("%7d" % (int(d[sidx, e]) - int(t0)) for e in range(7)),
                      " | signal done %7d, all flags seen %7d" % (int(d[sidx, 7]) - int(t0), int(d[sidx, 15]) - int(t0)))
                if tag == "fwd":
                    print("        group waits passed:", " ".join("%7d" % (int(d[sidx, e]) - int(t0)) for e in range(8, 11)),
                          " group MMAs issued:", " ".join("%7d" % (int(d[sidx, e]) - int(t0)) for e in range(12, 15)))
            per = np.diff(d[100:900, 0].astype(np.int64))
            print("   mean |P0(t+1)-P0(t)| = %.0f ns" % np.abs(per).mean())


def mma_bench():
    out = torch.zeros(2, dtype=torch.int64, device=dev)
    print("tcgen05.mma SS issue-rate (one CTA): cycles/MMA (issue), cycles/MMA (to completion)")
    for variant in (0, 1):
        for (M, N, nacc) in ((64, 32, 1), (64, 32, 4), (128, 32, 1), (128, 32, 4), (64, 16, 1), (64, 16, 4), (64, 64, 1),
                             (64, 64, 4), (128, 128, 1), (128, 256, 1)):
            count = 576
            for rep in range(2):
                out.zero_()
                rs._lib.diag_call("rs_tc_mma_bench", M, N, count, variant, nacc, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
            c = out.cpu().numpy()
            print("   %s M=%3d N=%3d nacc=%d: issue %.1f, complete %.1f cycles/MMA" % (
                "warp+elect " if variant else "one thread ", M, N, nacc, c[0] / count, c[1] / count))


def rec_diag():
    """cfg-2 recurrent kernels: per-variant kernel times, bitwise comparison against variant 0,
    and the in-kernel timelines (CTA 0 and the last CTA, all 16 stamps)."""
    L, H, F, C, B, T = 3, 768, 120, 80, 32, 998
    rng = np.random.default_rng(0)
    p = model.init_params(L, H, F, C, seed=0)
    flat = model.flatten(p, L, H, F, C)
    x = rng.standard_normal((T, B, F)).astype(np.float32)
    lens = np.full(B, T, np.int32)
    lens[1::4] = rng.integers(T // 2, T, size=len(lens[1::4]))
    if os.environ.get("RS_TS_GKB"):
        print("RS_TS_GKB =", os.environ["RS_TS_GKB"])
    os.environ.setdefault("RS_TC_CHUNK", "0")          # timelines are stamped by single-launch kernels only
    m = rs.AcousticModel(L, H, B, 1000, 600, F, False, C, device=dev)
    m.create_training_rnn(1.0, 1.0, 1, 3e-4, 0.33)
    m.load_flat_params(flat)
    m.enable_timing()
    dbg_f = torch.zeros((2 * T, 16), dtype=torch.int64, device=dev)
    dbg_b = torch.zeros((2 * T, 16), dtype=torch.int64, device=dev)
    rs._lib.call("rs_am_set_debug_timeline", m._handle, dbg_f.data_ptr(), dbg_b.data_ptr())
    xd, ld = torch.from_numpy(x).to(dev), torch.from_numpy(lens).to(dev)
    labs = [np.append(rng.integers(1, 79, size=rng.integers(60, 121)), 79).astype(np.int32) for _ in range(B)]
    ref = None
    variants = [int(v) for v in os.environ.get("RS_DIAG_VARIANTS", "0").split(",")]
    for variant in variants:
        os.environ["RS_TS_VARIANT"] = str(variant)
        for it in range(2):
            m.rnn_state.zero_()
            dbg_f.zero_(); dbg_b.zero_()
            logits = m.forward(xd, ld, training=True, keep_state=False)
            loss, grad = m.ctc_loss(logits, labs, ld)
            m.grads.zero_()
            m.backward(xd, ld, grad)
            torch.cuda.synchronize()
        got = (logits.clone(), m.grads.clone())
        if ref is None:
            ref = got
        print("variant %d: rec ms fwd/bwd %s ; logits bitwise == variant %d: %s ; grads bitwise: %s ; grads finite %s" % (
            variant, m.recurrent_ms(), variants[0], bool(torch.equal(got[0], ref[0])), bool(torch.equal(got[1], ref[1])),
            bool(torch.isfinite(got[1]).all())))
        for tag, d in (("fwd", dbg_f.cpu().numpy()), ("bwd", dbg_b.cpu().numpy())):
            steps = [400, 401] if tag == "fwd" else [405, 404]
            t0 = int(d[steps[0], 0])
            print("  timeline %s variant %d (ns relative to CTA0 step %d ev0); events 0..15" % (tag, variant, steps[0]))
            for cta, base in (("cta0", 0), ("ctaN", T)):
                for sidx in steps:
                    print("    %s step %d:" % (cta, sidx), " ".join(
                        "%6d" % (int(d[base + sidx, e]) - t0) if d[base + sidx, e] else "     ." for e in range(16)))
            per = np.diff(d[100:900, 0].astype(np.int64))
            print("    mean step period %.0f ns" % np.abs(per).mean())
            if tag == "fwd" and d[900, 15] and d[100, 15]:
                print("    SM clock during the kernel: %.0f MHz (clock64 vs globaltimer over steps 100..900)" % (
                    (int(d[900, 15]) - int(d[100, 15])) / max(1, int(d[900, 7]) - int(d[100, 7])) * 1e3))


def xchg_diag():
    """cfg-2 forward recurrent kernel with the validated exchange (rec_ts_fwd3_kernel): mean time between the stamped
    events of a step, per chain, for CTA 0 and the last CTA (steps 100..900 of a single launch over all T steps)."""
    L, H, F, C, B, T = 3, 768, 120, 80, 32, 998
    rng = np.random.default_rng(0)
    flat = model.flatten(model.init_params(L, H, F, C, seed=0), L, H, F, C)
    x = rng.standard_normal((T, B, F)).astype(np.float32)
    lens = np.full(B, T, np.int32)
    os.environ["RS_TC_CHUNK"] = "0"
    m = rs.AcousticModel(L, H, B, 1000, 600, F, False, C, device=dev)
    m.create_training_rnn(0.8, 0.5, 1, 3e-4, 0.33)
    m.load_flat_params(flat)
    m.enable_timing()
    dbg_f = torch.zeros((2 * T, 16), dtype=torch.int64, device=dev)
    dbg_b = torch.zeros((2 * T, 16), dtype=torch.int64, device=dev)
    rs._lib.call("rs_am_set_debug_timeline", m._handle, dbg_f.data_ptr(), dbg_b.data_ptr())
    xd, ld = torch.from_numpy(x).to(dev), torch.from_numpy(lens).to(dev)
    dl = torch.from_numpy((rng.standard_normal((T, B, C)) * 0.01).astype(np.float32)).to(dev)
    for it in range(2):
        m.rnn_state.zero_()
        dbg_f.zero_(); dbg_b.zero_()
        m.forward(xd, ld, training=True, keep_state=False)
        m.grads.zero_()
        m.backward(xd, ld, dl)
        torch.cuda.synchronize()
    print("rec ms fwd/bwd", m.recurrent_ms())
    db = dbg_b.cpu().numpy().astype(np.int64)
    bnames = ["counter seen", "fetch issued", "-", "MMAs done", "partials pushed", "partials received",
              "cell math done", "published"]
    two = bool(db[100, 8] > 10 ** 12)                  # rec_ts_bwd4_kernel stamps chain 1 in events 8..15
    for cta, base in (("cta0", 0), ("ctaN", T)):
        eb = db[base + 100:base + 900, :][::-1]          # backward runs t downwards: row = t - t0; the fetch for step t-1 is stamped at row t
        if not eb[:, 0].all():
            print("  bwd %s: no stamps (a stamped backward kernel did not run)" % cta)
            continue
        if two:
            for X in range(2):
                o = 8 * X
                print("  bwd %s chain %d: step period %.0f ns" % (cta, X, np.diff(eb[:, o]).mean()))
                dt = eb[:, o + 1] - eb[:, o]
                print("      %-26s %5.0f ns after 'counter seen' (p95 %5.0f)" % ("fetch issued", dt.mean(), np.percentile(dt, 95)))
                for k, nm in ((3, "MMAs done"), (4, "voted + partials pushed"), (5, "partials received"), (6, "cell math done"), (7, "published")):
                    dt = eb[1:, o + k] - eb[:-1, o]
                    print("      %-26s %5.0f ns after 'counter seen' (p95 %5.0f)" % (nm, dt.mean(), np.percentile(dt, 95)))
                nxt = eb[2:, o] - eb[2:, o + 7]
                print("      next counter seen +%5.0f ns after 'published' (p95 %5.0f)" % (nxt.mean(), np.percentile(nxt, 95)))
            print("  bwd %s: chain 1 'counter seen' - chain 0 'counter seen': mean %.0f ns" % (cta, (eb[:, 8] - eb[:, 0]).mean()))
            continue
        print("  bwd %s: step period %.0f ns" % (cta, np.diff(eb[:, 0]).mean()))
        dt = eb[:, 1] - eb[:, 0]
        print("      %-26s %5.0f ns after 'counter seen' (p95 %5.0f)" % ("fetch issued", dt.mean(), np.percentile(dt, 95)))
        for k, nm in ((3, "MMAs done"), (9, "accumulator read"), (10, "voted"), (4, "partials pushed"), (5, "partials received"),
                      (11, "partials summed"), (6, "cell math done"), (12, "barrier passed"), (7, "published")):
            dt = eb[1:, k] - eb[:-1, 0]                   # events of row t-1 belong to the fetch stamped at row t
            print("      %-26s %5.0f ns after 'counter seen' (p95 %5.0f)" % (nm, dt.mean(), np.percentile(dt, 95)))
        nxt = eb[2:, 0] - eb[2:, 7]                      # row t: 'published' of step t, then the counter for the fetch of tile t
        print("      next counter seen +%5.0f ns after 'published' (p95 %5.0f)" % (nxt.mean(), np.percentile(nxt, 95)))
    d = dbg_f.cpu().numpy().astype(np.int64)
    names = ["counter seen", "fetch issued", "tile landed", "MMAs done", "cell math done", "published", "end of step"]
    for cta, base in (("cta0", 0), ("ctaN", T)):
        for X in range(2):
            e = d[base + 100:base + 900, 8 * X:8 * X + 8]
            period = np.diff(e[:, 0]).mean()
            print("  %s chain %d: step period %.0f ns; attempts per step %.4f" % (cta, X, period, (e[-1, 7] - e[0, 7]) / (len(e) - 1)))
            for k in range(1, 7):
                dt = e[:, k] - e[:, 0]
                print("      %-26s %5.0f ns after 'counter seen' (p95 %5.0f)" % (names[k], dt.mean(), np.percentile(dt, 95)))
            nxt = e[1:, 0] - e[:-1, 5]
            print("      next counter seen +%5.0f ns after 'published' (p95 %5.0f)" % (nxt.mean(), np.percentile(nxt, 95)))
        print("  %s: chain 1 'counter seen' - chain 0 'counter seen': mean %.0f ns" % (cta, (d[base + 100:base + 900, 8] - d[base + 100:base + 900, 0]).mean()))
    print("  ctaN - cta0 'published' (chain 0): mean %.0f ns, std %.0f" % ((d[T + 100:T + 900, 5] - d[100:900, 5]).mean(), (d[T + 100:T + 900, 5] - d[100:900, 5]).std()))


def xchg2_diag():
    """The backward two-chain kernel's in-kernel timeline INSIDE the pipelined schedule (cfg-2): one launch -- layer and chunk
    from RS_TC_DBG_LC, default "1,3" -- records its stamps while the other layer's launch and the GEMMs run beside it."""
    L, H, F, C, B, T = 3, 768, 120, 80, 32, 998
    os.environ.setdefault("RS_TC_DBG_LC", "1,3")
    rng = np.random.default_rng(0)
    flat = model.flatten(model.init_params(L, H, F, C, seed=0), L, H, F, C)
    x = rng.standard_normal((T, B, F)).astype(np.float32)
    lens = np.full(B, T, np.int32)
    m = rs.AcousticModel(L, H, B, 1000, 600, F, False, C, device=dev)
    m.create_training_rnn(0.8, 0.5, 1, 3e-4, 0.33)
    m.load_flat_params(flat)
    m.enable_timing()
    n = 128
    dbg_f = torch.zeros((2 * T, 16), dtype=torch.int64, device=dev)
    dbg_b = torch.zeros((2 * T, 16), dtype=torch.int64, device=dev)
    rs._lib.call("rs_am_set_debug_timeline", m._handle, dbg_f.data_ptr(), dbg_b.data_ptr())
    xd, ld = torch.from_numpy(x).to(dev), torch.from_numpy(lens).to(dev)
    dl = torch.from_numpy((rng.standard_normal((T, B, C)) * 0.01).astype(np.float32)).to(dev)
    for it in range(2):
        m.rnn_state.zero_()
        dbg_b.zero_()
        m.forward(xd, ld, training=True, keep_state=False)
        m.grads.zero_()
        m.backward(xd, ld, dl)
        torch.cuda.synchronize()
    print("RS_TC_DBG_LC=%s RS_TC_WINDOW_BWD=%s: rec ms fwd/bwd %s" % (os.environ["RS_TC_DBG_LC"], os.environ.get("RS_TC_WINDOW_BWD", "default"), m.recurrent_ms()))
    db = dbg_b.cpu().numpy().astype(np.int64)
    for cta, base in (("cta0", 0), ("ctaN", n)):
        eb = db[base + 8:base + n - 8, :][::-1]
        if not eb[:, 0].all():
            print("  bwd %s: no stamps" % cta)
            continue
        for X in range(2):
            o = 8 * X
            print("  bwd %s chain %d: step period %.0f ns" % (cta, X, np.diff(eb[:, o]).mean()))
            dt = eb[:, o + 1] - eb[:, o]
            print("      %-26s %5.0f ns after 'counter seen' (p95 %5.0f)" % ("fetch issued", dt.mean(), np.percentile(dt, 95)))
            for k, nm in ((3, "MMAs done"), (4, "voted + partials pushed"), (5, "partials received"), (6, "cell math done"), (7, "published")):
                dt = eb[1:, o + k] - eb[:-1, o]
                print("      %-26s %5.0f ns after 'counter seen' (p95 %5.0f)" % (nm, dt.mean(), np.percentile(dt, 95)))
            nxt = eb[2:, o] - eb[2:, o + 7]
            print("      next counter seen +%5.0f ns after 'published' (p95 %5.0f)" % (nxt.mean(), np.percentile(nxt, 95)))


def trace_diag():
    """cfg-2 training step: where the recurrent launches of the pipelined schedule sit in time (CUDA events on the
    launching streams, ms after the top of the forward / backward call) and how long each phase of the step takes."""
    L, H, F, C, B, T = 3, 768, 120, 80, 32, 998
    if os.environ.get("RS_TRACE_CFG") == "4":            # BASELINE config 4's model and batch, an 11 s batch
        L, H, B, T = 5, 1024, 16, 1096
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((T, B, F)).astype(np.float32)).to(dev)
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    labs = [np.append(rng.integers(1, 79, size=rng.integers(60, 121)), 79).astype(np.int32) for _ in range(B)]
    m = rs.AcousticModel(L, H, B, max(T, 1000), 600, F, False, C, device=dev)
    m.create_training_rnn(0.8, 0.5, 1, 3e-4, 0.33)
    m.initialize(None)
    m.enable_timing()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    for it in range(3):
        m.grads.zero_()
        ev[0].record()
        logits = m.forward(x, lens, training=True, keep_state=False)
        ev[1].record()
        loss, grad = m.ctc_loss(logits, labs, lens)
        ev[2].record()
        m.backward(x, lens, grad)
        ev[3].record()
        m.apply_gradients()
        ev[4].record()
        torch.cuda.synchronize()
    print("RS_TC_CHUNK=%s RS_TC_WINDOW=%s: forward %.2f ms, ctc %.2f ms, backward %.2f ms, clip+adam %.2f ms" % (
        os.environ.get("RS_TC_CHUNK", "default"), os.environ.get("RS_TC_WINDOW", "default"),
        ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[3].elapsed_time(ev[4])))
    tr = m.recurrent_trace()
    for d, tag in ((0, "fwd"), (1, "bwd"), (2, "dK "), (3, "dx ")):
        for l in range(L):
            print("  %s layer %d: %s" % (tag, l, " ".join("[%.2f-%.2f]" % (a, b) for a, b in tr[d][l])))


def e2e_diag():
    """Host-side cost of the public-API training step at cfg-2: how long each call takes to RETURN (enqueue time)
    versus the device time of the step."""
    import time
    L, H, F, C, B, n = 3, 768, 120, 80, 32, 160000
    rng = np.random.default_rng(0)
    sigs = [(0.1 * rng.standard_normal(n)).astype(np.float32) for _ in range(B)]
    labs = [np.append(rng.integers(1, 79, size=rng.integers(60, 121)), 79).astype(np.int32) for _ in range(B)]
    ap = rs.AudioProcessor(1000, "fbank", device=dev)
    m = rs.AcousticModel(L, H, B, 1000, 600, F, False, C, device=dev)
    m.create_training_rnn(0.8, 0.5, 1, 3e-4, 0.33)
    m.initialize(None)
    for it in range(6):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        f, nf = ap.process_batch(sigs, 16000, time_major=True); t.append(time.perf_counter())
        m.start_batch(None, True); t.append(time.perf_counter())
        m.step_on_batch(f, nf, labs, compute_gradients=True, compute_error_rate=False); t.append(time.perf_counter())
        loss = m.end_batch(None, True, rnn_state_reset_ratio=1.0)[0]; t.append(time.perf_counter())
        if it >= 2:
            d = np.diff(t) * 1e3
            print("e2e host ms: process_batch %.2f  start_batch %.2f  step_on_batch (enqueue) %.2f  end_batch (sync + read) %.2f"
                  "  total %.2f" % (d[0], d[1], d[2], d[3], (t[-1] - t[0]) * 1e3))


def stress_diag():
    """Run-to-run determinism of the cfg-2 forward + backward under the pipelined schedule: every repetition must
    reproduce the first one bit for bit (logits) / to fp32 regrouping (gradients are summed chunk by chunk in a fixed
    order, so they are bitwise stable too).  Prints the repetitions that differ."""
    L, H, F, C, B, T = 3, 768, 120, 80, 32, 998
    reps = int(os.environ.get("RS_STRESS_REPS", "40"))
    rng = np.random.default_rng(0)
    p = model.init_params(L, H, F, C, seed=0)
    flat = model.flatten(p, L, H, F, C)
    x = torch.from_numpy(rng.standard_normal((T, B, F)).astype(np.float32)).to(dev)
    lens_np = np.full(B, T, np.int32)
    lens_np[1::4] = rng.integers(T // 2, T, size=len(lens_np[1::4]))
    lens = torch.from_numpy(lens_np).to(dev)
    dl = torch.from_numpy((rng.standard_normal((T, B, C)) * (np.arange(T)[:, None, None] < lens_np[None, :, None])).astype(np.float32)).to(dev)
    m = rs.AcousticModel(L, H, B, 1000, 600, F, False, C, device=dev)
    m.create_training_rnn(0.8, 0.5, 1, 3e-4, 0.33)
    m.load_flat_params(flat)
    # two different inputs alternate, so that data left over from the previous repetition is WRONG data: a read that
    # races ahead of its producer shows up as a difference from the first run of the same input
    xs = [x, torch.flip(x, dims=[0]) * 0.7]
    dls = [dl, torch.flip(dl, dims=[1]) * 1.3]
    refs = [None, None]
    bad = 0
    for it in range(reps):
        k = it & 1
        m.rnn_state.zero_()
        m._dropout_calls = 0
        logits = m.forward(xs[k], lens, training=True, keep_state=False)
        m.grads.zero_()
        m.backward(xs[k], lens, dls[k])
        junk = torch.randn(1 << 24, device=dev).sum()
        torch.cuda.synchronize()
        got = (logits.clone(), m.grads.clone())
        if refs[k] is None:
            refs[k] = got
            continue
        ref = refs[k]
        dlg = float((got[0] - ref[0]).abs().max())
        dgr = float((got[1] - ref[1]).abs().max() / ref[1].abs().max())
        if dlg != 0.0 or dgr != 0.0:
            bad += 1
            idx = (got[0] - ref[0]).abs().amax(dim=(1, 2)).nonzero().flatten()
            if bad <= 3:
                print("  rep %d: max |dlogit| %.3e (first differing step t=%s), grad rel diff %.3e" % (
                    it, dlg, int(idx[0]) if len(idx) else None, dgr))
    print("stress: %d of %d repetitions differ from the first run of their input (RS_TC_CHUNK=%s RS_TS_VARIANT=%s)" % (
        bad, reps - 2, os.environ.get("RS_TC_CHUNK", "default"), os.environ.get("RS_TS_VARIANT", "default")))


def bf16_round(x):
    return torch.from_numpy(np.asarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy().astype(np.float64)


def ts_diag():
    """TMEM-resident A operand: pin the TMEM layout of a bf16 A and time the M128 N64 TS MMA."""
    out = torch.zeros(2, dtype=torch.int64, device=dev)
    for K in (64, 256, 768):
        rng = np.random.default_rng(K)
        A = rng.standard_normal((64, K)).astype(np.float32)
        B = rng.standard_normal((32, K)).astype(np.float32)
        Ah = bf16_round(A); Al = bf16_round(A.astype(np.float64) - Ah)
        Bh = bf16_round(B); Bl = bf16_round(B.astype(np.float64) - Bh)
        want = np.zeros((128, 64))
        for q in range(4):
            rows = slice(16 * q, 16 * q + 16)
            want[32 * q:32 * q + 16, :32] = Ah[rows] @ Bh.T
            want[32 * q:32 * q + 16, 32:] = Ah[rows] @ Bl.T
            want[32 * q + 16:32 * q + 32, :32] = Al[rows] @ Bh.T
            want[32 * q + 16:32 * q + 32, 32:] = Al[rows] @ Bl.T
        Ad, Bd = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
        for variant in (0, 1):
            D = torch.full((128, 64), float("nan"), dtype=torch.float32, device=dev)
            rs._lib.diag_call("rs_tc_ts_selftest", Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), K, variant, 1, out.data_ptr(),
                         torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got = D.cpu().numpy().astype(np.float64)
            e = np.abs(got - want)
            quad = [e[np.ix_([32 * q + i + off for q in range(4) for i in range(16)], range(c0, c0 + 32))].max() / np.abs(want).max()
                    for off in (0, 16) for c0 in (0, 32)]
            full = got[[32 * q + i for q in range(4) for i in range(16)]]
            lo = got[[32 * q + 16 + i for q in range(4) for i in range(16)]]
            tot = full[:, :32] + full[:, 32:] + lo[:, :32]
            ref = A.astype(np.float64) @ B.astype(np.float64).T
            print("ts K=%d variant=%d: quadrant rel err hh %.2e hl %.2e lh %.2e ll %.2e ; x3 sum vs fp64 %.2e ; nan %d" % (
                K, variant, quad[0], quad[1], quad[2], quad[3], np.abs(tot - ref).max() / np.abs(ref).max(), int(np.isnan(got).sum())))
        for dcol in (448, 384, 416):
          for reps in (1, 8):
            rs._lib.diag_call("rs_tc_ts_selftest", Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), K, dcol << 8, reps, out.data_ptr(),
                         torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            c = out.cpu().numpy()
            n = reps * K // 16
            print("   ts timing K=%d reps=%d D@col %d: %d MMAs (M128 N64 K16, A in TMEM): issue %.1f, complete %.1f cycles/MMA" % (
                K, reps, dcol, n, c[0] / n, c[1] / n))


if __name__ == "__main__":
    which = sys.argv[1:] or ["ctc", "fbank"]
    if "ctc" in which:
        ctc_diag()
    if "xchg2" in which:
        xchg2_diag()
    if "fbank" in which:
        fbank_diag()
    if "tc" in which:
        tc_diag()
    if "gemm" in which:
        gemm_diag()
    if "rec" in which:
        rec_diag()
    if "ts" in which:
        ts_diag()
    if "gemmbench" in which:
        gemm_bench()
    if "stress" in which:
        stress_diag()
    if "e2e" in which:
        e2e_diag()
    if "xchg" in which:
        xchg_diag()
    if "trace" in which:
        trace_diag()
    if "mma" in which:
        mma_bench()
    if "cfg1" in which or "cfg2" in which:
        model_diag(which)
