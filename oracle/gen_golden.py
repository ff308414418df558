"""Generates the committed golden vectors under tests/golden/ (run HERE, where
/root/reference exists; the GPU box only sees the .npz files).

    python -m oracle.gen_golden

  fbank_*.npz   inputs + outputs of the REFERENCE's own
                util/audioprocessor.py::_extract_fbank (via oracle/ref_shim.py),
                both delta modes -> pins oracle/features.py and the CUDA kernels.
  labels.npz    outputs of the reference's util/dataprocessor.py label codec.
  ctc_*.npz     oracle/ctc.py outputs (+ torch.nn.functional.ctc_loss where the
                labels do not contain the blank id).
  model_cfg1.npz  oracle/model.py forward/backward on BASELINE config 1 shapes.
"""
import os

import numpy as np

from . import ctc, features, model, ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synth_pcm(rng, n):
    """Seeded synthetic audio: 0.1*N(0,1) plus a few sinusoids so that the
    log-mel surface is not flat."""
    t = np.arange(n)
    x = 0.1 * rng.standard_normal(n)
    for f, a in ((220.0, 0.2), (1370.0, 0.1), (3100.0, 0.05)):
        x += a * np.sin(2 * np.pi * f * t / 16000.0 + rng.uniform(0, 6.28))
    return x.astype(np.float32)


def gen_fbank():
    ap = ref_shim.load_reference_audioprocessor()
    rng = np.random.default_rng(0)
    cases = {
        "cfg1_1s_16k": (16000, [16000, 16000], 3510),
        "ragged_16k": (16000, [4000, 7777, 1600 + 400, 12345], 3510),
        "crop_22k": (22050, [22050, 11025], 3510),
        "trunc_16k": (16000, [16000], 50),
        "silence_16k": (16000, [3200], 3510),
    }
    for name, (sr, lens, tmax) in cases.items():
        sigs = [synth_pcm(rng, n) for n in lens]
        if name.startswith("silence"):
            sigs = [np.zeros(n, np.float32) for n in lens]
        out = {"sr": sr, "tmax": tmax, "n": np.array(lens)}
        for mode in ("interp", "edge"):
            ref_shim.set_delta_mode(mode)
            proc = ap.AudioProcessor(tmax, "fbank")
            for i, s in enumerate(sigs):
                feat, length = proc.process_signal(s, sr)
                out["feat_%s_%d" % (mode, i)] = feat.astype(np.float64)
                out["len_%s_%d" % (mode, i)] = length
        for i, s in enumerate(sigs):
            out["sig_%d" % i] = s
        np.savez_compressed(os.path.join(OUT, "fbank_%s.npz" % name), **out)
        print("fbank", name, [out["feat_interp_%d" % i].shape for i in range(len(sigs))])


def gen_labels():
    dp = ref_shim.load_reference_dataprocessor().DataProcessor
    from importlib import util as _u
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, "models", "SpeechRecognizer.py")).read()
    ns = {}
    exec(src[src.index("ENGLISH_CHAR_MAP"):src.index("class SpeechRecognizer")], ns)
    cm = ns["ENGLISH_CHAR_MAP"]
    texts = ["it'll", "'d", "i will", "hello world", "the cat's doors' are yellow", "o'clock", "zz top bbq",
             "a", "", "mississippi's coffee", "we've been there", "don't", "i'm"]
    out = {"char_map": np.array(cm), "texts": np.array(texts)}
    for i, t in enumerate(texts):
        ids = dp.get_str_labels(cm, t)
        out["ids_%d" % i] = np.array(ids, np.int32)
        out["back_%d" % i] = np.array(dp.get_labels_str(cm, ids))
    np.savez_compressed(os.path.join(OUT, "labels.npz"), **out)
    print("labels", len(texts))


def gen_ctc():
    import torch
    rng = np.random.default_rng(11)
    T, B, C = 40, 5, 80
    logits = (2.0 * rng.standard_normal((T, B, C))).astype(np.float32)
    lens = np.array([40, 33, 40, 9, 0], np.int32)
    base = [rng.integers(1, 79, size=n).astype(np.int32) for n in (8, 12, 1, 10, 3)]
    base[0][3] = base[0][2]           # a repeated label
    for tag, labs in (("noeos", base), ("eos", [np.append(l, 79).astype(np.int32) for l in base])):
        out = {"logits": logits, "lens": lens, "B": B}
        for i, l in enumerate(labs):
            out["lab_%d" % i] = l
        for mode in ("source", "dest"):
            loss, grad = ctc.ctc_loss_and_grad(logits, labs, lens, beta_skip=mode)
            out["loss_" + mode] = loss
            out["grad_" + mode] = grad
        if tag == "noeos":
            ok = [i for i in range(B) if lens[i] > 0 and len(labs[i]) <= lens[i]]
            lt = torch.tensor(logits.astype(np.float64)[:, ok], requires_grad=True)
            tl = torch.nn.functional.ctc_loss(torch.log_softmax(lt, -1), torch.tensor(np.concatenate([labs[i] for i in ok]).astype(np.int64)),
                                              torch.tensor(lens[ok].astype(np.int64)), torch.tensor([len(labs[i]) for i in ok]),
                                              blank=79, reduction="none", zero_infinity=False)
            out["torch_items"] = np.array(ok)
            out["torch_loss"] = tl.detach().numpy()
        out["greedy_len"] = np.array([len(g) for g in ctc.greedy_decode(logits, lens)])
        for i, g in enumerate(ctc.greedy_decode(logits, lens)):
            out["greedy_%d" % i] = g
        np.savez_compressed(os.path.join(OUT, "ctc_%s.npz" % tag), **out)
        print("ctc", tag, out["loss_source"])


def gen_model():
    # BASELINE config 1: 1x128 LSTM, B=2, 1 s (T=98 fbank frames), 8 labels + EOS
    L, H, F, C, T, B = 1, 128, 120, 80, 98, 2
    rng = np.random.default_rng(5)
    params = model.init_params(L, H, F, C, seed=0)
    x = rng.standard_normal((T, B, F)).astype(np.float32)
    lens = np.array([98, 71], np.int32)
    labs = [np.append(rng.integers(1, 79, size=8), 79).astype(np.int32) for _ in range(B)]
    logits, state, cache = model.forward(params, x, lens, L, H)
    loss, dlogits = ctc.ctc_loss_and_grad(logits, labs, lens)
    grads = model.backward(params, cache, dlogits, L, H)
    out = {"x": x, "lens": lens, "flat_params": model.flatten(params, L, H, F, C).astype(np.float32),
           "logits": logits, "loss": loss, "flat_grads": model.flatten(grads, L, H, F, C),
           "state_c": state[0][0], "state_h": state[0][1], "dims": np.array([L, H, F, C, T, B])}
    for i, l in enumerate(labs):
        out["lab_%d" % i] = l
    np.savez_compressed(os.path.join(OUT, "model_cfg1.npz"), **out)
    print("model cfg1 loss", loss)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_fbank()
    gen_labels()
    gen_ctc()
    gen_model()
