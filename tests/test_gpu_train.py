"""GPU parity of the update rule and of one complete training step (features ->
forward -> CTC -> backward -> clip -> Adam) against the oracle pipeline, plus the
step protocol of the reference (start_batch / run_step / end_batch)."""
import numpy as np
import pytest
import torch

from oracle import ctc, features, model, optim

pytestmark = pytest.mark.gpu


def test_clip_adam_matches_oracle(pkg, cuda):
    rng = np.random.default_rng(0)
    n = 100003
    theta = rng.standard_normal(n).astype(np.float32)
    m = pkg.AcousticModel(1, 8, 2, 4, 10, 8, False, 5, device=cuda)
    m.create_training_rnn(1.0, 1.0, 1, 3e-4, 0.33)
    lib = pkg._lib
    th = torch.from_numpy(theta.copy()).to(cuda)
    am, av = torch.zeros_like(th), torch.zeros_like(th)
    ss = torch.zeros(1, dtype=torch.float64, device=cuda)
    o_th, o_m, o_v = theta.astype(np.float64), np.zeros(n), np.zeros(n)
    stream = torch.cuda.current_stream().cuda_stream
    for step, scale in ((1, 5.0), (2, 1e-3), (3, 1.0)):                  # clipped, unclipped, clipped
        g = (rng.standard_normal(n) * scale).astype(np.float32)
        gd = torch.from_numpy(g).to(cuda)
        lib.call("rs_sumsq", gd.data_ptr(), n, ss.data_ptr(), stream)
        lib.call("rs_clip_adam_step", th.data_ptr(), gd.data_ptr(), am.data_ptr(), av.data_ptr(), n, ss.data_ptr(),
                 1.0, 3e-4, 0.9, 0.999, 1e-8, step, stream)
        o_th, o_m, o_v, norm = optim.clip_adam_step(o_th, g.astype(np.float64), o_m, o_v, step, 3e-4, 1.0)
        assert abs(float(ss.cpu()[0]) ** 0.5 - norm) < 1e-6 * norm
        np.testing.assert_allclose(th.cpu().numpy(), o_th, atol=2e-7)
        np.testing.assert_allclose(am.cpu().numpy(), o_m, atol=1e-7)
    with pytest.raises(ValueError):
        lib.call("rs_clip_adam_step", th.data_ptr(), gd.data_ptr(), am.data_ptr(), av.data_ptr(), n, ss.data_ptr(),
                 1.0, 3e-4, 0.9, 0.999, 1e-8, 0, stream)


def _synthetic_set(rng, n_items, seconds=1.0, sr=16000, n_labels=8):
    items = []
    for _ in range(n_items):
        sig = (0.1 * rng.standard_normal(int(seconds * sr))).astype(np.float32)
        lab = np.append(rng.integers(1, 79, size=n_labels), 79).astype(np.int32)
        items.append([(sig, sr), lab])
    return items


def test_full_training_step_matches_oracle_pipeline(pkg, cuda):
    """BASELINE config 1: 1x128 LSTM, batch 2 of 1 s audio, 8 labels + EOS, through
    the reference's step protocol; parameters after one step vs the oracle."""
    L, H, F, C, B, Tmax = 1, 128, 120, 80, 2, 100
    rng = np.random.default_rng(0)
    items = _synthetic_set(rng, 2)
    m = pkg.AcousticModel(L, H, B, Tmax, 600, F, False, C, device=cuda, seed=0)
    m.create_training_rnn(1.0, 1.0, 1, 3e-4, 0.33, use_iterator=True)
    m.initialize(None)
    theta0 = m.params.cpu().numpy().astype(np.float64)
    ds = pkg.AcousticModel.build_dataset(items, B, Tmax, 600, "fbank", pkg.ENGLISH_CHAR_MAP, device=cuda)
    m.add_datasets_input(ds, ds)
    mean_loss, err, step, empty = m.run_train_step(None, 1, 1.0)
    assert step == 1 and not empty and np.isfinite(mean_loss) and 0.0 <= err
    # oracle pipeline on the same inputs
    feats = []
    for (sig, sr), _ in items:
        f, n = features.fbank(sig, sr, Tmax)
        pad = np.zeros((Tmax, F))
        pad[:n] = f
        feats.append(pad)
    x = np.stack(feats, axis=1)
    lens = np.array([98, 98])
    p = model.unflatten(theta0, L, H, F, C)
    logits, _, cache = model.forward(p, x, lens, L, H)
    labs = [it[1] for it in items]
    loss, dlogits = ctc.ctc_loss_and_grad(logits, labs, lens)
    g = model.flatten(model.backward(p, cache, dlogits, L, H), L, H, F, C)
    assert abs(mean_loss - np.mean(loss / lens)) < 1e-3 * abs(np.mean(loss / lens))     # north-star gate
    np.testing.assert_allclose(m.grads.cpu().numpy(), g, atol=2e-3 * np.abs(g).max())
    th, _, _, _ = optim.clip_adam_step(theta0, g, np.zeros_like(g), np.zeros_like(g), 1, 3e-4, 1.0)
    upd_want, upd_got = th - theta0, m.params.cpu().numpy() - theta0
    # Adam's first step is lr * sign(g) wherever |g| >> eps: compare where the oracle gradient is not tiny
    # (entries with a near-zero gradient may legitimately take the other sign at fp32 resolution)
    # (the tensor-core path carries dh through a bf16 product: more small entries may flip)
    big = np.abs(g) / max(np.linalg.norm(g), 1.0) > (1e-2 if m.uses_tensor_cores else 1e-4)
    np.testing.assert_allclose(upd_got[big], upd_want[big], atol=3e-6)
    assert np.mean(np.abs(upd_got - upd_want) > 3e-6) < (5e-2 if m.uses_tensor_cores else 1e-3)
    # state reset ratio 1.0 -> zero state after the step (models/AcousticModel.py:681-682)
    assert float(m.rnn_state.abs().max()) == 0.0


def test_step_protocol_accumulates_minibatches_and_epoch_end(pkg, cuda):
    L, H, F, C, B, Tmax = 1, 32, 120, 80, 2, 100
    rng = np.random.default_rng(1)
    items = _synthetic_set(rng, 5)               # 3 mini-batches, the last one padded
    m = pkg.AcousticModel(L, H, B, Tmax, 600, F, False, C, device=cuda, seed=1)
    m.create_training_rnn(0.8, 0.5, 1, 1e-3, 0.33, use_iterator=True)
    m.initialize(None)
    ds = pkg.AcousticModel.build_dataset(items, B, Tmax, 600, "fbank", pkg.ENGLISH_CHAR_MAP, device=cuda)
    m.add_datasets_input(ds, ds)
    l1, e1, s1, empty1 = m.run_train_step(None, 2, 0.25)
    assert s1 == 1 and not empty1
    l2, e2, s2, empty2 = m.run_train_step(None, 2, 0.25)         # 1 mini-batch left, then OutOfRange
    assert s2 == 2 and empty2
    l3, e3, s3, empty3 = m.run_train_step(None, 2, 0.25)         # nothing left
    assert (l3, e3, s3, empty3) == (0.0, 0.0, 2, True)
    m.reset_train_iterator()
    losses = [m.run_train_step(None, 3, 1.0)[0] for _ in range(1)]
    ev_loss, ev_err, ev_step = m.run_evaluation(None)
    # the padded row has length 0: loss/len = 0/0 -> the reference's displayed mean is NaN too (:361-362)
    assert np.isnan(ev_loss) or np.isfinite(ev_loss)
    assert ev_step == m.global_step
    lr = m.get_learning_rate()
    m.learning_rate_decay_op()
    assert abs(m.get_learning_rate() - lr * 0.33) < 1e-12


def test_end_batch_early_read_equals_the_draining_read(pkg, cuda, monkeypatch):
    """end_batch copies the accumulators from a side stream behind the CTC kernel and returns while the backward pass is
    still running (RS_EARLY_READ, default) -- the losses, error rates, step counts and parameters of three accumulated
    train steps must be those of the path that drains the device first (RS_EARLY_READ=0), bit for bit; the dataset
    iterator's deferred feature kernels (launched by run_step in front of its forward pass) are part of both runs."""
    import torch
    L, H, F, C, B, Tmax = 2, 64, 120, 80, 4, 100
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("RS_EARLY_READ", mode)
        rng = np.random.default_rng(5)
        items = _synthetic_set(rng, 16)
        m = pkg.AcousticModel(L, H, B, Tmax, 600, F, False, C, device=cuda, seed=5)
        m.create_training_rnn(0.8, 0.5, 1, 1e-3, 0.33, use_iterator=True)
        m.initialize(None)
        ds = pkg.AcousticModel.build_dataset(items, B, Tmax, 600, "fbank", pkg.ENGLISH_CHAR_MAP, device=cuda)
        m.add_datasets_input(ds, ds)
        res = [m.run_train_step(None, 2, 1.0, compute_error_rate=(k == 1)) for k in range(2)]
        torch.cuda.synchronize()
        out[mode] = (res, m.params.clone(), m.global_step)
    (r1, p1, g1), (r0, p0, g0) = out["1"], out["0"]
    assert g1 == g0 == 2
    for a, b in zip(r1, r0):
        assert a[2] == b[2] and a[3] == b[3]
        assert np.array_equal(np.float32([a[0], a[1]]), np.float32([b[0], b[1]]), equal_nan=True), (a, b)
        assert np.isfinite(a[0])
    assert torch.equal(p1, p0)


def test_loss_decreases_on_a_fixed_batch(pkg, cuda):
    L, H, F, C, B, Tmax = 2, 64, 120, 80, 4, 100
    rng = np.random.default_rng(3)
    items = _synthetic_set(rng, 4)
    m = pkg.AcousticModel(L, H, B, Tmax, 600, F, False, C, device=cuda, seed=3)
    m.create_training_rnn(1.0, 1.0, 1, 3e-3, 0.33, use_iterator=True)
    m.initialize(None)
    ds = pkg.AcousticModel.build_dataset(items, B, Tmax, 600, "fbank", pkg.ENGLISH_CHAR_MAP, device=cuda)
    losses = []
    for _ in range(12):
        m.add_datasets_input(ds, ds)
        losses.append(m.run_train_step(None, 1, 1.0, compute_error_rate=False)[0])
    assert losses[-1] < 0.8 * losses[0], losses


def test_checkpoint_roundtrip(pkg, cuda, tmp_path):
    m = pkg.AcousticModel(2, 32, 2, 50, 60, 20, False, 80, device=cuda, seed=7)
    m.create_training_rnn(1.0, 1.0, 1, 3e-4, 0.33)
    m.initialize(None)
    m.global_step = 41
    m.save(None, str(tmp_path))
    n = pkg.AcousticModel(2, 32, 2, 50, 60, 20, False, 80, device=cuda, seed=8)
    n.create_training_rnn(1.0, 1.0, 1, 1e-2, 0.33)
    n.initialize(None)
    n.restore(None, str(tmp_path))
    assert torch.equal(m.params, n.params) and n.global_step == 41 and abs(n.get_learning_rate() - 3e-4) < 1e-9
    names = set(m.param_views())
    assert "rnn/multi_rnn_cell/cell_1/basic_lstm_cell/kernel" in names and "Input_Layer/input_w" in names


def test_restore_from_a_tensorflow_bundle(pkg, cuda, tmp_path):
    """AcousticModel.restore reads the reference's own checkpoint format (tensor bundle written here with the layout
    TF writes; the shipped checkpoint's index is parsed in tests/test_tf_checkpoint.py)."""
    from test_tf_checkpoint import _write_bundle
    L, H, F, C = 2, 64, 20, 30
    rng = np.random.default_rng(3)
    tensors = {"Input_Layer/input_w": rng.standard_normal((F, H)).astype(np.float32),
               "Input_Layer/input_b": rng.standard_normal(H).astype(np.float32),
               "Output_layer/output_w": rng.standard_normal((H, C)).astype(np.float32),
               "Output_layer/output_b": rng.standard_normal(C).astype(np.float32),
               "global_step": np.array(4321, np.int32), "learning_rate": np.array(2.5e-4, np.float32)}
    for l in range(L):
        tensors["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/kernel" % l] = rng.standard_normal((2 * H, 4 * H)).astype(np.float32)
        tensors["rnn/multi_rnn_cell/cell_%d/basic_lstm_cell/bias" % l] = rng.standard_normal(4 * H).astype(np.float32)
    _write_bundle(str(tmp_path / "acousticmodel.ckpt"), tensors)
    (tmp_path / "checkpoint").write_text('model_checkpoint_path: "acousticmodel.ckpt"\n')
    m = pkg.AcousticModel(L, H, 4, 50, 600, F, False, C, device=cuda)
    m.create_training_rnn(1.0, 1.0, 1, 3e-4, 0.33)
    m.initialize(None)
    m.restore(None, str(tmp_path))
    assert m.global_step == 4321 and abs(m.learning_rate_var - 2.5e-4) < 1e-9
    for k, v in m.param_views().items():
        np.testing.assert_array_equal(v.cpu().numpy(), tensors[k])


def test_save_in_the_reference_format_and_restore(pkg, cuda, tmp_path):
    """save(fmt="tf") writes the reference's tensor-bundle checkpoint (checksums through the library's rs_crc32c);
    every block verifies and restore() brings the parameters back bit for bit."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("tfck2", os.path.join(root, "rnn-speech_b200", "tf_checkpoint.py"))
    tfck = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tfck)
    L, H, F, C = 2, 128, 40, 30
    m = pkg.AcousticModel(L, H, 4, 50, 600, F, False, C, device=cuda, seed=3)
    m.create_training_rnn(1.0, 1.0, 1, 3e-4, 0.33)
    m.initialize(None)
    m.global_step = 77
    want = m.params.clone()
    path = m.save(None, str(tmp_path), fmt="tf")
    assert tfck.verify_table_checksums(path + ".index") == 3
    # payload checksum of one tensor as stored in its entry (accelerated CRC == pure-Python CRC)
    w = m.param_views()["Output_layer/output_w"].cpu().numpy().tobytes()
    assert tfck.struct.pack("<I", tfck.mask_crc(tfck.crc32c(w))) in open(path + ".index", "rb").read()
    m2 = pkg.AcousticModel(L, H, 4, 50, 600, F, False, C, device=cuda, seed=9)
    m2.create_training_rnn(1.0, 1.0, 1, 1e-3, 0.33)
    m2.initialize(None)
    m2.restore(None, str(tmp_path))
    assert m2.global_step == 77 and abs(m2.learning_rate_var - 3e-4) < 1e-9
    assert torch.equal(m2.params, want)
