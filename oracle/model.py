"""Oracle (test infrastructure, NOT product code): CPU restatement of the
reference's acoustic-model graph, forward and backward.

Follows /root/reference/models/AcousticModel.py:189-317 (_build_base_rnn) and
:319-407 (_add_training_on_rnn) plus the TensorFlow-1 cell/loop bodies those
lines call (TF is absent from /root/reference and this image -- restated from
its published source, PARITY UNPINNED upstream; the backward pass is
cross-checked against torch-CPU autograd in tests/test_oracle_model.py):

  * input dense  rnn_in[t] = x[t] @ w_i + b_i                     (:240-250)
  * L x DropoutWrapper(BasicLSTMCell(H)) in a MultiRNNCell         (:223-237)
      g = [x~, h] @ K + b,  K [2H,4H] rows = [input half; recurrent half]
      i, j, f, o = split(g, 4)                (TF gate order, forget_bias=1.0)
      c' = c*sigmoid(f + 1) + sigmoid(i)*tanh(j);   h' = tanh(c')*sigmoid(o)
      x~ = dropout(x, keep_in); the cell's output dropout(h', keep_out) goes
      up / out, the undropped (c', h') recur.
  * tf.nn.dynamic_rnn(sequence_length, initial_state, time_major)  (:277-278)
      for t >= len[b]: output row = 0, state row frozen.
  * output dense logits[t] = out[t] @ w_o + b_o                    (:301-309)
  * optional batch-norm over the batch axis, eps 1e-3, no scale/shift (:253-259)

Parameter order everywhere (and in the product's flat buffer):
  input_w [F,H], input_b [H], (kernel_l [2H,4H], bias_l [4H]) x L,
  output_w [H,C], output_b [C]
which are exactly the variables the reference checkpoints (:515-527).
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def dropout_mask(seed, stream_id, T, B, H, keep):
    """Bernoulli(keep) keep-mask [T,B,H] from a counter hash the CUDA kernels
    restate bit for bit (csrc/common.cuh: rs_dropout_keep).  TF's own RNG stream
    cannot be matched; any i.i.d. Bernoulli(keep) mask drawn fresh per time step
    is the same computation (tf.nn.dropout: x * floor(keep + U[0,1)) / keep).
    stream_id = 2*layer + (0: cell input, 1: cell output).  `keep` is taken at float32 precision, which is what
    the reference feeds (a float32 placeholder, models/AcousticModel.py:648-652) and what the C ABI receives: the
    threshold int(keep * 2^24) of 0.8 differs by one between float32 and float64."""
    keep = float(np.float32(keep))
    if keep >= 1.0:
        return None
    idx = np.arange(T * B * H, dtype=np.uint64)
    ctr = (np.uint64(stream_id) << np.uint64(40)) | idx
    z = _splitmix64(_splitmix64(np.uint64(seed)) ^ ctr)
    thr = np.uint64(int(keep * 16777216.0))
    return ((z >> np.uint64(40)) < thr).reshape(T, B, H)


def param_shapes(L, H, F, C):
    shapes = [("input_w", (F, H)), ("input_b", (H,))]
    for l in range(L):
        shapes += [("kernel_%d" % l, (2 * H, 4 * H)), ("bias_%d" % l, (4 * H,))]
    shapes += [("output_w", (H, C)), ("output_b", (C,))]
    return shapes


def param_count(L, H, F, C):
    return sum(int(np.prod(s)) for _, s in param_shapes(L, H, F, C))


def init_params(L, H, F, C, seed=0, dtype=np.float32):
    """Xavier/glorot-uniform weights, zero biases (models/AcousticModel.py:242-245,
    :303-306; BasicLSTMCell's default glorot_uniform kernel, zero bias)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in param_shapes(L, H, F, C):
        if len(shape) == 2:
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            out[name] = rng.uniform(-lim, lim, size=shape).astype(dtype)
        else:
            out[name] = np.zeros(shape, dtype)
    return out


def flatten(params, L, H, F, C):
    return np.concatenate([np.asarray(params[n]).reshape(-1) for n, _ in param_shapes(L, H, F, C)])


def unflatten(flat, L, H, F, C):
    out, off = {}, 0
    for n, s in param_shapes(L, H, F, C):
        k = int(np.prod(s))
        out[n] = np.asarray(flat[off:off + k]).reshape(s)
        off += k
    assert off == len(flat)
    return out


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def forward(params, x, seq_len, L, H, state=None, keep_in=1.0, keep_out=1.0, seed=0,
            normalization=False, dtype=np.float64, keep_cache=True):
    """x [T,B,F], seq_len [B].  state: list of L (c,h) pairs [B,H] or None (zeros).
    Returns logits [T,B,C], new_state, cache (for backward)."""
    x = np.asarray(x, dtype=dtype)
    T, B, F = x.shape
    seq_len = np.asarray(seq_len)
    keep_in, keep_out = float(np.float32(keep_in)), float(np.float32(keep_out))   # float32 placeholders in the reference
    p = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    valid = (np.arange(T)[:, None] < seq_len[None, :])                     # [T,B]
    cur = (x.reshape(T * B, F) @ p["input_w"] + p["input_b"]).reshape(T, B, H)
    cache = {"x": x, "valid": valid, "layers": [], "keep_in": keep_in, "keep_out": keep_out}
    if normalization:
        mu = cur.mean(axis=1, keepdims=True)
        var = cur.var(axis=1, keepdims=True)
        cache["bn"] = (cur, mu, var)
        cur = (cur - mu) / np.sqrt(var + 1e-3)
    new_state = []
    for l in range(L):
        K, b = p["kernel_%d" % l], p["bias_%d" % l]
        m_in = dropout_mask(seed, 2 * l, T, B, H, keep_in)
        m_out = dropout_mask(seed, 2 * l + 1, T, B, H, keep_out)
        xin = cur if m_in is None else cur * m_in / keep_in
        gx = (xin.reshape(T * B, H) @ K[:H] + b).reshape(T, B, 4 * H)
        Wh = K[H:]
        c = np.zeros((B, H), dtype) if state is None else np.asarray(state[l][0], dtype).copy()
        h = np.zeros((B, H), dtype) if state is None else np.asarray(state[l][1], dtype).copy()
        out = np.zeros((T, B, H), dtype)
        gates = np.empty((T, B, 4 * H), dtype) if keep_cache else None   # activated i,j,f,o
        cs = np.empty((T + 1, B, H), dtype) if keep_cache else None
        hs = np.empty((T + 1, B, H), dtype) if keep_cache else None
        if keep_cache:
            cs[0], hs[0] = c, h
        for t in range(T):
            g = gx[t] + h @ Wh
            i = _sigmoid(g[:, :H]); j = np.tanh(g[:, H:2 * H])
            f = _sigmoid(g[:, 2 * H:3 * H] + 1.0); o = _sigmoid(g[:, 3 * H:])
            cn = c * f + i * j
            hn = np.tanh(cn) * o
            v = valid[t][:, None]
            out[t] = np.where(v, hn, 0.0)
            c = np.where(v, cn, c)
            h = np.where(v, hn, h)
            if keep_cache:
                gates[t, :, :H] = i; gates[t, :, H:2 * H] = j
                gates[t, :, 2 * H:3 * H] = f; gates[t, :, 3 * H:] = o
                cs[t + 1], hs[t + 1] = c, h
        new_state.append((c, h))
        cache["layers"].append({"xin": xin, "m_in": m_in, "m_out": m_out, "gates": gates, "cs": cs, "hs": hs})
        cur = out if m_out is None else out * m_out / keep_out
    cache["top"] = cur
    C = p["output_w"].shape[1]
    logits = (cur.reshape(T * B, H) @ p["output_w"] + p["output_b"]).reshape(T, B, C)
    return logits, new_state, cache


def backward(params, cache, dlogits, L, H, dtype=np.float64):
    """Gradient of sum(loss) w.r.t. every parameter given dlogits [T,B,C]
    (models/AcousticModel.py:386-388: compute_gradients on the per-item loss
    VECTOR differentiates its sum).  No gradient flows into the persistent state
    variables (truncated BPTT).  Returns a dict of gradients."""
    p = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    dlogits = np.asarray(dlogits, dtype=dtype)
    T, B, C = dlogits.shape
    valid = cache["valid"]
    keep_in, keep_out = cache["keep_in"], cache["keep_out"]
    g = {}
    top = cache["top"]
    g["output_w"] = top.reshape(T * B, H).T @ dlogits.reshape(T * B, C)
    g["output_b"] = dlogits.reshape(T * B, C).sum(0)
    dcur = (dlogits.reshape(T * B, C) @ p["output_w"].T).reshape(T, B, H)
    for l in range(L - 1, -1, -1):
        lay = cache["layers"][l]
        K = p["kernel_%d" % l]
        Wh = K[H:]
        dout = dcur if lay["m_out"] is None else dcur * lay["m_out"] / keep_out
        gates, cs, hs = lay["gates"], lay["cs"], lay["hs"]
        dgates = np.zeros((T, B, 4 * H), dtype)
        dh = np.zeros((B, H), dtype)
        dc = np.zeros((B, H), dtype)
        for t in range(T - 1, -1, -1):
            v = valid[t][:, None]
            i = gates[t, :, :H]; j = gates[t, :, H:2 * H]
            f = gates[t, :, 2 * H:3 * H]; o = gates[t, :, 3 * H:]
            c_prev = cs[t]
            # for valid rows c_new == cs[t+1]; for frozen rows the values are unused
            tc = np.tanh(cs[t + 1])
            dh_tot = dh + dout[t]
            do = dh_tot * tc * o * (1 - o)
            dc_tot = dc + dh_tot * o * (1 - tc * tc)
            di = dc_tot * j * i * (1 - i)
            dj = dc_tot * i * (1 - j * j)
            df = dc_tot * c_prev * f * (1 - f)
            dg = np.concatenate([di, dj, df, do], axis=1)
            dg = np.where(v, dg, 0.0)
            dgates[t] = dg
            dh = np.where(v, dg @ Wh.T, dh)
            dc = np.where(v, dc_tot * f, dc)
        dg2 = dgates.reshape(T * B, 4 * H)
        gK = np.empty_like(K)
        gK[:H] = lay["xin"].reshape(T * B, H).T @ dg2
        gK[H:] = hs[:T].reshape(T * B, H).T @ dg2
        g["kernel_%d" % l] = gK
        g["bias_%d" % l] = dg2.sum(0)
        dxin = (dg2 @ K[:H].T).reshape(T, B, H)
        dcur = dxin if lay["m_in"] is None else dxin * lay["m_in"] / keep_in
    if "bn" in cache:
        raw, mu, var = cache["bn"]
        inv = 1.0 / np.sqrt(var + 1e-3)
        xhat = (raw - mu) * inv
        dcur = inv * (dcur - dcur.mean(axis=1, keepdims=True) - xhat * (dcur * xhat).mean(axis=1, keepdims=True))
    x = cache["x"]
    F = x.shape[2]
    g["input_w"] = x.reshape(T * B, F).T @ dcur.reshape(T * B, H)
    g["input_b"] = dcur.reshape(T * B, H).sum(0)
    return g
