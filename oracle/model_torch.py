"""Oracle (test infrastructure, NOT product code): the same restatement as oracle/model.py
(/root/reference/models/AcousticModel.py:189-317 forward, :386-401 gradients) written on torch-CPU
tensors, so that the GEMMs run on MKL with every host thread.  It exists for ONE purpose: the CPU
baseline legs of bench.py (BASELINE.md section 3 asked for a torch-CPU/MKL model leg next to the
numpy/OpenBLAS one; bench.py times both and keeps the faster).  tests/test_oracle_model.py checks it
against oracle/model.py; no product code imports it.

Same semantics as oracle/model.py: gate order i, j, f, o with forget_bias 1.0, kernel rows
[input half; recurrent half], dropout masks from oracle.model.dropout_mask, zero output / frozen state
past seq_len, gradient of the SUM of the per-item losses, no gradient into the carried state.
"""
import numpy as np
import torch

from . import model as np_model


def _t(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a)).to(dtype)


def forward(params, x, seq_len, L, H, keep_in=1.0, keep_out=1.0, seed=0, dtype=torch.float32):
    """x [T,B,F] numpy, seq_len [B].  Returns (logits [T,B,C] torch, cache)."""
    p = {k: _t(v, dtype) for k, v in params.items()}
    x = _t(x, dtype)
    keep_in, keep_out = float(np.float32(keep_in)), float(np.float32(keep_out))
    T, B, F = x.shape
    seq_len = np.asarray(seq_len)
    valid = torch.from_numpy(np.arange(T)[:, None] < seq_len[None, :])                 # [T,B] bool
    cur = torch.addmm(p["input_b"], x.reshape(T * B, F), p["input_w"]).reshape(T, B, H)
    cache = {"x": x, "valid": valid, "layers": [], "keep_in": keep_in, "keep_out": keep_out, "p": p}
    for l in range(L):
        K, b = p["kernel_%d" % l], p["bias_%d" % l]
        m_in = np_model.dropout_mask(seed, 2 * l, T, B, H, keep_in)
        m_out = np_model.dropout_mask(seed, 2 * l + 1, T, B, H, keep_out)
        xin = cur if m_in is None else cur * torch.from_numpy(m_in).to(dtype) / keep_in
        gx = torch.addmm(b, xin.reshape(T * B, H), K[:H]).reshape(T, B, 4 * H)
        Wh = K[H:].contiguous()
        c = torch.zeros((B, H), dtype=dtype)
        h = torch.zeros((B, H), dtype=dtype)
        out = torch.zeros((T, B, H), dtype=dtype)
        gates = torch.empty((T, B, 4 * H), dtype=dtype)
        cs = torch.empty((T + 1, B, H), dtype=dtype)
        hs = torch.empty((T + 1, B, H), dtype=dtype)
        cs[0], hs[0] = c, h
        for t in range(T):
            g = torch.addmm(gx[t], h, Wh)
            i = torch.sigmoid(g[:, :H]); j = torch.tanh(g[:, H:2 * H])
            f = torch.sigmoid(g[:, 2 * H:3 * H] + 1.0); o = torch.sigmoid(g[:, 3 * H:])
            cn = c * f + i * j
            hn = torch.tanh(cn) * o
            v = valid[t][:, None]
            out[t] = torch.where(v, hn, torch.zeros_like(hn))
            c = torch.where(v, cn, c)
            h = torch.where(v, hn, h)
            gates[t, :, :H] = i; gates[t, :, H:2 * H] = j
            gates[t, :, 2 * H:3 * H] = f; gates[t, :, 3 * H:] = o
            cs[t + 1], hs[t + 1] = c, h
        cache["layers"].append({"xin": xin, "m_in": m_in, "m_out": m_out, "gates": gates, "cs": cs, "hs": hs})
        cur = out if m_out is None else out * torch.from_numpy(m_out).to(dtype) / keep_out
    cache["top"] = cur
    C = p["output_w"].shape[1]
    logits = torch.addmm(p["output_b"], cur.reshape(T * B, H), p["output_w"]).reshape(T, B, C)
    return logits, cache


def backward(cache, dlogits, L, H):
    """Gradients of sum(loss) given dlogits [T,B,C]; returns a dict of numpy arrays."""
    p = cache["p"]
    dtype = p["input_w"].dtype
    dlogits = _t(dlogits, dtype)
    T, B, C = dlogits.shape
    valid = cache["valid"]
    keep_in, keep_out = cache["keep_in"], cache["keep_out"]
    g = {}
    top = cache["top"]
    dl2 = dlogits.reshape(T * B, C)
    g["output_w"] = top.reshape(T * B, H).t() @ dl2
    g["output_b"] = dl2.sum(0)
    dcur = (dl2 @ p["output_w"].t()).reshape(T, B, H)
    for l in range(L - 1, -1, -1):
        lay = cache["layers"][l]
        K = p["kernel_%d" % l]
        WhT = K[H:].t().contiguous()
        dout = dcur if lay["m_out"] is None else dcur * torch.from_numpy(lay["m_out"]).to(dtype) / keep_out
        gates, cs, hs = lay["gates"], lay["cs"], lay["hs"]
        dgates = torch.zeros((T, B, 4 * H), dtype=dtype)
        dh = torch.zeros((B, H), dtype=dtype)
        dc = torch.zeros((B, H), dtype=dtype)
        for t in range(T - 1, -1, -1):
            v = valid[t][:, None]
            i = gates[t, :, :H]; j = gates[t, :, H:2 * H]
            f = gates[t, :, 2 * H:3 * H]; o = gates[t, :, 3 * H:]
            tc = torch.tanh(cs[t + 1])
            dh_tot = dh + dout[t]
            do = dh_tot * tc * o * (1 - o)
            dc_tot = dc + dh_tot * o * (1 - tc * tc)
            di = dc_tot * j * i * (1 - i)
            dj = dc_tot * i * (1 - j * j)
            df = dc_tot * cs[t] * f * (1 - f)
            dg = torch.cat([di, dj, df, do], dim=1)
            dg = torch.where(v, dg, torch.zeros_like(dg))
            dgates[t] = dg
            dh = torch.where(v, dg @ WhT, dh)
            dc = torch.where(v, dc_tot * f, dc)
        dg2 = dgates.reshape(T * B, 4 * H)
        gK = torch.empty_like(K)
        gK[:H] = lay["xin"].reshape(T * B, H).t() @ dg2
        gK[H:] = hs[:T].reshape(T * B, H).t() @ dg2
        g["kernel_%d" % l] = gK
        g["bias_%d" % l] = dg2.sum(0)
        dxin = (dg2 @ K[:H].t()).reshape(T, B, H)
        dcur = dxin if lay["m_in"] is None else dxin * torch.from_numpy(lay["m_in"]).to(dtype) / keep_in
    x = cache["x"]
    F = x.shape[2]
    g["input_w"] = x.reshape(T * B, F).t() @ dcur.reshape(T * B, H)
    g["input_b"] = dcur.reshape(T * B, H).sum(0)
    return {k: v.numpy() for k, v in g.items()}
