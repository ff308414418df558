"""Oracle (test infrastructure, NOT product code): CPU restatement of
``tf.nn.ctc_loss(sparse_labels, logits, seq_len, ignore_longer_outputs_than_inputs=True)``
as called at /root/reference/models/AcousticModel.py:357, and of the greedy
decode used as the parity surface for prediction (BASELINE.json north_star).

TensorFlow 1.x is absent from /root/reference (requirements.txt:1-5, unpinned;
README.md:49 says >= 1.4) and cannot be installed here: the lattice rules below
restate tensorflow/core/util/ctc/ctc_loss_calculator.{h,cc} (TF 1.4) from its
published source.  PARITY UNPINNED upstream (the reference has no CTC test);
cross-checked in tests/test_oracle_ctc.py against torch.nn.functional.ctc_loss
for label sequences that do not contain the blank id.

TF rules restated:
  * time-major logits [T,B,C], blank = C-1, ctc_merge_repeated=True,
    preprocess_collapse_repeated=False, softmax inside the op;
  * l' = [b, l1, b, ..., lN, b]; labels may legally equal the blank id (the
    validity check is 0 <= l < C) -- the reference's EOS id 79 IS the blank
    (util/dataprocessor.py:174-175), so this is live on the hot path;
  * alpha: skip (u-2 -> u) iff l'[u] != blank and l'[u] != l'[u-2]   (tests the DESTINATION)
  * beta : skip (u -> u+2) iff l'[u] != blank and l'[u] != l'[u+2]   (tests the SOURCE)
  * only u in [max(0, U-2(T-t)), min(U, 2(t+1))) is updated, the rest stays log 0;
  * log p(z|x) = LSE_u(alpha[0,u] + beta[0,u])   (evaluated at t = 0);
  * loss = -log p;  grad[t,k] = y[t,k] - exp(LSE_{u: l'u=k}(alpha+beta)[t] - log p);
  * seq_len == 0 or len(labels) > seq_len  -> loss 0, grad 0 (item skipped);
  * log p == log 0 ("No valid path found")  -> loss +inf, grad = y.
``beta_skip`` selects the beta rule: "source" (TF as recalled) or "dest"
(the textbook-symmetric rule); one named switch, mirrored by the kernel.
"""
import numpy as np

NEG_INF = -np.inf


def _lse2(a, b):
    m = np.maximum(a, b)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = m + np.log(np.exp(a - m) + np.exp(b - m))
    return np.where(np.isneginf(m), NEG_INF, out)


def log_softmax(x):
    m = x.max(axis=-1, keepdims=True)
    e = x - m
    return e - np.log(np.exp(e).sum(axis=-1, keepdims=True))


def sparse_from_dense(dense_labels, num_labels, batch_size=None):
    """models/AcousticModel.py:151-156 / :174-178: the dense zero-padded label
    batch becomes a sparse tensor by DROPPING EVERY 0 (so a real label id 0 is
    lost), and (iterator path) rows that end up empty are filled with
    [num_labels-1].  Returns a list of 1-D int arrays."""
    dense_labels = np.asarray(dense_labels)
    rows = [row[row != 0].astype(np.int32) for row in dense_labels]
    if batch_size is not None:
        while len(rows) < batch_size:
            rows.append(np.zeros((0,), np.int32))
        rows = [r if len(r) else np.array([num_labels - 1], np.int32) for r in rows]
    return rows


def ctc_item(logp, labels, blank, beta_skip="source", want_grad=True):
    """One batch item.  logp: [T, C] log-softmax (float64), labels: 1-D ints.
    Returns (loss, grad[T,C] or None, alpha, beta)."""
    T, C = logp.shape
    N = len(labels)
    lp = np.full(2 * N + 1, blank, dtype=np.int64)
    lp[1::2] = labels
    U = len(lp)
    is_blank = lp == blank
    a_skip = np.zeros(U, bool)
    a_skip[2:] = (~is_blank[2:]) & (lp[2:] != lp[:-2])
    b_skip = np.zeros(U, bool)      # indexed by SOURCE u: may go u -> u+2
    if beta_skip == "source":
        b_skip[:-2] = (~is_blank[:-2]) & (lp[:-2] != lp[2:])
    elif beta_skip == "dest":
        b_skip[:-2] = (~is_blank[2:]) & (lp[:-2] != lp[2:])
    else:
        raise ValueError(beta_skip)
    uidx = np.arange(U)

    alpha = np.full((T, U), NEG_INF)
    alpha[0, 0] = logp[0, blank]
    if U > 1:          # (TF indexes row 1 unconditionally; N == 0 is undefined there)
        alpha[0, 1] = logp[0, lp[1]]
    for t in range(1, T):
        prev = alpha[t - 1]
        s = prev.copy()
        s[1:] = _lse2(s[1:], prev[:-1])
        s2 = np.full(U, NEG_INF)
        s2[2:] = prev[:-2]
        s = np.where(a_skip, _lse2(s, s2), s)
        lo, hi = max(0, U - 2 * (T - t)), min(U, 2 * (t + 1))
        live = (uidx >= lo) & (uidx < hi)
        alpha[t] = np.where(live, s + logp[t, lp], NEG_INF)

    beta = np.full((T, U), NEG_INF)
    beta[T - 1, max(U - 2, 0):] = 0.0
    for t in range(T - 2, -1, -1):
        nxt = beta[t + 1] + logp[t + 1, lp]
        s = nxt.copy()
        s[:-1] = _lse2(s[:-1], nxt[1:])
        s2 = np.full(U, NEG_INF)
        s2[:-2] = nxt[2:]
        s = np.where(b_skip, _lse2(s, s2), s)
        lo, hi = max(0, U - 2 * (T - t)), min(U, 2 * (t + 1))
        live = (uidx >= lo) & (uidx < hi)
        beta[t] = np.where(live, s, NEG_INF)

    ab0 = alpha[0] + beta[0]
    m = ab0.max()
    log_p = NEG_INF if np.isneginf(m) else m + np.log(np.exp(ab0 - m).sum())
    if not want_grad:
        return -log_p, None, alpha, beta
    y = np.exp(logp)
    if np.isneginf(log_p):
        return np.inf, y.copy(), alpha, beta
    ab = alpha + beta                                  # [T, U]
    prob = np.zeros((T, C))
    with np.errstate(under="ignore"):
        np.add.at(prob, (np.arange(T)[:, None].repeat(U, 1), np.broadcast_to(lp, (T, U))), np.exp(ab - log_p))
    return -log_p, y - prob, alpha, beta


def ctc_loss_and_grad(logits, labels_list, seq_len, blank=None, beta_skip="source", want_grad=True):
    """logits [T,B,C] (any float dtype; computed in float64), labels_list: list of B
    int arrays, seq_len [B].  Returns (loss[B] f64, grad[T,B,C] f64 or None)."""
    logits = np.asarray(logits, dtype=np.float64)
    T, B, C = logits.shape
    blank = C - 1 if blank is None else blank
    loss = np.zeros(B)
    grad = np.zeros_like(logits) if want_grad else None
    for b in range(B):
        L = int(seq_len[b])
        lab = np.asarray(labels_list[b], dtype=np.int64)
        if L < 0 or L > T:
            raise ValueError("sequence_length(%d) <= %d violated" % (b, T))
        if np.any(lab < 0) or np.any(lab >= C):
            raise ValueError("labels must be in [0, num_classes)")
        if L == 0 or len(lab) > L:
            continue
        logp = log_softmax(logits[:L, b, :])
        l, g, _, _ = ctc_item(logp, lab, blank, beta_skip, want_grad)
        loss[b] = l
        if want_grad:
            grad[:L, b, :] = g
    return loss, grad


def greedy_decode(logits, seq_len, blank=None):
    """tf.nn.ctc_greedy_decoder(merge_repeated=True) restated: per item argmax
    over classes for t < len (first index on ties), collapse repeats, drop blank.
    Returns a list of B int32 arrays."""
    logits = np.asarray(logits)
    T, B, C = logits.shape
    blank = C - 1 if blank is None else blank
    out = []
    for b in range(B):
        L = int(seq_len[b])
        path = logits[:L, b, :].argmax(axis=-1)
        keep = np.ones(L, bool)
        keep[1:] = path[1:] != path[:-1]
        seq = path[keep]
        out.append(seq[seq != blank].astype(np.int32))
    return out


def top2_margin(logits, seq_len):
    """Per-frame gap between the two largest logits (inf for t >= len): used by
    the parity tests' margin rule for argmax comparisons."""
    logits = np.asarray(logits, dtype=np.float64)
    T, B, C = logits.shape
    part = np.partition(logits, C - 2, axis=-1)
    gap = part[..., C - 1] - part[..., C - 2]
    mask = np.arange(T)[:, None] < np.asarray(seq_len)[None, :]
    return np.where(mask, gap, np.inf)


# --------------------------------------------------------------------------------------
# CTC beam search: tf.nn.ctc_beam_search_decoder(logits, seq_len, beam_width=100, top_paths=1,
# merge_repeated=True) as called at /root/reference/models/AcousticModel.py:312 (defaults).
# Restates tensorflow/core/util/ctc/ctc_beam_search.h (CTCBeamSearchDecoder::Step / TopPaths,
# BeamEntry::LabelSeq) with the default scorer (all expansion scores 0) [3P, TensorFlow is absent
# from the reference tree and not installable here: parity unpinned upstream].
#   * per step the class scores are the logits minus their maximum (TF >= 1.12 also subtracts the
#     log-sum-exp; every beam holds exactly one emission per step, so either way adds the same
#     constant to all beams: the ranking and the decoded path do not depend on it -- `normalize`);
#   * each beam entry keeps log P(blank-ended), log P(label-ended) and their sum;
#   * an entry's label-ended mass also receives its parent's mass when the parent is still in the
#     beam ("Active"); an extension that already exists as an active entry is not created twice;
#   * the beam is the `beam_width` best of (updated entries + new extensions) by total, incumbents
#     win ties (TopN is filled with the updated entries first, a new leaf must be strictly better
#     than the bottom);
#   * top path = best total; merge_repeated=True collapses equal neighbours of the OUTPUT labels
#     (LabelSeq), which is TF's documented quirk.
# float32 arithmetic like TF; tie order among exactly equal totals is unspecified in TF (heap) and
# fixed here as: updated entries before new extensions, then lower beam slot, then lower label.
class _BeamEntry(object):
    __slots__ = ("parent", "label", "children", "oldp", "newp")

    def __init__(self, parent, label):
        self.parent, self.label, self.children = parent, label, None
        ninf = np.float32(-np.inf)
        self.oldp = [ninf, ninf, ninf]      # total, blank, label
        self.newp = [ninf, ninf, ninf]

    def active(self):
        return self.newp[0] != -np.inf


def _lse32(a, b):
    a, b = np.float32(a), np.float32(b)
    if a == -np.inf:
        return b
    if b == -np.inf:
        return a
    m = max(a, b)
    return np.float32(m + np.log1p(np.exp(np.float32(-abs(a - b)), dtype=np.float32), dtype=np.float32))


def beam_search_decode(logits, seq_len, beam_width=100, merge_repeated=True, blank=None, normalize=True):
    """Returns (list of B int32 arrays: the top path, float32 [B]: its log score)."""
    logits = np.asarray(logits, dtype=np.float32)
    T, B, C = logits.shape
    blank = C - 1 if blank is None else blank
    assert blank == C - 1, "TF's decoder fixes the blank at num_classes - 1"
    outs, scores = [], np.zeros(B, np.float32)
    for b in range(B):
        root = _BeamEntry(None, -1)
        root.newp = [np.float32(0.0), np.float32(0.0), np.float32(-np.inf)]
        leaves = [root]
        for t in range(int(seq_len[b])):
            raw = logits[t, b]
            inp = raw - raw.max()
            if normalize:
                inp = inp - np.log(np.exp(inp, dtype=np.float32).sum(dtype=np.float32), dtype=np.float32)
            inp = inp.astype(np.float32)
            branches = leaves           # kept sorted by newp.total, descending
            for e in branches:
                e.oldp = list(e.newp)
            for e in branches:
                if e.parent is not None:
                    if e.parent.active():
                        prev = e.parent.oldp[1] if e.label == e.parent.label else e.parent.oldp[0]
                        e.newp[2] = _lse32(e.newp[2], prev)
                    e.newp[2] = np.float32(e.newp[2] + inp[e.label])
                e.newp[1] = np.float32(e.oldp[0] + inp[blank])
                e.newp[0] = _lse32(e.newp[1], e.newp[2])
            cands = [(e.newp[0], 0, i, -1, e) for i, e in enumerate(branches)]
            for i, e in enumerate(branches):
                if e.children is None:
                    e.children = {}
                for c in range(C - 1):
                    ch = e.children.get(c)
                    if ch is not None and ch.active():
                        continue
                    prev = e.oldp[1] if c == e.label else e.oldp[0]
                    tot = np.float32(inp[c] + prev)
                    if tot > -np.inf:
                        cands.append((tot, 1, i, c, e))
            # beam_width best by total; incumbents first on ties, then slot, then label
            cands.sort(key=lambda x: (-float(x[0]), x[1], x[2], x[3]))
            keep = cands[:beam_width]
            new_leaves, kept_ids = [], set()
            for tot, kind, i, c, e in keep:
                if kind == 0:
                    new_leaves.append(e)
                    kept_ids.add(id(e))
                else:
                    ch = e.children.get(c)
                    if ch is None:
                        ch = e.children[c] = _BeamEntry(e, c)
                    ch.newp = [tot, np.float32(-np.inf), tot]
                    new_leaves.append(ch)
                    kept_ids.add(id(ch))
            for e in branches:
                if id(e) not in kept_ids:           # evicted: "bottom->newp.Reset()"
                    e.newp = [np.float32(-np.inf)] * 3
            leaves = new_leaves
        best = leaves[0]
        scores[b] = best.newp[0]
        seq, prev_label, e = [], -1, best
        while e.parent is not None:
            if not merge_repeated or e.label != prev_label:
                seq.append(e.label)
            prev_label = e.label
            e = e.parent
        outs.append(np.array(seq[::-1], dtype=np.int32))
    return outs, scores
