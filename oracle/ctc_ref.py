"""CPU oracle for the CTC loss-and-gradient path and the greedy decoder.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker.

Parity status: **parity unpinned**.  The arithmetic of the reference's hot path
lives in a third-party dependency that is absent from ``/root/reference``
(``github.com/SeanNaren/warp-ctc`` + ``pytorch_binding``, cloned at HEAD with no
pinned version by ``tools/install_warpctc_pytorch.sh:7``), and the reference's
own tests assert no loss value, gradient or decoded sequence for this path
(``models/test/pytorch/ctc/test_ctc.py`` only trains until ``ler < 0.05``).  This
file therefore restates the published CTC algorithm (Graves et al. 2006) in
float64, following the in-tree description of the same algorithm in
``models/chainer/ctc/ctc_loss_from_chainer.py`` and the calling convention of
``models/pytorch_v3/ctc/ctc.py:30-66``; it is cross-checked in ``tests/`` against
``torch.nn.functional.ctc_loss`` (CPU, float64), against hand-derived
known-answer cases, and -- for the decoder -- against golden vectors generated
by importing the reference's own ``GreedyDecoder`` (``tests/golden/``).

Reference anchors (paths relative to /root/reference):
  * softmax over the vocabulary inside the op ........ chainer ctc :32-35, 269
  * blank-interleaved extended label sequence .......... chainer ctc :38-42
  * transition rule (stay / advance / skip unless equal
    symbols or blank) ................................... chainer ctc :185-199
  * forward and backward sweeps ......................... chainer ctc :234-264
  * cost = -log p(labels | acts) ........................ chainer ctc :283
  * gradient = softmax - posterior occupancy, masked by
    the input length ..................................... chainer ctc :288-304
  * call surface (acts[T,B,V], flat labels, lens; costs
    summed; grads stashed in forward) .................... pytorch_v3 ctc.py:30-66
  * greedy decoder ...................................... greedy_decoder.py:19-47
"""

from itertools import groupby

import numpy as np

NEG_INF = -np.inf


def log_softmax(x, axis=-1):
    """lp = x - logsumexp(x) in float64 (chainer ctc :32-35 computes the
    max-subtracted softmax and then takes its log, :269-270)."""
    x = np.asarray(x, dtype=np.float64)
    m = np.max(x, axis=axis, keepdims=True)
    return x - m - np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True))


def extended_labels(labels, blank=0):
    """[blank, l1, blank, l2, ..., blank] (chainer ctc :38-42)."""
    labels = np.asarray(labels, dtype=np.int64)
    ext = np.full(2 * len(labels) + 1, blank, dtype=np.int64)
    ext[1::2] = labels
    return ext


def count_repeats(labels):
    labels = np.asarray(labels)
    return int(np.sum(labels[1:] == labels[:-1])) if len(labels) > 1 else 0


def _alpha_beta(lp, ext, blank):
    """Log-space forward/backward variables for one utterance.

    lp:  [T, V] float64 log-probabilities (only the first T_b frames).
    ext: [S] extended label sequence.
    Both alpha and beta include the emission at their own frame (the
    warp-ctc / Graves convention; chainer's beta excludes it -- same
    posterior, different bookkeeping).
    """
    T = lp.shape[0]
    S = len(ext)
    em = lp[:, ext]                                   # [T, S] label-indexed gather
    # skip transition s-2 -> s allowed iff ext[s] != blank and ext[s] != ext[s-2]
    skip = np.zeros(S, dtype=bool)
    skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    # skip transition s -> s+2 (for beta) allowed iff ext[s+2] != blank and != ext[s]
    skip_b = np.zeros(S, dtype=bool)
    skip_b[:-2] = skip[2:]

    alpha = np.full((T, S), NEG_INF)
    alpha[0, 0] = em[0, 0]
    if S > 1:
        alpha[0, 1] = em[0, 1]
    for t in range(1, T):
        prev = alpha[t - 1]
        acc = prev.copy()
        acc[1:] = np.logaddexp(acc[1:], prev[:-1])
        via_skip = np.full(S, NEG_INF)
        via_skip[2:] = np.where(skip[2:], prev[:-2], NEG_INF)
        acc = np.logaddexp(acc, via_skip)
        alpha[t] = acc + em[t]

    beta = np.full((T, S), NEG_INF)
    beta[T - 1, S - 1] = em[T - 1, S - 1]
    if S > 1:
        beta[T - 1, S - 2] = em[T - 1, S - 2]
    for t in range(T - 2, -1, -1):
        nxt = beta[t + 1]
        acc = nxt.copy()
        acc[:-1] = np.logaddexp(acc[:-1], nxt[1:])
        via_skip = np.full(S, NEG_INF)
        via_skip[:-2] = np.where(skip_b[:-2], nxt[2:], NEG_INF)
        acc = np.logaddexp(acc, via_skip)
        beta[t] = acc + em[t]
    return alpha, beta, em


def ctc_cost_and_grad_single(acts_b, labels_b, blank=0):
    """One utterance: acts_b [T_b, V] unnormalised logits, labels_b [L_b].

    Returns (cost, grad[T_b, V]) in float64 with
      cost = -log p(labels | acts)
      grad = softmax(acts) - posterior occupancy per symbol
    (chainer ctc :283, :288-304).
    Infeasible alignments (L_b + repeats > T_b, or T_b == 0 with L_b > 0) give
    cost = +inf and an all-zero gradient (policy documented in DESIGN.md).
    """
    acts_b = np.asarray(acts_b, dtype=np.float64)
    T, V = acts_b.shape
    labels_b = np.asarray(labels_b, dtype=np.int64)
    L = len(labels_b)
    grad = np.zeros((T, V))
    if T == 0:
        return (0.0 if L == 0 else np.inf), grad
    if L + count_repeats(labels_b) > T:
        return np.inf, grad
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        lp = log_softmax(acts_b, axis=1)
        ext = extended_labels(labels_b, blank)
        S = len(ext)
        alpha, beta, em = _alpha_beta(lp, ext, blank)
        ll = alpha[T - 1, S - 1]
        if S > 1:
            ll = np.logaddexp(ll, alpha[T - 1, S - 2])
        if not np.isfinite(ll):
            return np.inf, grad
        # posterior occupancy of lattice state (t, s)
        post = np.exp(alpha + beta - em - ll)             # [T, S]
        post = np.where(np.isfinite(post), post, 0.0)
        occ = np.zeros((T, V))
        for s in range(S):                                # sum over states of equal symbol
            occ[:, ext[s]] += post[:, s]
        grad = np.exp(lp) - occ
    return -ll, grad


def ctc_cost_and_grad(acts, labels, act_lens, label_lens, blank=0):
    """Mini-batch oracle with the warp-ctc calling convention.

    acts [T, B, V] logits; labels flat [sum(label_lens)]; act_lens, label_lens [B].
    Returns (costs[B] float64, grads[T, B, V] float64); rows t >= act_lens[b]
    of the gradient are exactly zero (the reference wrapper pre-zeros grads,
    pytorch_v3 ctc.py:36).
    """
    acts = np.asarray(acts, dtype=np.float64)
    T, B, V = acts.shape
    labels = np.asarray(labels, dtype=np.int64)
    act_lens = np.asarray(act_lens, dtype=np.int64)
    label_lens = np.asarray(label_lens, dtype=np.int64)
    assert act_lens.shape == (B,) and label_lens.shape == (B,)
    assert int(label_lens.sum()) == len(labels)
    costs = np.zeros(B)
    grads = np.zeros((T, B, V))
    off = 0
    for b in range(B):
        Tb, Lb = int(act_lens[b]), int(label_lens[b])
        c, g = ctc_cost_and_grad_single(acts[:Tb, b, :], labels[off:off + Lb], blank)
        costs[b] = c
        grads[:Tb, b, :] = g
        off += Lb
    return costs, grads


def reduce_costs(costs, act_lens, size_average=False, length_average=False):
    """Scalar loss reductions.

    sum ............................ pytorch_v3 ctc.py:50
    size_average (mean over B) ..... pytorch_v3 ctc.py:46-48
    length_average (/ sum act_lens)  upstream warpctc_pytorch.CTCLoss [recollection]
    """
    costs = np.asarray(costs, dtype=np.float64)
    total = costs.sum()
    if length_average:
        return total / float(np.sum(act_lens))
    if size_average:
        return total / len(costs)
    return total


def greedy_decode(logits, x_lens, blank=0):
    """Best-path decoding (greedy_decoder.py:19-47): per-frame argmax (first
    index wins ties, numpy semantics), collapse repeats, then drop blanks.

    logits [B, T, V] (batch-major, raw logits -- pytorch_v3 ctc.py:436-437);
    returns a list of B int64 arrays (ragged).
    """
    logits = np.asarray(logits)
    hyps = []
    for b in range(logits.shape[0]):
        n = int(x_lens[b])
        best = np.argmax(logits[b, :n], axis=1) if n > 0 else np.zeros(0, dtype=np.int64)
        collapsed = [k for k, _ in groupby(best.tolist())]
        hyps.append(np.array([k for k in collapsed if k != blank], dtype=np.int64))
    return hyps


def concatenate_labels(ys, y_lens):
    """Flat int32 label vector from padded ys[B, Lmax] (pytorch_v3 ctc.py:532-549)."""
    ys = np.asarray(ys)
    return np.concatenate([ys[b, :int(y_lens[b])] for b in range(len(y_lens))]
                          or [np.zeros(0, dtype=np.int32)]).astype(np.int32)
