"""CPU oracle for the CTC prefix beam search (TEST INFRASTRUCTURE ONLY; see oracle/ctc_ref.py).

A restatement of models/pytorch_v3/ctc/decoders/beam_search_decoder.py:33-124 (no language model) in float64:
  :57      the beam starts with the empty prefix, p_blank = log 1, p_non_blank = log 0
  :66-81   a blank keeps the prefix and feeds its p_blank
  :86-101  a non-blank symbol c extends the prefix (only from p_blank when c repeats the last symbol) ...
  :105-109 ... and a repeated symbol also keeps the prefix, feeding its p_non_blank
  :113-116 sort by logaddexp(p_blank, p_non_blank), descending, stable; keep beam_width
  :118     the first prefix of the last beam is the hypothesis
Prefixes are tuples in an insertion-ordered dict, exactly as the reference keeps them, so the tie order of its
stable sort is reproduced.  Pinned by tests/golden/beam_golden.npz (outputs of the reference's own class).
"""

import numpy as np

LOG_0 = -np.inf


def beam_search(log_probs, n_frames, beam_width, blank=0):
    """log_probs [T, V]; returns (hypothesis as a list of ints, its log-probability)."""
    lp = np.asarray(log_probs, dtype=np.float64)
    V = lp.shape[1]
    beam = [((), (0.0, LOG_0))]
    with np.errstate(all="ignore"):
        for t in range(int(n_frames)):
            nxt = {}
            for c in range(V):
                p_t = lp[t, c]
                for prefix, (p_b, p_nb) in beam:
                    if c == blank:
                        n_b, n_nb = nxt.get(prefix, (LOG_0, LOG_0))
                        nxt[prefix] = (np.logaddexp(n_b, np.logaddexp(p_b + p_t, p_nb + p_t)), n_nb)
                        continue
                    end = prefix[-1] if prefix else None
                    new_prefix = prefix + (c,)
                    n_b, n_nb = nxt.get(new_prefix, (LOG_0, LOG_0))
                    if c != end:
                        n_nb = np.logaddexp(n_nb, np.logaddexp(p_b + p_t, p_nb + p_t))
                    else:
                        n_nb = np.logaddexp(n_nb, p_b + p_t)
                    nxt[new_prefix] = (n_b, n_nb)
                    if c == end:
                        n_b, n_nb = nxt.get(prefix, (LOG_0, LOG_0))
                        nxt[prefix] = (n_b, np.logaddexp(n_nb, p_nb + p_t))
            beam = sorted(nxt.items(), key=lambda x: np.logaddexp(*x[1]), reverse=True)[:beam_width]
    return list(beam[0][0]), float(np.logaddexp(*beam[0][1]))
