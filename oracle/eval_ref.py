"""CPU oracle for the evaluation kernels (edit distance with error counts, posterior softmax).

TEST INFRASTRUCTURE ONLY (see oracle/ctc_ref.py): only tests/ may import it.

Restates ``compute_wer`` of the reference (utils/evaluation/edit_distance.py:53-126):
  :66-72   first row / column of the matrix = j / i
  :75-83   d[i][j] = d[i-1][j-1] when the tokens match, else 1 + min(substitution, insertion, deletion)
  :88-117  backtrace from (len(ref), len(hyp)), testing in this order: match on the diagonal ("C"), insertion
           (d[x][y] == d[x][y-1] + 1), substitution (d[x][y] == d[x-1][y-1] + 1), else deletion
Pinned by tests/golden/edit_distance_golden.npz, produced by the reference's own function.  One deviation, which
the fixture marks with ok = 0: the reference evaluates ``d[x-1]`` / ``d[x][y-1]`` with x == 0 or y == 0, i.e. with
python's index -1 (the LAST row / column), and raises IndexError / AssertionError on some inputs (its callers
swallow the exception and skip the utterance, examples/timit/s5/exp/metrics/phone.py:92-103).  Here the first
row backtraces as insertions and the first column as deletions, which is what the reference computes whenever
it does not raise.
"""

import numpy as np


def compute_wer(ref, hyp):
    """(distance, substitutions, insertions, deletions) for two token sequences."""
    R, H = len(ref), len(hyp)
    d = np.zeros((R + 1, H + 1), dtype=np.int64)
    d[0, :] = np.arange(H + 1)
    d[:, 0] = np.arange(R + 1)
    for i in range(1, R + 1):
        for j in range(1, H + 1):
            if ref[i - 1] == hyp[j - 1]:
                d[i, j] = d[i - 1, j - 1]
            else:
                d[i, j] = min(d[i - 1, j - 1], d[i, j - 1], d[i - 1, j]) + 1
    x, y, sub, ins, dele = R, H, 0, 0, 0
    while x > 0 or y > 0:
        if x == 0:
            ins += 1; y -= 1
        elif y == 0:
            dele += 1; x -= 1
        elif d[x, y] == d[x - 1, y - 1] and ref[x - 1] == hyp[y - 1]:
            x -= 1; y -= 1
        elif d[x, y] == d[x, y - 1] + 1:
            ins += 1; y -= 1
        elif d[x, y] == d[x - 1, y - 1] + 1:
            sub += 1; x -= 1; y -= 1
        else:
            dele += 1; x -= 1
    return int(d[R, H]), sub, ins, dele


def posteriors(logits, temperature=1.0):
    """softmax(logits / temperature) in float64 (models/pytorch_v3/ctc/ctc.py:486)."""
    z = np.asarray(logits, dtype=np.float64) / float(temperature)
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)
