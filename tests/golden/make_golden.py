"""Generates the committed golden fixtures (run in the build container, where /root/reference exists).

  greedy_golden.npz : inputs and outputs of the REFERENCE's own GreedyDecoder
                      (models/pytorch_v3/ctc/decoders/greedy_decoder.py), imported from
                      /root/reference and called per utterance (B=1 slices: its final
                      np.array(best_hyps) raises on ragged results under numpy >= 1.24).
  ctc_golden.npz    : CTC costs/gradients from torch.nn.functional.ctc_loss (CPU, float64) -- an
                      implementation independent of oracle/ -- on small seeded cases.  The
                      reference's own loss (warp-ctc) cannot be run: it is not vendored.

Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def greedy_cases():
    sys.path.insert(0, "/root/reference")
    from models.pytorch_v3.ctc.decoders.greedy_decoder import GreedyDecoder  # the reference itself
    dec = GreedyDecoder(blank_index=0)
    rng = np.random.RandomState(1623)
    out = {}
    shapes = [(4, 37, 6), (3, 64, 30), (2, 50, 62), (2, 20, 700), (5, 9, 3)]
    for i, (B, T, V) in enumerate(shapes):
        logits = rng.randn(B, T, V).astype(np.float32)
        if i == 0:                       # ties, all-blank frames, leading/trailing blanks, repeats
            logits[0, :, :] = 0.0        # all ties -> argmax 0 (blank) everywhere
            logits[1, :5, 0] = 10.0
            logits[1, -5:, 0] = 10.0
            logits[2, 10:20, 3] = 9.0    # long run of one symbol
            logits[2, 14, 0] = 20.0      # ... split by a blank -> symbol emitted twice
            logits[3, ::2, 2] = 7.0
            logits[3, 1::2, 2] = 7.0
            logits[3, 5, 4] = 7.0        # exact tie at frame 5 between 2 and 4 -> first index (2)
        x_lens = rng.randint(T // 2, T + 1, size=B)
        x_lens[0] = T
        if i == 4:
            x_lens[1] = 0                # empty utterance
        hyps = []
        for b in range(B):
            h = dec(logits[b:b + 1], x_lens[b:b + 1])
            hyps.append(np.asarray(h[0], dtype=np.int64).reshape(-1))
        out["logits_%d" % i] = logits
        out["x_lens_%d" % i] = x_lens.astype(np.int32)
        out["hyp_lens_%d" % i] = np.array([len(h) for h in hyps], dtype=np.int32)
        out["hyp_flat_%d" % i] = np.concatenate(hyps) if hyps else np.zeros(0, np.int64)
    out["n_cases"] = np.array(len(shapes))
    np.savez_compressed(os.path.join(HERE, "greedy_golden.npz"), **out)


def ctc_cases():
    rng = np.random.RandomState(1623)
    out = {}
    shapes = [(1, 1, 2, 1), (3, 12, 5, 4), (4, 40, 30, 12), (2, 25, 62, 10), (3, 30, 200, 8), (2, 60, 7, 30)]
    for i, (B, T, V, Lmax) in enumerate(shapes):
        acts = (rng.randn(T, B, V) * (3.0 if i == 3 else 1.0)).astype(np.float32)
        act_lens = rng.randint(max(1, T // 2), T + 1, size=B)
        act_lens[0] = T
        labels, label_lens = [], []
        for b in range(B):
            L = rng.randint(0 if i == 1 else 1, Lmax + 1)
            lab = rng.randint(1, V, size=L)
            for j in range(1, L):
                if rng.uniform() < 0.2:
                    lab[j] = lab[j - 1]
            while L + int(np.sum(lab[1:L] == lab[:L - 1])) > act_lens[b]:
                L -= 1
            labels.append(lab[:L])
            label_lens.append(L)
        flat = np.concatenate(labels).astype(np.int32)
        a = torch.tensor(acts, dtype=torch.float64, requires_grad=True)
        lp = torch.log_softmax(a, dim=2)
        costs = torch.nn.functional.ctc_loss(lp, torch.tensor(flat, dtype=torch.long), torch.tensor(act_lens),
                                             torch.tensor(label_lens), blank=0, reduction="none")
        costs.sum().backward()
        out["acts_%d" % i] = acts
        out["labels_%d" % i] = flat
        out["act_lens_%d" % i] = act_lens.astype(np.int32)
        out["label_lens_%d" % i] = np.array(label_lens, dtype=np.int32)
        out["costs_%d" % i] = costs.detach().numpy()
        out["grads_%d" % i] = a.grad.numpy().astype(np.float32)
    out["n_cases"] = np.array(len(shapes))
    np.savez_compressed(os.path.join(HERE, "ctc_golden.npz"), **out)


if __name__ == "__main__":
    greedy_cases()
    ctc_cases()
    print("golden fixtures written to", HERE)
