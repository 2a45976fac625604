"""Generates the committed golden fixtures (run in the build container, where /root/reference exists).

  greedy_golden.npz : inputs and outputs of the REFERENCE's own GreedyDecoder
                      (models/pytorch_v3/ctc/decoders/greedy_decoder.py), imported from
                      /root/reference and called per utterance (B=1 slices: its final
                      np.array(best_hyps) raises on ragged results under numpy >= 1.24).
  ctc_golden.npz    : CTC costs/gradients from torch.nn.functional.ctc_loss (CPU, float64) -- an
                      implementation independent of oracle/ -- on small seeded cases.  The
                      reference's own loss (warp-ctc) cannot be run: it is not vendored.

  ctc_reference_golden.npz : per-utterance costs and gradients from the REFERENCE'S OWN in-tree CTC,
                      ``models/chainer/ctc/ctc_loss_from_chainer.py`` (class
                      ConnectionistTemporalClassification, numpy branch, float32, reduce='no'),
                      imported from /root/reference and run here.  The file only needs the
                      ``chainer`` package for a base class, ``cuda.get_array_module`` and
                      ``type_check``; chainer is not installed, so a ten-line stub module stands in
                      for those three names -- every line of CTC arithmetic that runs is the
                      reference's.  This is what pins oracle/ to the reference.

  edit_distance_golden.npz : (distance, substitutions, insertions, deletions) returned by the REFERENCE's own
                      ``compute_wer`` (utils/evaluation/edit_distance.py:53-126, imported from /root/reference;
                      the module-level ``import Levenshtein`` -- a package that is not installed and that
                      compute_wer never calls -- is satisfied by an empty stub) on seeded token lists.  Pairs on
                      which the reference raises (its backtrace indexes the matrix with -1 at the borders; its
                      callers swallow the exception, phone.py:92-103) are recorded with ``ok = 0``.

  beam_golden.npz   : hypotheses of the REFERENCE's own BeamSearchDecoder
                      (models/pytorch_v3/ctc/decoders/beam_search_decoder.py), imported from /root/reference and
                      called per utterance, beam widths 1 / 2 / 10 / 20, fed float32 log-softmax outputs upcast to
                      float64 -- the arithmetic the reference performs under the numpy it was written for (python
                      float + float32 scalar = float64 before NEP 50).  The hypotheses it returns for the float32
                      array itself under this container's numpy 2 (float32 arithmetic) are stored next to them
                      (``hyp32_*``) and are identical on every case.

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


def _import_reference_chainer_ctc():
    """Import models/chainer/ctc/ctc_loss_from_chainer.py from /root/reference with a stub for the
    (absent) chainer package: Function base class, cuda.get_array_module -> numpy, type_check,
    utils.force_array, is_debug.  No arithmetic lives in the stub."""
    import collections
    import collections.abc
    import importlib.util
    import types
    if not hasattr(collections, "Sequence"):          # the 2018 file uses the pre-3.10 alias
        collections.Sequence = collections.abc.Sequence
    chainer = types.ModuleType("chainer")
    chainer.is_debug = lambda: False
    backends = types.ModuleType("chainer.backends")
    cuda = types.ModuleType("chainer.backends.cuda")
    cuda.get_array_module = lambda *a: np
    function = types.ModuleType("chainer.function")
    function.Function = type("Function", (object,), {})
    utils = types.ModuleType("chainer.utils")
    utils.force_array = lambda x, dtype=None: np.asarray(x, dtype=dtype)
    type_check = types.ModuleType("chainer.utils.type_check")
    type_check.expect = lambda *a, **k: None
    chainer.backends, backends.cuda = backends, cuda
    chainer.function, chainer.utils, utils.type_check = function, utils, type_check
    for name, mod in (("chainer", chainer), ("chainer.backends", backends), ("chainer.backends.cuda", cuda),
                      ("chainer.function", function), ("chainer.utils", utils),
                      ("chainer.utils.type_check", type_check)):
        sys.modules[name] = mod
    spec = importlib.util.spec_from_file_location(
        "reference_chainer_ctc", "/root/reference/models/chainer/ctc/ctc_loss_from_chainer.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_ctc_cases():
    """Costs and gradients of the reference's in-tree CTC on seeded cases: variable input lengths,
    L=0, adjacent repeats, L+repeats == T (tight), a V=62 phone-sized and a V=200 case."""
    ref = _import_reference_chainer_ctc()
    rng = np.random.RandomState(2018)
    out = {}
    shapes = [(1, 1, 3, 1), (1, 2, 4, 1), (3, 12, 5, 4), (4, 40, 30, 12), (3, 31, 62, 14), (2, 30, 200, 8),
              (2, 64, 7, 30), (4, 96, 30, 40)]
    for i, (B, T, V, Lmax) in enumerate(shapes):
        acts = (rng.randn(T, B, V) * (2.5 if i == 4 else 1.0)).astype(np.float32)
        act_lens = rng.randint(max(1, T // 2), T + 1, size=B)
        act_lens[0] = T
        padded = np.zeros((B, max(Lmax, 1)), dtype=np.int32)
        label_lens = []
        for b in range(B):
            L = rng.randint(0 if i == 2 else 1, Lmax + 1)
            if i == 6 and b == 0:
                L = Lmax
            lab = rng.randint(1, V, size=L)
            for j in range(1, L):
                if rng.uniform() < 0.2:
                    lab[j] = lab[j - 1]
            while L + int(np.sum(lab[1:L] == lab[:L - 1])) > act_lens[b]:
                L -= 1
            padded[b, :L] = lab[:L]
            label_lens.append(L)
        label_lens = np.array(label_lens, dtype=np.int32)
        fn = ref.ConnectionistTemporalClassification(0, reduce="no")       # blank = 0 as in ctc.py:267-269
        inputs = (act_lens.astype(np.int32), label_lens, padded, acts.copy())
        loss, = fn.forward(inputs)
        grad = fn.backward(inputs, (np.ones(B, dtype=np.float32),))[3]
        out["acts_%d" % i] = acts
        out["labels_%d" % i] = np.concatenate([padded[b, :label_lens[b]] for b in range(B)]).astype(np.int32)
        out["act_lens_%d" % i] = act_lens.astype(np.int32)
        out["label_lens_%d" % i] = label_lens
        out["costs_%d" % i] = np.asarray(loss, dtype=np.float32)
        out["grads_%d" % i] = np.asarray(grad, dtype=np.float32)
    out["n_cases"] = np.array(len(shapes))
    np.savez_compressed(os.path.join(HERE, "ctc_reference_golden.npz"), **out)


def edit_distance_cases():
    import importlib.util
    import types
    sys.modules.setdefault("Levenshtein", types.ModuleType("Levenshtein"))      # never called by compute_wer
    spec = importlib.util.spec_from_file_location("reference_edit_distance",
                                                  "/root/reference/utils/evaluation/edit_distance.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.RandomState(53)
    refs, hyps, outs, oks = [], [], [], []
    for i in range(160):
        V = int(rng.choice([2, 3, 5, 30, 62]))
        R, H = int(rng.randint(0 if i % 9 == 0 else 1, 40)), int(rng.randint(0 if i % 7 == 0 else 1, 40))
        ref = rng.randint(0, V, size=R)
        if i % 3 == 0 and R > 0:                      # hypothesis = corrupted reference (the realistic case)
            hyp = list(ref)
            for _ in range(rng.randint(0, 6)):
                op, pos = rng.randint(3), rng.randint(0, max(len(hyp), 1))
                if op == 0 and hyp:
                    hyp[min(pos, len(hyp) - 1)] = int(rng.randint(0, V))
                elif op == 1:
                    hyp.insert(pos, int(rng.randint(0, V)))
                elif hyp:
                    del hyp[min(pos, len(hyp) - 1)]
            hyp = np.array(hyp, dtype=np.int64)
        else:
            hyp = rng.randint(0, V, size=H)
        try:
            with np.errstate(all="ignore"):
                wer, sub, ins, dele = mod.compute_wer([int(x) for x in ref], [int(x) for x in hyp], normalize=False)
            res, ok = (int(wer), int(sub), int(ins), int(dele)), 1
        except Exception:
            res, ok = (0, 0, 0, 0), 0
        refs.append(ref); hyps.append(hyp); outs.append(res); oks.append(ok)
    rmax, hmax = max(len(r) for r in refs), max(len(h) for h in hyps)
    ref_pad = np.full((len(refs), rmax), -1, np.int32)
    hyp_pad = np.full((len(hyps), hmax), -2, np.int32)
    for b, (r, h) in enumerate(zip(refs, hyps)):
        ref_pad[b, :len(r)] = r
        hyp_pad[b, :len(h)] = h
    np.savez_compressed(os.path.join(HERE, "edit_distance_golden.npz"), refs=ref_pad, hyps=hyp_pad,
                        ref_lens=np.array([len(r) for r in refs], np.int32),
                        hyp_lens=np.array([len(h) for h in hyps], np.int32),
                        out=np.array(outs, np.int32), ok=np.array(oks, np.int32))


def beam_cases():
    sys.path.insert(0, "/root/reference")
    from models.pytorch_v3.ctc.decoders.beam_search_decoder import BeamSearchDecoder  # the reference itself
    dec = BeamSearchDecoder(blank_index=0)
    rng = np.random.RandomState(77)
    out = {}
    shapes = [(3, 30, 6, 2), (2, 40, 30, 10), (2, 50, 30, 20), (2, 35, 62, 10), (3, 25, 5, 1), (2, 20, 30, 2),
              (2, 28, 3, 10), (1, 60, 30, 10)]
    n_same = 0
    for i, (B, T, V, W) in enumerate(shapes):
        scale = [1.0, 2.0, 3.0, 1.5, 1.0, 4.0, 1.0, 2.5][i]
        logits = (rng.randn(B, T, V) * scale).astype(np.float32)
        if i in (1, 7):
            logits[:, :, 0] += 2.0                    # blank-dominated, as a trained CTC model's output
        if i == 4:
            logits[0, :, :] = 0.0                     # exact ties everywhere: the tie order of the sort decides
        lp = torch.log_softmax(torch.from_numpy(logits), dim=-1).numpy()           # float32, as ctc.py:439-441 feeds it
        x_lens = rng.randint(T // 2, T + 1, size=B)
        x_lens[0] = T
        if i == 6:
            x_lens[1] = 0
        hyps, hyps32 = [], []
        for b in range(B):
            with np.errstate(all="ignore"):
                h = dec(lp[b:b + 1].astype(np.float64), x_lens[b:b + 1], beam_width=W)
                h32 = dec(lp[b:b + 1], x_lens[b:b + 1], beam_width=W)
            hyps.append(np.asarray(h[0], dtype=np.int64).reshape(-1))
            hyps32.append(np.asarray(h32[0], dtype=np.int64).reshape(-1))
            n_same += int(np.array_equal(hyps[-1], hyps32[-1]))
        out["log_probs_%d" % i] = lp
        out["x_lens_%d" % i] = x_lens.astype(np.int32)
        out["beam_%d" % i] = np.array(W)
        out["hyp_lens_%d" % i] = np.array([len(h) for h in hyps], dtype=np.int32)
        out["hyp_flat_%d" % i] = np.concatenate(hyps) if hyps else np.zeros(0, np.int64)
        out["hyp32_lens_%d" % i] = np.array([len(h) for h in hyps32], dtype=np.int32)
        out["hyp32_flat_%d" % i] = np.concatenate(hyps32) if hyps32 else np.zeros(0, np.int64)
    out["n_cases"] = np.array(len(shapes))
    print("beam: float64 and float32 runs of the reference agree on %d of %d utterances" % (n_same, sum(s[0] for s in shapes)))
    np.savez_compressed(os.path.join(HERE, "beam_golden.npz"), **out)


if __name__ == "__main__":
    greedy_cases()
    ctc_cases()
    reference_ctc_cases()
    edit_distance_cases()
    beam_cases()
    print("golden fixtures written to", HERE)
