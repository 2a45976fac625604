"""Synthetic CTC workloads for the five BASELINE.json configs.

The reference ships no corpus and no network is available, so every benchmark
and full-size parity case uses seeded synthetic inputs of the named shape:
unnormalised fp32 logits ``acts[T,B,V] ~ N(0,1)``, flat int32 labels uniform in
``[1, V-1]`` with 10 % forced adjacent repeats (to exercise the CTC repeat
rule), blank = 0 (reference: models/pytorch_v3/ctc/ctc.py:267-269,299-300), and
every utterance feasible (``L_b + repeats_b <= T_b``).  Utterances are sorted by
descending input length, as the reference's encoder does before the loss
(models/pytorch_v3/encoders/rnn.py:318-321).
"""

from collections import namedtuple

import numpy as np

Config = namedtuple("Config", "name B T V Lmax kind")

CONFIGS = {
    "C1": Config("TIMIT-shaped B=32 T=775 V=62 L<=75", 32, 775, 62, 75, "var"),
    "C2": Config("WSJ-shaped B=64 T=1500 V=33 L<=200", 64, 1500, 33, 200, "var"),
    "C3": Config("LibriSpeech-shaped B=128 T=800 V=30 L<=400", 128, 800, 30, 400, "full"),
    "C4": Config("CSJ-shaped B=64 T=1000 V=3386 L<=150", 64, 1000, 3386, 150, "var"),
    "C5": Config("LibriSpeech-960h sweep B=512 T=400-1600 V=30", 512, 1600, 30, 400, "sweep"),
}
_SEED_INDEX = {"C1": 0, "C2": 1, "C3": 2, "C4": 3, "C5": 4}

Workload = namedtuple("Workload", "name T B V labels label_lens act_lens seed")


def _repeats(lab):
    return int(np.sum(lab[1:] == lab[:-1])) if len(lab) > 1 else 0


def make_lengths_and_labels(cfg_key, B=None, T=None, V=None, Lmax=None, kind=None, seed=None):
    """Lengths and labels (host, numpy) for a named config or a custom shape."""
    if cfg_key is not None:
        cfg = CONFIGS[cfg_key]
        B, T, V, Lmax, kind = cfg.B, cfg.T, cfg.V, cfg.Lmax, cfg.kind
        seed = 1623 + _SEED_INDEX[cfg_key]
        name = cfg.name
    else:
        name = "custom B=%d T=%d V=%d L<=%d" % (B, T, V, Lmax)
        seed = 1623 if seed is None else seed
    rng = np.random.RandomState(seed)
    if kind == "full":
        act_lens = np.full(B, T, dtype=np.int64)
    elif kind == "sweep":
        act_lens = rng.randint(T // 4, T + 1, size=B)
        act_lens = np.sort(act_lens)[::-1].copy()
    else:
        act_lens = rng.randint((T + 1) // 2, T + 1, size=B)
        act_lens[0] = T
        act_lens = np.sort(act_lens)[::-1].copy()
    if kind == "sweep":
        label_lens = np.clip(np.floor(act_lens * rng.uniform(0.15, 0.30, size=B)), 1, Lmax).astype(np.int64)
    else:
        label_lens = rng.randint((Lmax + 1) // 2, Lmax + 1, size=B)
    labels = []
    for b in range(B):
        L = int(label_lens[b])
        while True:
            lab = rng.randint(1, V, size=L)
            if L > 1:
                copy_prev = rng.uniform(size=L) < 0.1
                for i in range(1, L):
                    if copy_prev[i]:
                        lab[i] = lab[i - 1]
            if L + _repeats(lab) <= int(act_lens[b]):
                break
            L -= 1                      # reduce until feasible
        label_lens[b] = L
        labels.append(lab.astype(np.int32))
    flat = np.concatenate(labels).astype(np.int32) if labels else np.zeros(0, np.int32)
    return Workload(name, T, B, V, flat, label_lens.astype(np.int32), act_lens.astype(np.int32), seed)


def make_acts(wl, device="cpu", copy_index=0):
    """fp32 logits [T, B, V] ~ N(0,1) from a seeded torch.Generator."""
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(wl.seed + 7919 * copy_index)
    acts = torch.randn(wl.T, wl.B, wl.V, generator=g, dtype=torch.float32)
    return acts.to(device)


def algorithmic_bytes(wl):
    """read acts + write grad + label-indexed gathers, valid frames only.
    Returns (bytes, strict_dram_bytes, frames)."""
    T_b = wl.act_lens.astype(np.int64)
    L_b = wl.label_lens.astype(np.int64)
    strict = int(np.sum(8 * T_b * wl.V)) + 4 * int(L_b.sum()) + 8 * wl.B
    gathers = int(np.sum(8 * T_b * (2 * L_b + 1)))
    return strict + gathers, strict, int(T_b.sum())
