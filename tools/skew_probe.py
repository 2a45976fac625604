"""Developer aid (GPU box): how much does a skewed symbol distribution (text-like: a few symbols carry most of
the labels) cost the reducers?  C3 shape, labels re-drawn from p_k ~ 1 / (k + 1)^s.   python tools/skew_probe.py [s ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_end2end_speech_recognition_b200 as b200
from pytorch_end2end_speech_recognition_b200 import ctc as ctc_mod, workloads

wl0 = workloads.make_lengths_and_labels("C3")
acts = [workloads.make_acts(wl0, copy_index=i).cuda() for i in range(8)]
for s in [float(a) for a in sys.argv[1:]] or [0.0, 0.7, 1.0, 1.5]:
    rng = np.random.RandomState(5)
    p = 1.0 / (np.arange(wl0.V - 1) + 1.0) ** s
    p /= p.sum()
    labels = (1 + rng.choice(wl0.V - 1, size=wl0.labels.size, p=p)).astype(np.int32)
    wl = wl0._replace(labels=labels)
    cnt = np.bincount(labels[:int(wl.label_lens[0])], minlength=wl.V)
    for i in range(10):
        b200.ctc_loss_and_grad(acts[i % 8], wl.labels, wl.act_lens, wl.label_lens)
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(64):
            b200.ctc_loss_and_grad(acts[i % 8], wl.labels, wl.act_lens, wl.label_lens)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 64)
    print("exponent %.1f: largest symbol group of utterance 0: %d of %d labels; %.4f ms per step; fallbacks %s" % (
        s, cnt.max(), int(wl.label_lens[0]), best, ctc_mod.last_fallbacks()))
