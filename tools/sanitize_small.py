"""Small cases for compute-sanitizer (memcheck / racecheck): every kernel path once, sizes kept tiny.
  compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pytorch_end2end_speech_recognition_b200 as b200
from pytorch_end2end_speech_recognition_b200 import workloads

for kw in (dict(B=3, T=37, V=30, Lmax=12), dict(B=2, T=90, V=30, Lmax=40), dict(B=2, T=300, V=30, Lmax=140),
           dict(B=2, T=60, V=200, Lmax=20), dict(B=1, T=5, V=4, Lmax=2), dict(B=3, T=50, V=301, Lmax=16),
           dict(B=2, T=40, V=1003, Lmax=12)):
    wl = workloads.make_lengths_and_labels(None, kind="var", seed=3, **kw)
    acts = workloads.make_acts(wl).cuda()
    costs, loss, grads = b200.ctc_loss_and_grad(acts, wl.labels, wl.act_lens, wl.label_lens)
    c2, _, _ = b200.ctc_loss_and_grad(acts, wl.labels, wl.act_lens, wl.label_lens, need_grad=False)
    torch.cuda.synchronize()
    assert torch.allclose(costs, c2, rtol=1e-5)
    print(kw, float(loss))
logits = torch.randn(3, 40, 30, device="cuda")
print(b200.greedy_decode(logits, np.array([40, 33, 20], np.int32))[1].tolist())
