"""Developer aid (GPU box): time the greedy decoder call and its kernels in isolation."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytorch_end2end_speech_recognition_b200 as b200
from pytorch_end2end_speech_recognition_b200 import workloads
for key in sys.argv[1:] or ["C3", "C4"]:
    wl = workloads.make_lengths_and_labels(key)
    logits = [workloads.make_acts(wl, copy_index=100 + i).transpose(0, 1).contiguous().cuda() for i in range(4)]
    lens = torch.from_numpy(wl.act_lens.astype(np.int32)).cuda()
    for i in range(8):
        b200.greedy_decode(logits[i % 4], lens)
    torch.cuda.synchronize()
    for keep in (False, True):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        outs = []
        for i in range(100):
            o = b200.greedy_decode(logits[i % 4], lens)
            if keep: outs.append(o)
        e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(key, "keep" if keep else "drop", "device %.4f ms/call  enqueue %.4f ms/call  wall %.4f" % (e0.elapsed_time(e1) / 100, (t1 - t0) * 10, (t2 - t0) * 10))
