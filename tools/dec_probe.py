import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import pytorch_end2end_speech_recognition_b200 as b200
from pytorch_end2end_speech_recognition_b200 import workloads
for key in ("C3", "C4"):
    wl = workloads.make_lengths_and_labels(key)
    logits = [workloads.make_acts(wl, copy_index=100 + i).transpose(0, 1).contiguous().cuda() for i in range(3)]
    lens = torch.from_numpy(wl.act_lens.astype(np.int32)).cuda()
    for i in range(3): b200.greedy_decode(logits[i % 3], lens)
    torch.cuda.synchronize()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for i in range(30): b200.greedy_decode(logits[i % 3], lens)
        e1.record(); h = time.perf_counter() - t0; torch.cuda.synchronize()
        print(key, "device %.4f ms/call host %.4f ms/call" % (e0.elapsed_time(e1) / 30, h * 1e3 / 30))
