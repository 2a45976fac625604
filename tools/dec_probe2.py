import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench
import torch.distributed as dist
import pytorch_end2end_speech_recognition_b200 as b200
from pytorch_end2end_speech_recognition_b200 import workloads
dev = torch.device("cuda", 0)
barrier, timed = bench.make_timer(torch, dist, dev, 1)
for k in ("C3", "C3", "C4"):
    print(bench.decoder_line(torch, dev, timed, workloads, b200, k))
