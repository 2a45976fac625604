"""Developer tool (GPU box): H2D time of each of bench.py's rotating pinned logits buffers (is one of them
slow, e.g. on the other NUMA node?), and the CPU affinity NVML reports for the GPU."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from pytorch_end2end_speech_recognition_b200 import workloads


def main():
    dev = torch.device("cuda", 0)
    wl = workloads.make_lengths_and_labels("C3")
    hosts = [workloads.make_acts(wl, copy_index=i) for i in range(14)]
    pinned = [a.pin_memory() for a in hosts]
    d = torch.empty_like(pinned[0], device=dev)
    for rep in range(2):
        out = []
        for i, p in enumerate(pinned):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(); d.copy_(p, non_blocking=True); e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1))
        print("H2D ms per pinned buffer:", " ".join("%.3f" % x for x in out))
    print("process affinity:", sorted(os.sched_getaffinity(0)))
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        print("GPU 0 CPU affinity (NVML):", cpus[:8], "...", len(cpus), "cpus; cpu_count", os.cpu_count())
    except Exception as e:
        print("nvml affinity failed:", e)
    try:
        print(open("/sys/devices/system/node/online").read().strip(), "numa nodes online")
    except Exception as e:
        print(e)


if __name__ == "__main__":
    main()
