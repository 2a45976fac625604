"""Developer tool (GPU box): where the time of bench.py's end-to-end step goes (C3): duration of the H2D copy
of the logits when split over n copy streams, alone and while the previous step's kernels run."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import pytorch_end2end_speech_recognition_b200 as b200
from pytorch_end2end_speech_recognition_b200 import workloads

def main():
    dev = torch.device("cuda", 0)
    wl = workloads.make_lengths_and_labels("C3")
    hosts = [workloads.make_acts(wl, copy_index=i).pin_memory() for i in range(4)]
    stage = [torch.empty_like(hosts[0], device=dev) for _ in range(2)]
    grads = torch.empty_like(stage[0])
    costs = torch.empty(wl.B, device=dev); loss = torch.empty(1, device=dev)
    comp = torch.cuda.current_stream(dev)
    for nsplit in (1, 2, 4):
        streams = [torch.cuda.Stream(device=dev) for _ in range(nsplit)]
        bounds = np.linspace(0, wl.T, nsplit + 1).astype(int)
        def copy(i, evs=None):
            for k, cs in enumerate(streams):
                with torch.cuda.stream(cs):
                    if evs: evs[k][0].record(cs)
                    stage[i % 2][bounds[k]:bounds[k + 1]].copy_(hosts[i % 4][bounds[k]:bounds[k + 1]], non_blocking=True)
                    if evs: evs[k][1].record(cs)
        for with_kernel in (False, True):
            torch.cuda.synchronize()
            durs = []
            t0 = time.perf_counter()
            for i in range(20):
                evs = [[torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)] for _ in streams]
                if with_kernel:
                    b200.ctc_loss_and_grad(stage[(i + 1) % 2], wl.labels, wl.act_lens, wl.label_lens, grads=grads, costs=costs, loss_sum=loss)
                copy(i, evs)
                torch.cuda.synchronize()
                durs.append(max(e[0].elapsed_time(e[1]) for e in evs))
            wall = (time.perf_counter() - t0) / 20 * 1e3
            print("split %d  kernels concurrently: %-5s  slowest part %.3f ms (median)  wall per iteration %.3f ms"
                  % (nsplit, with_kernel, float(np.median(durs)), wall))

def host_cost():
    """Host time of one ctc_loss_and_grad call (enqueue only; the GPU is idle when the call starts)."""
    dev = torch.device("cuda", 0)
    for key in ("C1", "C3", "C5"):
        wl = workloads.make_lengths_and_labels(key)
        acts = workloads.make_acts(wl).to(dev)
        grads = torch.empty_like(acts); costs = torch.empty(wl.B, device=dev); loss = torch.empty(1, device=dev)
        ts = []
        for _ in range(20):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            b200.ctc_loss_and_grad(acts, wl.labels, wl.act_lens, wl.label_lens, grads=grads, costs=costs, loss_sum=loss)
            ts.append(time.perf_counter() - t0)
        print("%s: host time per call (enqueue) median %.1f us" % (key, 1e6 * float(np.median(ts))))


if __name__ == "__main__":
    if "--host" in sys.argv:
        host_cost()
    else:
        main()
