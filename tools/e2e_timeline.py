"""Developer tool (GPU box): host-side timeline of bench.py's pipelined end-to-end step (C3): where the host
spends its time per step (enqueue of the call, of the copies, waiting for the previous loss)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import pytorch_end2end_speech_recognition_b200 as b200
from pytorch_end2end_speech_recognition_b200 import workloads


def main():
    dev = torch.device("cuda", 0)
    wl = workloads.make_lengths_and_labels("C3")
    NB = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    pinned = [workloads.make_acts(wl, copy_index=i).pin_memory() for i in range(NB)]
    stage = [torch.empty_like(pinned[0], device=dev) for _ in range(2)]
    grads = [torch.empty_like(stage[0]) for _ in range(NB)]
    costs = torch.empty(wl.B, device=dev)
    cs = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    copied = [[torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()]]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    comp = torch.cuda.current_stream(dev)
    half = wl.T // 2
    ldev = [torch.empty(1, device=dev) for _ in range(2)]
    lhost = [torch.empty(1).pin_memory() for _ in range(2)]
    lread = [torch.cuda.Event(), torch.cuda.Event()]
    gpu_ev = [[torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)] for _ in range(64)]
    cp_ev = [[torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)] for _ in range(64)]

    def issue_copy(i):
        d = i % 2
        for k, (lo, hi) in enumerate(((0, half), (half, wl.T))):
            with torch.cuda.stream(cs[k]):
                cs[k].wait_event(consumed[d])
                if k == 0: cp_ev[i % 64][0].record(cs[k])
                stage[d][lo:hi].copy_(pinned[i % NB][lo:hi], non_blocking=True)
                if k == 0: cp_ev[i % 64][1].record(cs[k])
                copied[d][k].record(cs[k])

    for d in (0, 1):
        consumed[d].record(comp)
    stamps = []
    torch.cuda.synchronize()
    issue_copy(0)
    n = 40
    t_start = time.perf_counter()
    for i in range(n):
        d = i % 2
        t0 = time.perf_counter()
        comp.wait_event(copied[d][0]); comp.wait_event(copied[d][1])
        gpu_ev[i][0].record(comp)
        b200.ctc_loss_and_grad(stage[d], wl.labels, wl.act_lens, wl.label_lens, grads=grads[i % NB], costs=costs, loss_sum=ldev[d])
        gpu_ev[i][1].record(comp)
        t1 = time.perf_counter()
        consumed[d].record(comp)
        lhost[d].copy_(ldev[d], non_blocking=True)
        lread[d].record(comp)
        t2 = time.perf_counter()
        issue_copy(i + 1)
        t3 = time.perf_counter()
        if i > 0:
            lread[(i - 1) % 2].synchronize()
            float(lhost[(i - 1) % 2][0])
        t4 = time.perf_counter()
        stamps.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    torch.cuda.synchronize()
    total = time.perf_counter() - t_start
    a = np.array(stamps[4:]) * 1e6
    print("per step (us, median): call enqueue %.0f | loss D2H enqueue %.0f | logits H2D enqueue %.0f | wait previous loss %.0f | wall %.0f"
          % (tuple(np.median(a, axis=0)) + (total / n * 1e6,)))
    print("per step (us, max):    call enqueue %.0f | loss D2H enqueue %.0f | logits H2D enqueue %.0f | wait previous loss %.0f" % tuple(a.max(axis=0)))
    print("per-step host totals (us):", " ".join("%.0f" % x for x in np.array(stamps).sum(axis=1) * 1e6))
    g = [gpu_ev[i][0].elapsed_time(gpu_ev[i][1]) for i in range(4, n)]
    gap = [gpu_ev[i][1].elapsed_time(gpu_ev[i + 1][0]) for i in range(4, n - 1)]
    c = [cp_ev[i][0].elapsed_time(cp_ev[i][1]) for i in range(5, n)]
    print("H2D of the first half of the logits: %.3f ms (median)" % float(np.median(c)))
    print("GPU: call %.3f ms (median), gap between calls %.3f ms" % (float(np.median(g)), float(np.median(gap))))


if __name__ == "__main__":
    main()
