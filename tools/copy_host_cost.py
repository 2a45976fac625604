"""Developer tool (GPU box): HOST time of enqueueing a pinned host->device copy of C3's logits (12.3 MB, in two
halves), through torch's Tensor.copy_ and through cudaMemcpyAsync (cuda-python), with the GPU idle and with a
lattice kernel in flight."""
import sys, time
import numpy as np
import torch
from cuda.bindings import runtime as cudart
sys.path.insert(0, ".")
import pytorch_end2end_speech_recognition_b200 as b200
from pytorch_end2end_speech_recognition_b200 import workloads


def main():
    dev = torch.device("cuda", 0)
    wl = workloads.make_lengths_and_labels("C3")
    host = workloads.make_acts(wl).pin_memory()
    work = workloads.make_acts(wl).to(dev)
    stage = torch.empty_like(work)
    grads = torch.empty_like(work); costs = torch.empty(wl.B, device=dev); loss = torch.empty(1, device=dev)
    s = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    half = wl.T // 2
    nb = host[:half].numel() * 4
    H2D = cudart.cudaMemcpyKind.cudaMemcpyHostToDevice

    def torch_copy():
        for k, (lo, hi) in enumerate(((0, half), (half, wl.T))):
            with torch.cuda.stream(s[k]):
                stage[lo:hi].copy_(host[lo:hi], non_blocking=True)

    def raw_copy():
        for k, lo in enumerate((0, half)):
            cudart.cudaMemcpyAsync(stage[lo:].data_ptr(), host[lo:].data_ptr(), nb, H2D, s[k].cuda_stream)

    for name, fn in (("torch copy_", torch_copy), ("cudaMemcpyAsync", raw_copy)):
        for busy in (False, True):
            ts = []
            for _ in range(20):
                torch.cuda.synchronize()
                if busy:
                    b200.ctc_loss_and_grad(work, wl.labels, wl.act_lens, wl.label_lens, grads=grads, costs=costs, loss_sum=loss)
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            torch.cuda.synchronize()
            print("%-16s GPU %-5s: host time to enqueue both halves: median %.1f us" % (name, "busy" if busy else "idle", 1e6 * float(np.median(ts))))
    assert torch.equal(stage.cpu(), host)


if __name__ == "__main__":
    main()
