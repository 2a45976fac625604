"""bench.py -- CTC fwd+bwd frames/s on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3] [--impl b200|reference]

A "step" is one fused loss+gradient evaluation (two kernels: softmax rows -> lattice, whose last CTA
also sums the costs; the backward
of the op is an elementwise scale of the gradient computed here) over one mini-batch
of synthetic logits of the named shape.  N > 1: one process per GPU (torchrun), every rank owns a
mini-batch of the same shape (data-parallel training: utterances never cross GPUs) and the scalar
loss is all-reduced over NCCL every step -- weak scaling.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ctc_fwd_bwd_frames_per_sec"
UNIT = "frames/s"
N_ROTATE = 16          # distinct acts/grads buffer sets cycled through the timed loop (> L2 in total)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_traffic(workload_key):
    """DRAM bytes per lattice launch from the committed ncu capture (profiles/r01_traffic.json, C3 only)."""
    if workload_key != "C3":
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            k = json.load(f)["lattice_kernel"]
        return int(k["dram_bytes_read"]) + int(k["dram_bytes_write"])
    except Exception:
        return None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.active = threading.Event()          # set while the timed region runs: only those samples count
        self.ready = threading.Event()           # NVML is initialised and has answered once (its start-up is slow)
        self.sm, self.reasons, self.sm_max = [], set(), None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                getattr(pynvml, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.ready.set()
            while not self.stop_flag.is_set():
                if self.active.is_set():
                    self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    try:
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for bit, name in names.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                # every 4 ms: an NVML query holds a driver lock that kernel launches also take (worse with more
                # GPUs in the box) -- at 1 kHz the sampler itself cost one rank of two 7 % of its timed region
                time.sleep(0.004)
        except Exception:
            self.ready.set()
            self._smi()

    def _smi(self):
        import subprocess
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            if not self.active.is_set():
                time.sleep(0.001)
                continue
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.sm.append(int(f[0])); self.sm_max = int(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                break
            time.sleep(0.1)

    def sample_now(self):
        """One sample taken by the caller's thread (the GPU is busy: the caller keeps its queue full)."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            if self.sm_max is None:
                self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for bit, name in ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                              (0x4, "sw_power_cap")):
                if r & bit:
                    self.reasons.add(name)
            return True
        except Exception:
            pass
        try:
            import subprocess
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            f = [x.strip() for x in out.strip().split(",")]
            self.sm.append(int(f[0])); self.sm_max = int(f[1])
            return True
        except Exception:
            return False

    def result(self):
        self.stop_flag.set()
        self.join(timeout=5)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def _host_threads():
    """All the host cores this process may use (torchrun exports OMP_NUM_THREADS=1: ignore it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(wl, acts_np, budget_s=12.0, kind_note=""):
    """The oracle's C++/OpenMP restatement of the warp-ctc CPU path (float), timed on this host."""
    from oracle import ctc_cpu
    threads = _host_threads()
    ctc_cpu.ctc_cpu(acts_np[:, :min(8, wl.B)], *_slice(wl, min(8, wl.B)))      # warm-up (page in, build)
    reps, t0 = 0, time.perf_counter()
    while True:
        ctc_cpu.ctc_cpu(acts_np, wl.labels, wl.act_lens, wl.label_lens, precision="f32", num_threads=threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or reps >= 20:
            break
    frames = int(wl.act_lens.sum())
    return {"value": frames * reps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d full mini-batches of %s (%d frames each), C++/OpenMP fp32 restatement of the "
                      "warp-ctc CPU path%s" % (reps, wl.name, frames, kind_note)}


def _slice(wl, nb):
    L = int(wl.label_lens[:nb].sum())
    return wl.labels[:L], wl.act_lens[:nb], wl.label_lens[:nb]


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  warp-ctc itself is not in
    /root/reference (un-vendored, un-pinned), so the oracle's C++/OpenMP port is timed, with all the
    host threads, on the same workload, one full mini-batch per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pytorch_end2end_speech_recognition_b200 import workloads
    from oracle import ctc_cpu
    wl = workloads.make_lengths_and_labels(args.workload)
    acts = workloads.make_acts(wl).numpy()
    frames = int(wl.act_lens.sum())
    threads = _host_threads()
    for _ in range(max(args.warmup, 1)):
        ctc_cpu.ctc_cpu(acts, wl.labels, wl.act_lens, wl.label_lens, precision="f32", num_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctc_cpu.ctc_cpu(acts, wl.labels, wl.act_lens, wl.label_lens, precision="f32", num_threads=threads)
    dt = time.perf_counter() - t0
    value = frames * args.steps / dt
    total, strict, _ = workloads.algorithmic_bytes(wl)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "frames_per_step": frames, "algorithmic_bytes_per_step": total},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "one full mini-batch per step; C++/OpenMP fp32 restatement of the warp-ctc CPU path "
                                   "(warp-ctc is not vendored in the reference)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import pytorch_end2end_speech_recognition_b200 as b200
    from pytorch_end2end_speech_recognition_b200 import ctc as ctc_mod
    from pytorch_end2end_speech_recognition_b200 import workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = workloads.make_lengths_and_labels(args.workload)
    frames = int(wl.act_lens.sum())
    total_bytes, strict_bytes, _ = workloads.algorithmic_bytes(wl)
    lattice_bytes = total_bytes - int(np.sum(8 * wl.act_lens.astype(np.int64) * wl.V))
    acts_bytes = wl.T * wl.B * wl.V * 4
    n_rot = max(2, min(N_ROTATE, int(np.ceil(160e6 / acts_bytes)))) if acts_bytes < 160e6 else 2
    acts_host = [workloads.make_acts(wl, copy_index=rank * N_ROTATE + i) for i in range(n_rot)]
    acts_dev = [a.to(dev) for a in acts_host]
    grads_dev = [torch.empty_like(a) for a in acts_dev]
    costs = torch.empty(wl.B, device=dev)
    loss = torch.empty(1, device=dev)
    # N > 1: the scalar loss of every step is all-reduced asynchronously (NCCL's own stream) while the next
    # steps compute -- a training loop only needs the number for logging.  The losses of four consecutive
    # steps travel in ONE all-reduce of four floats (the enqueue of an asynchronous all-reduce costs the host
    # ~70 us, which with the ~130 us of a call left the host slower than the 0.2 ms GPU step: 0.22-0.24 ms
    # per step at 4 ranks).  Two groups of four loss slots alternate; a group is rewritten only after its
    # all-reduce has completed.
    GROUP = 4
    loss_groups = [torch.zeros(GROUP, device=dev) for _ in range(2)]
    pending = [None, None]
    step_count = [0]

    def step(i):
        j = i % n_rot
        if world == 1:
            b200.ctc_loss_and_grad(acts_dev[j], wl.labels, wl.act_lens, wl.label_lens, grads=grads_dev[j],
                                   costs=costs, loss_sum=loss)
            return loss
        n = step_count[0]
        g, k = (n // GROUP) % 2, n % GROUP
        if k == 0 and pending[g] is not None:
            pending[g].wait()
            pending[g] = None
        slot = loss_groups[g][k:k + 1]
        b200.ctc_loss_and_grad(acts_dev[j], wl.labels, wl.act_lens, wl.label_lens, grads=grads_dev[j],
                               costs=costs, loss_sum=slot)
        if k == GROUP - 1:
            pending[g] = dist.all_reduce(loss_groups[g], async_op=True)   # the one exchange step of the path: the scalar losses
        step_count[0] = n + 1
        return slot

    def drain():
        if world == 1:
            return
        n = step_count[0]
        if n % GROUP:                                    # a partly filled group: its losses are exchanged now
            g = (n // GROUP) % 2
            pending[g] = dist.all_reduce(loss_groups[g], async_op=True)
            step_count[0] = (n // GROUP + 1) * GROUP
        for g in range(2):
            if pending[g] is not None:
                pending[g].wait()
                pending[g] = None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        ev_probe = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if os.environ.get("B200CTC_TIMED_DEBUG") else None
        for i in range(steps):
            fn(i)
            if ev_probe is not None and i in (4, steps - 1):
                ev_probe[0 if i == 4 else 1].record()
        drain()                          # every all-reduce issued in the timed region completes inside it
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if ev_probe is not None and steps > 8:
            print("rank %d timed: first 5 steps %.3f ms, steps 5..%d %.3f ms (%.4f each), drain %.3f ms" % (
                rank, ev0.elapsed_time(ev_probe[0]), steps - 1, ev_probe[0].elapsed_time(ev_probe[1]),
                ev_probe[0].elapsed_time(ev_probe[1]) / (steps - 5), ev_probe[1].elapsed_time(ev1)), file=sys.stderr)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.ready.wait(timeout=5.0)           # NVML start-up (slow, worse with eight ranks) stays out of the timed region
    # warm-up: W steps as asked; with several ranks at least 20, so that the first NCCL launches and the rank
    # skew after process start-up are behind us when the K timed steps begin
    for i in range(max(args.warmup, 3 if world == 1 else 20)):
        step(i)
    drain()
    sampler.active.set()
    t_host0 = time.perf_counter()
    ms = timed(step, args.steps)
    host_ms_per_step = (time.perf_counter() - t_host0) * 1e3 / args.steps   # wall clock of the same loop (diagnostic)
    sampler.active.clear()
    def agree_min(n):
        # every rank must take the same branch below: the extra steps contain collectives
        if world == 1:
            return n
        t = torch.tensor([n], device=dev, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t[0])

    n_in_region = len(sampler.sm)
    if os.environ.get("B200CTC_TEST_CLOCK_SKEW") and rank == 0:
        n_in_region = 0                       # test hook: one rank alone wants the fallback
    n_in_region = agree_min(n_in_region)
    if n_in_region < 3:
        # The timed region lasts ~10 ms and the sampling thread rarely gets the interpreter while the main thread
        # enqueues: take the remaining samples from the main thread under the SAME load right after the timed
        # region (every query is issued with >= 20 steps queued on the GPU).
        for _ in range(12):
            for i in range(24):
                step(i)
            sampler.sample_now()
        drain()
        torch.cuda.synchronize()
    if agree_min(len(sampler.sm)) == 0:
        sampler.active.set()
        timed(step, max(args.steps, 200))
        sampler.active.clear()
    clocks = sampler.result()
    clocks["sampled"] = ("in the timed region" if n_in_region >= 3 else
                         "%d in the timed region, the rest under the same load right after it" % n_in_region)
    ms_per_step = ms / args.steps
    value = frames * world / (ms_per_step * 1e-3)

    # ---- per-kernel device time (events inside the C library, separate pass) ----
    ctc_mod.set_profiling(True)
    k_ms = np.zeros(3)
    n_prof = min(args.steps, 20)
    for i in range(n_prof):
        step(i)
        k_ms += np.array(ctc_mod.last_kernel_ms())
    ctc_mod.set_profiling(False)
    k_ms /= n_prof
    peak, peak_kind = peaks()
    lattice_gbs = lattice_bytes / (k_ms[1] * 1e-3) / 1e9 if k_ms[1] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "lattice (alpha/beta recursion + occupancy update)",
                "achieved": lattice_gbs, "peak": peak, "peak_kind": peak_kind + " hbm copy GB/s", "unit": "GB/s",
                "frac": lattice_gbs / peak, "traffic": measured_traffic(args.workload),
                "kernel_ms": {"softmax_rows": k_ms[0], "lattice_and_cost_sum": k_ms[1]},
                "algorithmic_bytes_per_launch": lattice_bytes,
                "whole_step": {"algorithmic_bytes": total_bytes, "strict_dram_bytes": strict_bytes,
                               "achieved": total_bytes / (ms_per_step * 1e-3) / 1e9,
                               "frac": total_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                               "frac_of_8000": total_bytes / (ms_per_step * 1e-3) / 1e9 / 8000.0}}

    # ---- end to end: host buffers, H2D of the step's inputs and D2H of the loss inside the timed region ----
    # Double-buffered input pipeline (what a training loop with a prefetching data loader does): while
    # step i computes, the logits of step i+1 travel host -> device.  Every step still pays its own H2D copy
    # of the pinned logits and its own D2H read of the loss inside the timed loop.  The copy is split in two
    # halves on two copy streams: one stream moves a 12 MB pinned buffer at 17-28 GB/s on this box, two
    # concurrent copies at 53-55 GB/s (PCIe gen5 x16; tools/h2d_bandwidth.py).
    # one pinned allocation for all rotating host buffers: separate pin_memory() calls gave one buffer out of
    # fourteen that copies 2-6x slower than the others on this box (tools/pinned_probe.py)
    pinned_all = torch.empty((len(acts_host),) + tuple(acts_host[0].shape), dtype=acts_host[0].dtype).pin_memory()
    for i, a in enumerate(acts_host):
        pinned_all[i].copy_(a)
    pinned = [pinned_all[i] for i in range(len(acts_host))]
    stage = [torch.empty_like(a) for a in acts_dev[:2]]
    h2d = acts_bytes + wl.labels.nbytes + wl.act_lens.nbytes + wl.label_lens.nbytes
    e2e_loss = []
    copy_streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    copied = [[torch.cuda.Event(), torch.cuda.Event()], [torch.cuda.Event(), torch.cuda.Event()]]
    half = wl.T // 2
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    compute_stream = torch.cuda.current_stream(dev)

    def issue_copy(i):
        d = i % 2
        src = pinned[i % n_rot]
        for k, (cs, lo, hi) in enumerate(((copy_streams[0], 0, half), (copy_streams[1], half, wl.T))):
            with torch.cuda.stream(cs):
                cs.wait_event(consumed[d])                               # the step that last read stage[d] is done
                stage[d][lo:hi].copy_(src[lo:hi], non_blocking=True)      # H2D of step i's logits, frames [lo, hi)
                copied[d][k].record(cs)

    # The loss of every step is copied device -> host (pinned) right behind its kernels and READ on the host one
    # step later, after the next step has been enqueued: the host never idles the GPU while it prepares a call
    # (what a training loop that logs the previous step's loss does).  Every step's loss is read inside the
    # timed region; the last one after the loop.
    e2e_dev = [torch.empty(1, device=dev) for _ in range(2)]
    e2e_host = [torch.empty(1).pin_memory() for _ in range(2)]
    e2e_read = [torch.cuda.Event(), torch.cuda.Event()]
    # The 4-byte D2H goes on its own stream: on the compute stream it would sit in a copy-engine queue behind
    # the prefetch of the next logits and hold the next step's kernels back by up to one copy (0.1-0.2 ms).
    result_stream = torch.cuda.Stream(device=dev)
    step_done = [torch.cuda.Event(), torch.cuda.Event()]

    e2e_debug = [] if os.environ.get("B200CTC_E2E_DEBUG") else None      # developer aid: host time of every step

    def read_loss(i):
        e2e_read[i % 2].synchronize()
        e2e_loss.append(float(e2e_host[i % 2][0]))                      # host read of step i's result

    def e2e_step(i, first):
        d = i % 2
        compute_stream.wait_event(copied[d][0])
        compute_stream.wait_event(copied[d][1])
        b200.ctc_loss_and_grad(stage[d], wl.labels, wl.act_lens, wl.label_lens, grads=grads_dev[i % n_rot],
                               costs=costs, loss_sum=e2e_dev[d])        # labels/lens go host -> device inside the call
        consumed[d].record(compute_stream)
        if world > 1:
            dist.all_reduce(e2e_dev[d])
        step_done[d].record(compute_stream)
        with torch.cuda.stream(result_stream):
            result_stream.wait_event(step_done[d])
            e2e_host[d].copy_(e2e_dev[d], non_blocking=True)             # D2H of the step's result
            e2e_read[d].record(result_stream)
        issue_copy(i + 1)                                               # next step's logits travel while this one computes
        if i > first:
            read_loss(i - 1)

    def e2e_run(steps, first):
        # `first`..`first+steps-1`; the copy of step `first` is issued here, inside the timed region
        issue_copy(first)
        for i in range(first, first + steps):
            t0 = time.perf_counter()
            e2e_step(i, first)
            if e2e_debug is not None:
                e2e_debug.append((time.perf_counter() - t0) * 1e6)
        read_loss(first + steps - 1)
        for cs in copy_streams:
            cs.synchronize()

    for d in (0, 1):
        consumed[d].record(compute_stream)
    e2e_run(n_rot + 2, 0)                     # warm-up: every rotating host buffer has been copied once
    torch.cuda.synchronize()
    e2e_ms = timed(lambda i: e2e_run(args.steps, n_rot + 2) if i == 0 else None, 1) / args.steps
    if e2e_debug is not None and rank == 0:
        print("e2e host us per step:", " ".join("%.0f" % x for x in e2e_debug), file=sys.stderr)
    e2e = {"value": frames * world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
           "pipeline": "double-buffered: H2D of step i+1 (two halves on two copy streams) overlaps the kernels of step i; the loss of step i is copied D2H behind its kernels (own stream) and read on the host after step i+1 is enqueued (one extra prefetch copy per run is also inside the timed region)"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3 if world == 1 else 20), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "per_gpu_batch": wl.B, "frames_per_step_per_gpu": frames,
                   "utterances_per_sec": wl.B * world / (ms_per_step * 1e-3),
                   "l2": "rotating %d acts/grads buffer sets (%.0f MB > L2)" % (n_rot, 2 * n_rot * acts_bytes / 1e6),
                   "parallelism": "utterance-sharded dp%d, scalar loss all-reduce (asynchronous, four steps per message)" % world},
        "host": {"wall_ms_per_step": host_ms_per_step, "cpus": _host_threads()},
        "roofline": roofline, "e2e": e2e, "gpu_launches": 2 * args.steps, "clocks": clocks,
    }
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(wl, acts_host[0].numpy())
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
