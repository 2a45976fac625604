"""Developer tool (GPU box): timeline of the lattice kernel's CTA 0 (the longest utterance).

Builds a -DB200CTC_TRACE variant of the library (lib/libb200ctc_trace.so), runs one C3-shaped call
through it and prints, per warp role, where the cycles of a chunk go:
  lattice warps: [2] phase-1 chunk start -> [3] frames done -> (barrier, halo) -> [2] ...; [4] midpoint;
                 [13] phase-2 chunk start -> [14] frames + posteriors done -> (barrier, halo) -> [13] ...
  helper warps:  [7] phase-2 chunk start -> [8] prefetch for the next chunk issued -> [9] previous chunk's
                 frame reduced -> (cp.async wait, barrier) -> [7] ...
Usage: python tools/trace_lattice.py [C3] [--build-only]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_end2end_speech_recognition_b200 import build as b  # noqa: E402

LIB = os.environ.get("B200CTC_TRACE_LIB") or os.path.join(b.LIB_DIR, "libb200ctc_trace.so")


def main():
    key = next((a for a in sys.argv[1:] if a.startswith("C")), "C3")
    if not os.path.exists(LIB) or "--rebuild" in sys.argv or "--build-only" in sys.argv:
        b.build_library(extra_flags=["-DB200CTC_TRACE"], lib_path=LIB)
    if "--build-only" in sys.argv:
        return
    import numpy as np
    import torch
    b.LIB_PATH = LIB                      # make the package load the trace variant ...
    os.environ["B200CTC_LIB"] = LIB       # ... as it is (no staleness rebuild without -DB200CTC_TRACE)
    import pytorch_end2end_speech_recognition_b200 as eng
    from pytorch_end2end_speech_recognition_b200 import _lib, workloads
    wl = workloads.make_lengths_and_labels(key)
    acts = workloads.make_acts(wl).cuda()
    lib = _lib.load()
    cap = 4096
    host = (ctypes.c_longlong * (64 * cap))()
    cnt = (ctypes.c_int * 64)()
    cta = next((int(a[6:]) for a in sys.argv[1:] if a.startswith("--cta=")), 0)
    assert lib.b200ctc_debug_set_trace_cta(cta) == 0
    for _ in range(3):
        eng.ctc_loss_and_grad(acts, wl.labels, wl.act_lens, wl.label_lens)
        torch.cuda.synchronize()
        assert lib.b200ctc_debug_read_trace(host, cnt) == 0
    ev = np.frombuffer(host, dtype=np.int64).reshape(64, cap)
    counts = np.frombuffer(cnt, dtype=np.int32)
    t_all = [ev[w, :counts[w]] >> 8 for w in range(64)]
    tag_all = [ev[w, :counts[w]] & 0xff for w in range(64)]
    t0 = min(int(t[0]) for t in t_all if len(t))
    t1 = max(int(t[-1]) for t in t_all if len(t))
    print("CTA %d: %d cycles traced" % (cta, t1 - t0))
    for w in range(64):
        t, tag = t_all[w], tag_all[w]
        if not len(t):
            continue
        segs = {}
        for i in range(1, len(t)):
            k = (int(tag[i - 1]), int(tag[i]))
            d = int(t[i] - t[i - 1])
            s = segs.setdefault(k, [0, 0])
            s[0] += d
            s[1] += 1
        span = int(t[-1] - t[0])
        desc = ", ".join("%d->%d: %d x %.0f" % (k[0], k[1], v[1], v[0] / v[1]) for k, v in sorted(segs.items()) if v[0] > 0.01 * span)
        print("warp %2d  first %7d  span %7d  %s" % (w, int(t[0]) - t0, span, desc))
    # lattice warps, per phase and per quarter of the phase: run / publish / barrier wait / import (cycles per chunk)
    for w in range(63):
        t, tag = t_all[w], tag_all[w]
        if not len(t) or 2 not in tag:
            continue
        for phase, start_tag, done_tag in ((1, 2, 3), (2, 13, 14)):
            idx = [i for i in range(len(t)) if tag[i] == start_tag]
            rows = []
            for a, i in enumerate(idx):
                if i + 3 < len(t) and tag[i + 1] == done_tag and tag[i + 2] == 30 and tag[i + 3] == 31:
                    nxt = int(t[i + 4] - t[i + 3]) if i + 4 < len(t) and tag[i + 4] == start_tag else 0
                    rows.append((int(t[i + 1] - t[i]), int(t[i + 2] - t[i + 1]), int(t[i + 3] - t[i + 2]), nxt))
            if not rows:
                continue
            r = np.array(rows, dtype=np.float64)
            q = max(1, len(r) // 4)
            parts = ["q%d run %4.0f pub %3.0f wait %4.0f imp %3.0f" % ((k,) + tuple(r[k * q:(k + 1) * q].mean(axis=0))) for k in range(4)]
            print("warp %2d phase %d: %s" % (w, phase, " | ".join(parts)))
    if "--ctas" in sys.argv:           # per-CTA wall time (ns, %globaltimer): who finishes last
        nb = min(wl.B, 2048)
        tt = (ctypes.c_longlong * (2 * nb))()
        assert lib.b200ctc_debug_read_cta_times(tt, nb) == 0
        tt = np.frombuffer(tt, dtype=np.int64).reshape(nb, 2)
        order = np.argsort(-(wl.act_lens.astype(np.int64) * (2 * wl.label_lens + 1)), kind="stable")
        t_first = tt[:, 0].min()
        dur = tt[:, 1] - tt[:, 0]
        print("kernel span %.1f us; CTA start spread %.1f us" % ((tt[:, 1].max() - t_first) / 1e3, (tt[:, 0].max() - t_first) / 1e3))
        for i in np.argsort(-tt[:, 1])[:8]:
            bb = order[i]
            print("  CTA %3d (T=%d L=%d): start +%.1f us, runs %.1f us, ends +%.1f us" % (i, wl.act_lens[bb], wl.label_lens[bb], (tt[i, 0] - t_first) / 1e3, dur[i] / 1e3, (tt[i, 1] - t_first) / 1e3))
        ls = wl.label_lens[order[:nb]]
        for lo, hi in ((0, 128), (128, 240), (240, 252), (252, 352), (352, 376), (376, 464), (464, 2000)):
            sel = (ls >= lo) & (ls < hi)
            if sel.any():
                print("  L in [%d,%d): %d CTAs, run time mean %.1f us max %.1f us" % (lo, hi, sel.sum(), dur[sel].mean() / 1e3, dur[sel].max() / 1e3))
    if "--raw" in sys.argv:            # raw event sequences (tag@cycle) of every warp around the middle of phase 1 and of phase 2
        for lo_frac in (0.2, 0.7):
            lo = t0 + int((t1 - t0) * lo_frac)
            print("---- events in [%d, %d)" % (lo - t0, lo - t0 + 8000))
            for w in range(64):
                t, tag = t_all[w], tag_all[w]
                sel = [(int(x) - t0, int(g)) for x, g in zip(t, tag) if lo <= x < lo + 8000]
                if sel:
                    print("warp %2d: %s" % (w, " ".join("%d@%d" % (g, x - (lo - t0)) for x, g in sel)))


if __name__ == "__main__":
    main()
