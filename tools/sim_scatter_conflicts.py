"""Offline model (no GPU): shared-memory wavefronts of the phase-2 posterior scatter for the C3 label sequences,
with the shipped slot assignment (rank inside the symbol group) and with a conflict-aware one that uses the
freedom inside every symbol group (greedy: sets in order, each label takes a free slot of its group whose bank
is not yet used by its store instruction; if none, the least loaded bank).  One store instruction = one
(side, window, label slot m) set of 32 lanes; label index of lane g (global position group) = 4 g + m."""
import sys
import numpy as np
sys.path.insert(0, ".")
from pytorch_end2end_speech_recognition_b200 import workloads

NS, KX = 8, 16
OWN = 32 * NS - 2 * KX           # positions owned per window
HL = 2 * KX // NS


def sets_of(L, side):
    """label indices of every store instruction (list of arrays), forward side: label of position q = (q-1)/2"""
    S = 2 * L + 1
    JG = (S + NS - 1) // NS
    P = NS * JG
    nw = 1 if P <= 32 * NS else 1 + (P - 32 * NS + OWN - 1) // OWN
    out = []
    for w in range(nw):
        for m in range(4):
            labs = []
            for lane in range(32):
                if w > 0 and lane < HL:
                    continue                                  # halo lane: dump slot
                g = (w * OWN) // NS + lane
                if g >= JG:
                    continue
                q = NS * g + (2 * m if side else 2 * m + 1)
                s = P - 1 - q if side else q
                if 0 <= s < S and s % 2 == 1:
                    labs.append(s >> 1)
            if labs:
                out.append(np.array(labs))
    return out


def wavefronts(sets, slot):
    tot = 0
    for labs in sets:
        banks = slot[labs] % 32
        tot += np.bincount(banks, minlength=32).max()
    return tot, len(sets)


def main():
    wl = workloads.make_lengths_and_labels("C3")
    off = 0
    base_w = aware_w = n_sets = 0
    for b in range(wl.B):
        L = int(wl.label_lens[b]); lab = wl.labels[off:off + L]; off += L
        order = np.lexsort((np.arange(L), lab))
        syms, counts = np.unique(lab, return_counts=True)
        pad = (counts + 3) & ~3
        seg_slot = np.concatenate([[0], np.cumsum(pad)])
        seg_of = {s: u for u, s in enumerate(syms)}
        slot0 = np.empty(L, dtype=np.int64)
        rank = {}
        for i in order:
            u = seg_of[lab[i]]
            slot0[i] = seg_slot[u] + rank.get(u, 0)
            rank[u] = rank.get(u, 0) + 1
        for side in (0, 1):
            sets = sets_of(L, side)
            w0, n = wavefronts(sets, slot0)
            base_w += w0; n_sets += n
            # conflict-aware greedy
            free = [list(range(seg_slot[u], seg_slot[u] + pad[u])) for u in range(len(syms))]
            need = counts.copy()
            slot1 = np.full(L, -1, dtype=np.int64)
            for labs in sets:
                used = np.zeros(32, dtype=np.int64)
                for i in labs:
                    u = seg_of[lab[i]]
                    # keep enough free slots for the group's remaining labels: any free slot is fine (pad >= count)
                    best = min(free[u], key=lambda s: (used[s % 32], s))
                    free[u].remove(best)
                    slot1[i] = best
                    used[best % 32] += 1
            w1, _ = wavefronts(sets, slot1)
            aware_w += w1
    print("C3: %d store instructions per frame (both sides, all utterances)" % n_sets)
    print("wavefronts per store instruction: rank-in-group slots %.2f, conflict-aware slots %.2f" % (base_w / n_sets, aware_w / n_sets))


if __name__ == "__main__":
    main()
