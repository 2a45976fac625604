import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytorch_end2end_speech_recognition_b200 as b200
from pytorch_end2end_speech_recognition_b200 import ctc as ctc_mod
from oracle import ctc_ref
rng = np.random.RandomState(7)
for it in range(25):
    B, T, V = rng.randint(1, 7), rng.randint(1, 70), rng.randint(2, 40)
    act_lens = rng.randint(1, T + 1, size=B); act_lens[0] = T
    labels, label_lens = [], []
    for b in range(B):
        L = rng.randint(0, act_lens[b] + 1)
        lab = rng.randint(1, V, size=L)
        for j in range(1, L):
            if rng.uniform() < 0.15:
                lab[j] = lab[j - 1]
        if it % 5 != 0:
            while L + ctc_ref.count_repeats(lab[:L]) > act_lens[b]:
                L -= 1
        labels.append(lab[:L]); label_lens.append(L)
    flat = np.concatenate(labels).astype(np.int32)
    acts = (rng.randn(T, B, V) * rng.choice([0.3, 1.0, 4.0])).astype(np.float32)
    costs, loss, grads = b200.ctc_loss_and_grad(torch.from_numpy(acts).cuda(), flat, act_lens, label_lens)
    g = grads.cpu().numpy(); c = costs.cpu().numpy()
    c_ref, g_ref = ctc_ref.ctc_cost_and_grad(acts, flat, act_lens, label_lens)
    err = np.abs(g - g_ref)
    print('it', it, 'B,T,V', B, T, V, 'max err %.2e' % err.max(), 'fb', ctc_mod.last_fallbacks())
    if err.max() > 1e-4:
        for b in range(B):
            e = err[:, b]
            if e.max() > 1e-4:
                fr = np.where(e.max(1) > 1e-4)[0]
                off = int(np.sum(label_lens[:b]))
                print('   b', b, 'T_b', act_lens[b], 'L_b', label_lens[b], 'repeats', ctc_ref.count_repeats(labels[b]), 'cost', c[b], c_ref[b],
                      'bad frames', fr[:20], 'n', len(fr))
                t = fr[0]
                print('   labels', labels[b][:12], ' row', t, 'got', np.round(g[t, b, :10], 4), 'ref', np.round(g_ref[t, b, :10], 4))
    if it == 14:
        b = 4
        off = int(np.sum(label_lens[:b])); Lb = label_lens[b]; Tb = act_lens[b]
        lab = flat[off:off+Lb]
        a1 = np.ascontiguousarray(acts[:Tb, b:b+1])
        c1, l1, g1 = b200.ctc_loss_and_grad(torch.from_numpy(a1).cuda(), lab, [Tb], [Lb])
        g1 = g1.cpu().numpy(); cr, gr = ctc_ref.ctc_cost_and_grad(a1, lab, [Tb], [Lb])
        print('   single-utterance rerun: max err %.2e' % np.abs(g1-gr).max(), 'labels', lab.tolist())
        for (t, k) in zip(*np.where(np.abs(g1[:,0]-gr[:,0]) > 1e-4)):
            print('     t', t, 'col', k, 'got %.4f ref %.4f' % (g1[t,0,k], gr[t,0,k]))
        # expected path
        print('   softmax-occupancy ref at bad frames: argmin of ref grad per frame', [int(np.argmin(gr[t,0])) for t in range(Tb)])
