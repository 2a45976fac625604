"""Developer diagnostic (GPU box): runs a ladder of cases through the CUDA path and prints the
error against the oracle plus how many utterances fell back to the safe lattice."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pytorch_end2end_speech_recognition_b200 as b200
from pytorch_end2end_speech_recognition_b200 import workloads, ctc as ctc_mod
from oracle import ctc_ref
from oracle.ctc_cpu import ctc_cpu

def run(name, wl, acts, oracle='numpy'):
    a = acts.cuda()
    torch.cuda.synchronize()
    t0 = time.time()
    costs, loss, grads = b200.ctc_loss_and_grad(a, wl.labels, wl.act_lens, wl.label_lens)
    torch.cuda.synchronize()
    dt = time.time() - t0
    fb = ctc_mod.last_fallbacks()
    if oracle == 'numpy':
        c_ref, g_ref = ctc_ref.ctc_cost_and_grad(acts.numpy(), wl.labels, wl.act_lens, wl.label_lens)
    else:
        c_ref, g_ref = ctc_cpu(acts.numpy(), wl.labels, wl.act_lens, wl.label_lens, precision='f64')
    c = costs.cpu().numpy(); g = grads.cpu().numpy()
    fin = np.isfinite(c_ref)
    rel = np.max(np.abs(c[fin]-c_ref[fin])/np.maximum(np.abs(c_ref[fin]),1e-3)) if fin.any() else 0
    gerr = np.abs(g-g_ref)
    worst = np.unravel_index(np.argmax(gerr), gerr.shape)
    print('%-28s fallbacks=%s cost_rel=%.2e grad_abs=%.2e at %s nan=%d  (%.1f ms incl. launch)' % (
        name, fb, rel, gerr.max(), worst, int(np.isnan(g).sum()), dt*1e3), flush=True)
    return rel, gerr.max()

cases = [
    ('T1 L0', dict(B=1,T=1,V=3,Lmax=0)), ('T1 L1', dict(B=1,T=1,V=3,Lmax=1)), ('T2 L1', dict(B=1,T=2,V=4,Lmax=1)),
    ('T5 L2', dict(B=2,T=5,V=4,Lmax=2)), ('T9 L4', dict(B=3,T=9,V=5,Lmax=4)), ('T17 L7', dict(B=2,T=17,V=6,Lmax=7)),
    ('T40 L15', dict(B=4,T=40,V=9,Lmax=15)), ('T100 L31', dict(B=4,T=100,V=30,Lmax=31)), ('T100 L32', dict(B=4,T=100,V=30,Lmax=32)),
    ('T130 L60', dict(B=4,T=130,V=30,Lmax=60)), ('T300 L64 V33', dict(B=3,T=300,V=33,Lmax=64)),
    ('T300 L130 V62', dict(B=3,T=300,V=62,Lmax=130)), ('T200 L40 V300', dict(B=3,T=200,V=300,Lmax=40)),
    ('T800 L400', dict(B=4,T=800,V=30,Lmax=400)),
]
for name, kw in cases:
    for seed in (1, 2):
        wl = workloads.make_lengths_and_labels(None, kind='var', seed=seed, **kw) if kw['Lmax'] > 0 else \
             workloads.Workload('x', kw['T'], kw['B'], kw['V'], np.zeros(0,np.int32), np.zeros(kw['B'],np.int32), np.full(kw['B'],kw['T'],np.int32), seed)
        acts = workloads.make_acts(wl)
        try:
            run(name + ' s%d' % seed, wl, acts)
        except Exception as e:
            print(name, 'EXC', e, flush=True)
if '--full' in sys.argv:
    for key in ['C1','C2','C3','C4','C5']:
        wl = workloads.make_lengths_and_labels(key); acts = workloads.make_acts(wl)
        run(key, wl, acts, oracle='cpp')
