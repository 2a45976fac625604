"""CPU tests of the oracle itself: known-answer cases, the committed golden vectors, and an
independent implementation (torch.nn.functional.ctc_loss, CPU float64)."""
import os

import numpy as np
import pytest
import torch

from oracle import ctc_ref
from oracle.ctc_cpu import ctc_cpu


def lsm(x):
    return ctc_ref.log_softmax(np.asarray(x, dtype=np.float64), axis=-1)


def test_kat_T1_L0():
    acts = np.array([[[0.3, -1.2, 2.0]]])          # [T=1,B=1,V=3]
    c, g = ctc_ref.ctc_cost_and_grad(acts, [], [1], [0])
    lp = lsm(acts[0, 0])
    assert np.isclose(c[0], -lp[0], rtol=1e-14)
    y = np.exp(lp)
    expect = y.copy(); expect[0] -= 1.0           # all posterior mass on blank
    assert np.allclose(g[0, 0], expect, atol=1e-14)


def test_kat_T1_L1():
    acts = np.array([[[0.3, -1.2, 2.0]]])
    c, _ = ctc_ref.ctc_cost_and_grad(acts, [2], [1], [1])
    assert np.isclose(c[0], -lsm(acts[0, 0])[2], rtol=1e-14)


def test_kat_T2_L1_three_paths():
    rng = np.random.RandomState(0)
    acts = rng.randn(2, 1, 4)
    lp = lsm(acts[:, 0])
    a = 3
    p = np.exp(lp[0, a] + lp[1, 0]) + np.exp(lp[0, 0] + lp[1, a]) + np.exp(lp[0, a] + lp[1, a])
    c, _ = ctc_ref.ctc_cost_and_grad(acts, [a], [2], [1])
    assert np.isclose(c[0], -np.log(p), rtol=1e-13)


def test_kat_uniform_path_count():
    # uniform logits, V=2, label "1": p = (#valid paths) / 2^T ; valid = 0*1+0* -> T(T+1)/2 paths
    for T in (1, 2, 3, 5, 8):
        acts = np.zeros((T, 1, 2))
        c, _ = ctc_ref.ctc_cost_and_grad(acts, [1], [T], [1])
        assert np.isclose(c[0], -np.log(T * (T + 1) / 2 / 2.0 ** T), rtol=1e-12)


def test_kat_repeated_label():
    rng = np.random.RandomState(1)
    acts = rng.randn(3, 1, 3)
    # "aa" with T=2 has no valid alignment
    c, g = ctc_ref.ctc_cost_and_grad(acts[:2], [1, 1], [2], [2])
    assert np.isinf(c[0]) and c[0] > 0 and np.all(g == 0)
    # "aa" with T=3: exactly one path a-blank-a
    c, _ = ctc_ref.ctc_cost_and_grad(acts, [1, 1], [3], [2])
    lp = lsm(acts[:, 0])
    assert np.isclose(c[0], -(lp[0, 1] + lp[1, 0] + lp[2, 1]), rtol=1e-13)


def test_kat_L_equals_T():
    rng = np.random.RandomState(2)
    acts = rng.randn(4, 1, 5)
    lab = [1, 2, 3, 4]
    c, _ = ctc_ref.ctc_cost_and_grad(acts, lab, [4], [4])
    lp = lsm(acts[:, 0])
    assert np.isclose(c[0], -sum(lp[t, lab[t]] for t in range(4)), rtol=1e-13)


def test_padding_rows_zero_and_ignored():
    rng = np.random.RandomState(3)
    acts = rng.randn(10, 2, 6)
    labels = [1, 2, 3, 4, 5]
    c1, g1 = ctc_ref.ctc_cost_and_grad(acts, labels, [10, 6], [3, 2])
    acts2 = acts.copy(); acts2[6:, 1] = rng.randn(4, 6) * 10
    c2, g2 = ctc_ref.ctc_cost_and_grad(acts2, labels, [10, 6], [3, 2])
    assert np.all(g1[6:, 1] == 0) and np.all(g2[6:, 1] == 0)
    assert np.array_equal(c1, c2) and np.array_equal(g1, g2)


def test_gradient_rows_sum_to_zero_and_stability():
    rng = np.random.RandomState(4)
    acts = rng.randn(30, 3, 8) * 50          # logits +-50 and beyond
    c, g = ctc_ref.ctc_cost_and_grad(acts, [1, 2, 2, 3, 7, 7, 1], [30, 20, 25], [3, 2, 2])
    assert np.all(np.isfinite(c)) and np.all(np.isfinite(g))
    assert np.max(np.abs(g.sum(axis=2))) < 1e-12


def test_permutation_invariance():
    rng = np.random.RandomState(5)
    acts = rng.randn(12, 3, 5)
    labs = [[1, 2], [3], [4, 4, 1]]
    al, ll = [12, 9, 11], [2, 1, 3]
    c, g = ctc_ref.ctc_cost_and_grad(acts, sum(labs, []), al, ll)
    perm = [2, 0, 1]
    c2, g2 = ctc_ref.ctc_cost_and_grad(acts[:, perm], sum([labs[i] for i in perm], []),
                                       [al[i] for i in perm], [ll[i] for i in perm])
    assert np.allclose(c[perm], c2, rtol=1e-14) and np.allclose(g[:, perm], g2, atol=1e-14)


def test_finite_difference():
    rng = np.random.RandomState(6)
    acts = rng.randn(7, 1, 4)
    lab = [1, 3, 3]
    c, g = ctc_ref.ctc_cost_and_grad(acts, lab, [7], [3])
    eps = 1e-6
    for (t, k) in [(0, 0), (3, 3), (6, 1), (2, 2)]:
        ap = acts.copy(); ap[t, 0, k] += eps
        am = acts.copy(); am[t, 0, k] -= eps
        fd = (ctc_ref.ctc_cost_and_grad(ap, lab, [7], [3])[0][0] - ctc_ref.ctc_cost_and_grad(am, lab, [7], [3])[0][0]) / (2 * eps)
        assert abs(fd - g[t, 0, k]) < 1e-7


def test_against_torch_fp64_random():
    rng = np.random.RandomState(7)
    for _ in range(15):
        B, T, V = rng.randint(1, 5), rng.randint(1, 40), rng.randint(2, 9)
        act_lens = rng.randint(1, T + 1, size=B); act_lens[0] = T
        labels, label_lens = [], []
        for b in range(B):
            L = rng.randint(0, act_lens[b] // 2 + 1)
            lab = rng.randint(1, V, size=L)
            while L + ctc_ref.count_repeats(lab[:L]) > act_lens[b]:
                L -= 1
            labels.append(lab[:L]); label_lens.append(L)
        flat = np.concatenate(labels).astype(np.int64)
        acts = rng.randn(T, B, V)
        c, g = ctc_ref.ctc_cost_and_grad(acts, flat, act_lens, label_lens)
        a = torch.tensor(acts, requires_grad=True)
        tc = torch.nn.functional.ctc_loss(torch.log_softmax(a, 2), torch.tensor(flat), torch.tensor(act_lens),
                                          torch.tensor(label_lens), blank=0, reduction="none")
        tc.sum().backward()
        assert np.allclose(c, tc.detach().numpy(), rtol=1e-12)
        assert np.allclose(g, a.grad.numpy(), atol=1e-12)


def test_ctc_golden_vectors(golden_dir):
    z = np.load(os.path.join(golden_dir, "ctc_golden.npz"))
    for i in range(int(z["n_cases"])):
        c, g = ctc_ref.ctc_cost_and_grad(z["acts_%d" % i], z["labels_%d" % i], z["act_lens_%d" % i], z["label_lens_%d" % i])
        assert np.allclose(c, z["costs_%d" % i], rtol=1e-10), i
        assert np.max(np.abs(g - z["grads_%d" % i])) < 1e-6, i       # fixture stored in float32
        # C++ restatement, fp64 instantiation
        c2, g2 = ctc_cpu(z["acts_%d" % i], z["labels_%d" % i], z["act_lens_%d" % i], z["label_lens_%d" % i], precision="f64")
        assert np.allclose(c2, z["costs_%d" % i], rtol=1e-6), i
        assert np.max(np.abs(g2 - z["grads_%d" % i])) < 1e-6, i


def test_reference_chainer_golden_vectors(golden_dir):
    """PINS THE ORACLE TO THE REFERENCE: costs and gradients computed by the reference's own in-tree
    CTC (models/chainer/ctc/ctc_loss_from_chainer.py:214-304, float32, imported from /root/reference
    by tests/golden/make_golden.py).  The reference code is a float32 log-space recursion, so its
    gradients differ from exact arithmetic by up to ~1.3e-4 at T=96; tolerances: cost 1e-5 relative,
    gradient 1e-4 absolute up to T=64 and 5e-4 beyond."""
    z = np.load(os.path.join(golden_dir, "ctc_reference_golden.npz"))
    assert int(z["n_cases"]) >= 8
    for i in range(int(z["n_cases"])):
        acts = z["acts_%d" % i]
        args = (acts, z["labels_%d" % i], z["act_lens_%d" % i], z["label_lens_%d" % i])
        tol = 1e-4 if acts.shape[0] <= 64 else 5e-4
        c, g = ctc_ref.ctc_cost_and_grad(*args)
        assert np.allclose(c, z["costs_%d" % i], rtol=1e-5), i
        assert np.max(np.abs(g - z["grads_%d" % i])) < tol, i
        c2, g2 = ctc_cpu(*args, precision="f64")
        assert np.allclose(c2, z["costs_%d" % i], rtol=1e-5), i
        assert np.max(np.abs(g2 - z["grads_%d" % i])) < tol, i
        c3, g3 = ctc_cpu(*args, precision="f32")       # the timed CPU baseline
        assert np.allclose(c3, z["costs_%d" % i], rtol=1e-4), i
        assert np.max(np.abs(g3 - z["grads_%d" % i])) < 2e-3, i


def test_cpp_restatement_fp32_close():
    # the float instantiation (warp-ctc's ProbT=float) is only accurate to ~1e-3: loose bound
    rng = np.random.RandomState(8)
    acts = rng.randn(60, 4, 12).astype(np.float32)
    labels = rng.randint(1, 12, size=40).astype(np.int32)
    c64, g64 = ctc_ref.ctc_cost_and_grad(acts, labels, [60, 50, 45, 40], [10, 10, 10, 10])
    c, g = ctc_cpu(acts, labels, [60, 50, 45, 40], [10, 10, 10, 10], precision="f32")
    assert np.allclose(c, c64, rtol=1e-5)
    assert np.max(np.abs(g - g64)) < 2e-3


def test_greedy_golden_vectors(golden_dir):
    z = np.load(os.path.join(golden_dir, "greedy_golden.npz"))
    for i in range(int(z["n_cases"])):
        hyps = ctc_ref.greedy_decode(z["logits_%d" % i], z["x_lens_%d" % i], blank=0)
        lens = z["hyp_lens_%d" % i]
        flat = z["hyp_flat_%d" % i]
        assert [len(h) for h in hyps] == lens.tolist(), i
        assert np.array_equal(np.concatenate(hyps) if len(flat) else np.zeros(0, np.int64), flat), i


def test_greedy_kat():
    V = 4
    def onehot(seq):
        x = np.zeros((1, len(seq), V), np.float32)
        for t, k in enumerate(seq):
            x[0, t, k] = 1
        return x
    assert ctc_ref.greedy_decode(onehot([0, 0, 0]), [3])[0].tolist() == []
    assert ctc_ref.greedy_decode(onehot([0, 1, 1, 0, 1, 2, 2, 0]), [8])[0].tolist() == [1, 1, 2]
    assert ctc_ref.greedy_decode(onehot([1, 1, 2, 3]), [2])[0].tolist() == [1]          # truncation
    assert ctc_ref.greedy_decode(np.zeros((1, 5, V), np.float32), [5])[0].tolist() == []  # ties -> index 0


def test_reductions():
    costs = np.array([1.0, 2.0, 6.0])
    assert ctc_ref.reduce_costs(costs, [4, 4, 4]) == 9.0
    assert ctc_ref.reduce_costs(costs, [4, 4, 4], size_average=True) == 3.0
    assert ctc_ref.reduce_costs(costs, [4, 4, 4], length_average=True) == 0.75


def test_call_site_restatement_matches_torch_autograd():
    """oracle.ctc_ref.ctc_loss_call_site (ctc.py:299-337 + criterion.py:51-80 in fp64) against the same
    computation written with torch ops on the CPU in fp64 (log_softmax, F.ctc_loss, the label-smoothing sum)."""
    import torch
    rng = np.random.RandomState(0)
    B, T, V, Lmax = 4, 30, 9, 7
    logits = rng.randn(B, T, V)
    y_lens = rng.randint(1, Lmax + 1, size=B)
    ys = np.zeros((B, Lmax), np.int64)
    for b in range(B):
        ys[b, :y_lens[b]] = rng.randint(0, V - 1, size=y_lens[b])
    x_lens = np.array([30, 28, 22, 19])
    for temp, ls in [(1.0, 0.0), (2.0, 0.0), (1.0, 0.1), (1.7, 0.2)]:
        loss, g = ctc_ref.ctc_loss_call_site(logits, ys, x_lens, y_lens, temp, ls)
        x = torch.tensor(logits, requires_grad=True)
        lp = torch.log_softmax(x / temp, dim=2)
        flat = torch.tensor(np.concatenate([ys[b, :y_lens[b]] + 1 for b in range(B)]))
        ref = torch.nn.functional.ctc_loss(lp.transpose(0, 1), flat, torch.tensor(x_lens), torch.tensor(y_lens),
                                           blank=0, reduction="sum") / B
        if ls > 0:
            ref = ref * (1 - ls) + sum([(-(ls / V) * lp[b, :x_lens[b]]).sum() for b in range(B)]) / B
        ref.backward()
        assert abs(float(ref.detach()) - loss) < 1e-12 * abs(loss)
        assert np.max(np.abs(x.grad.numpy() - g)) < 1e-12


def test_edit_distance_oracle_matches_the_reference_function(golden_dir):
    """oracle/eval_ref.compute_wer against the outputs of the reference's own compute_wer
    (tests/golden/edit_distance_golden.npz); where the reference raised (ok = 0) only the distance identity
    sub + ins + del == distance and the length identity are checked."""
    from oracle import eval_ref
    z = np.load(os.path.join(golden_dir, "edit_distance_golden.npz"))
    assert int(z["ok"].sum()) >= 100
    for b in range(len(z["ok"])):
        ref, hyp = z["refs"][b, :z["ref_lens"][b]], z["hyps"][b, :z["hyp_lens"][b]]
        got = eval_ref.compute_wer(list(ref), list(hyp))
        if z["ok"][b]:
            assert got == tuple(int(x) for x in z["out"][b]), b
        d, sub, ins, dele = got
        assert d == sub + ins + dele and len(ref) - sub - dele == len(hyp) - sub - ins


def test_beam_search_oracle_matches_the_reference_decoder(golden_dir):
    from oracle import beam_ref
    z = np.load(os.path.join(golden_dir, "beam_golden.npz"))
    for i in range(int(z["n_cases"])):
        lp, x_lens, W = z["log_probs_%d" % i], z["x_lens_%d" % i], int(z["beam_%d" % i])
        offs = np.concatenate([[0], np.cumsum(z["hyp_lens_%d" % i])])
        assert np.array_equal(z["hyp_lens_%d" % i], z["hyp32_lens_%d" % i])          # fp32 run of the reference: same result
        assert np.array_equal(z["hyp_flat_%d" % i], z["hyp32_flat_%d" % i])
        for b in range(lp.shape[0]):
            if lp.shape[1] * lp.shape[2] * W > 20000:
                continue                                                            # keep the CPU suite short
            hyp, _ = beam_ref.beam_search(lp[b], x_lens[b], W)
            assert hyp == list(z["hyp_flat_%d" % i][offs[b]:offs[b + 1]]), (i, b)
