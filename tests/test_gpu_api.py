"""GPU tests of the reference-facing surface: the warpctc_pytorch names the reference imports
(models/pytorch_v3/ctc/ctc.py:11,30,35,39-45,69) behave as that code expects."""
import numpy as np
import pytest
import torch

import pytorch_end2end_speech_recognition_b200 as b200
from oracle import ctc_ref
from pytorch_end2end_speech_recognition_b200 import workloads

pytestmark = pytest.mark.gpu


def small():
    wl = workloads.make_lengths_and_labels(None, B=4, T=30, V=9, Lmax=8, kind="var", seed=3)
    acts = workloads.make_acts(wl)
    c_ref, g_ref = ctc_ref.ctc_cost_and_grad(acts.numpy(), wl.labels, wl.act_lens, wl.label_lens)
    return wl, acts, c_ref, g_ref


def as_i32(x):
    return torch.from_numpy(np.asarray(x, dtype=np.int32))


def test_gpu_ctc_fills_grads_and_costs_in_place():
    wl, acts, c_ref, g_ref = small()
    a = acts.cuda()
    grads = torch.zeros(a.size()).type_as(a)                     # ctc.py:36
    costs = torch.zeros(wl.B).cpu()                              # ctc.py:38
    st = b200.gpu_ctc(a, grads, as_i32(wl.labels), as_i32(wl.label_lens), as_i32(wl.act_lens), wl.B, costs)
    assert st == 0 and not costs.is_cuda
    assert np.allclose(costs.numpy(), c_ref, rtol=1e-5)
    assert np.max(np.abs(grads.cpu().numpy() - g_ref)) < 1e-4


def test_reference_style_subclass_of_CTC():
    """The reference overrides forward and inherits backward (ctc.py:30-52)."""
    from torch.autograd import Variable

    class _CTC(b200._CTC):
        @staticmethod
        def forward(ctx, acts, labels, act_lens, label_lens, size_average=False):
            is_cuda = True if acts.is_cuda else False
            acts = acts.contiguous()
            loss_func = b200.gpu_ctc if is_cuda else b200.cpu_ctc
            grads = torch.zeros(acts.size()).type_as(acts)
            minibatch_size = acts.size(1)
            costs = torch.zeros(minibatch_size).cpu()
            loss_func(acts, grads, labels, label_lens, act_lens, minibatch_size, costs)
            if size_average:
                costs = torch.FloatTensor([costs.mean()])
            else:
                costs = torch.FloatTensor([costs.sum()])
            ctx.grads = Variable(grads)
            return costs

    wl, acts, c_ref, g_ref = small()
    logits = acts.transpose(0, 1).contiguous().cuda().requires_grad_(True)   # [B,T,V] like the model output
    loss = _CTC.apply(logits.transpose(0, 1), as_i32(wl.labels), as_i32(wl.act_lens), as_i32(wl.label_lens), False)
    assert loss.shape == (1,) and not loss.is_cuda
    loss = loss.cuda() / wl.B                                    # ctc.py:323-326
    loss.backward()
    assert abs(float(loss.detach()) * wl.B - c_ref.sum()) < 1e-5 * c_ref.sum()
    got = logits.grad.transpose(0, 1).cpu().numpy() * wl.B
    assert np.max(np.abs(got - g_ref)) < 1e-4


@pytest.mark.parametrize("size_average,length_average", [(False, False), (True, False), (False, True)])
def test_CTCLoss_module_reductions(size_average, length_average):
    wl, acts, c_ref, g_ref = small()
    a = acts.cuda().requires_grad_(True)
    crit = b200.CTCLoss(size_average=size_average, length_average=length_average)
    loss = crit(a, as_i32(wl.labels), as_i32(wl.act_lens), as_i32(wl.label_lens))
    assert loss.shape == (1,) and not loss.is_cuda
    (loss * 2.0).sum().backward()                                # grad_output = 2
    div = float(wl.act_lens.sum()) if length_average else (wl.B if size_average else 1.0)
    assert abs(float(loss) - c_ref.sum() / div) < 1e-5 * c_ref.sum() / div
    assert np.max(np.abs(a.grad.cpu().numpy() - 2.0 * g_ref / div)) < 1e-4


def test_device_loss_api_and_reductions():
    wl, acts, c_ref, g_ref = small()
    for red in ("sum", "mean", "none"):
        a = acts.cuda().requires_grad_(True)
        out = b200.ctc_loss(a, wl.labels, wl.act_lens, wl.label_lens, reduction=red)
        assert out.is_cuda
        out.sum().backward()
        div = wl.B if red == "mean" else 1.0
        assert abs(float(out.sum()) - c_ref.sum() / div) < 1e-5 * c_ref.sum()
        assert np.max(np.abs(a.grad.cpu().numpy() - g_ref / div)) < 1e-4


def test_invalid_inputs_raise_runtime_error():
    """train_step's guard catches RuntimeError and skips the mini-batch
    (utils/training/training_loop.py:69-76): every bad input must raise, never abort."""
    wl, acts, _, _ = small()
    a = acts.cuda()
    good = (wl.labels.copy(), wl.act_lens.copy(), wl.label_lens.copy())
    bad = good[0].copy(); bad[0] = wl.V                 # label out of range
    with pytest.raises(RuntimeError):
        b200.ctc_loss_and_grad(a, bad, good[1], good[2])
    bad = good[0].copy(); bad[0] = 0                    # blank inside the labels
    with pytest.raises(RuntimeError):
        b200.ctc_loss_and_grad(a, bad, good[1], good[2])
    bad = good[1].copy(); bad[0] = wl.T + 1             # act_len > T
    with pytest.raises(RuntimeError):
        b200.ctc_loss_and_grad(a, good[0], bad, good[2])
    with pytest.raises(RuntimeError):                   # sum(label_lens) != len(labels)
        b200.ctc_loss_and_grad(a, good[0][:-1], good[1], good[2])
    with pytest.raises(RuntimeError):                   # wrong dtype
        b200.ctc_loss_and_grad(a.double(), *good)
    with pytest.raises(RuntimeError):                   # blank out of range
        b200.ctc_loss_and_grad(a, *good, blank=wl.V)
    # and the engine still works afterwards
    c, _, _ = b200.ctc_loss_and_grad(a, *good)
    assert torch.isfinite(c).all()


def test_ctc_loss_from_padded_matches_the_reference_call_site():
    """ctc.py:299-326: ys + 1, concatenate, time-major view, sum of costs / len(xs); gradient w.r.t. logits."""
    rng = np.random.RandomState(21)
    B, T, V, Lmax = 5, 40, 11, 9
    logits = torch.randn(B, T, V, device="cuda", requires_grad=True)
    y_lens = rng.randint(1, Lmax + 1, size=B).astype(np.int32)
    ys = np.zeros((B, Lmax), dtype=np.int64)
    for b in range(B):
        ys[b, :y_lens[b]] = rng.randint(0, V - 1, size=y_lens[b])      # 0-based, blank offset not applied yet
    x_lens = np.array([40, 38, 33, 30, 25], dtype=np.int32)
    loss = b200.ctc_loss_from_padded(logits, ys, x_lens, y_lens)
    loss.backward()
    flat = np.concatenate([ys[b, :y_lens[b]] + 1 for b in range(B)]).astype(np.int32)
    acts = logits.detach().cpu().numpy().transpose(1, 0, 2)
    c_ref, g_ref = ctc_ref.ctc_cost_and_grad(acts, flat, x_lens, y_lens)
    assert abs(float(loss) - c_ref.sum() / B) < 1e-5 * c_ref.sum()
    assert np.max(np.abs(logits.grad.cpu().numpy().transpose(1, 0, 2) - g_ref / B)) < 1e-4
