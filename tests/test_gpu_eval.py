"""GPU tests of the evaluation kernels: batched edit distance with error counts against
the reference's own compute_wer (golden fixture) and the numpy oracle, the posterior softmax, and the batched
decode-and-score call."""
import os

import numpy as np
import pytest
import torch

import pytorch_end2end_speech_recognition_b200 as b200
from oracle import ctc_ref, eval_ref
from pytorch_end2end_speech_recognition_b200 import evaluation

pytestmark = pytest.mark.gpu


def test_edit_distance_matches_the_reference_compute_wer(golden_dir):
    z = np.load(os.path.join(golden_dir, "edit_distance_golden.npz"))
    out = evaluation.edit_distance(torch.from_numpy(z["refs"]).cuda(), z["ref_lens"], torch.from_numpy(z["hyps"]).cuda(),
                                   z["hyp_lens"]).cpu().numpy()
    ok = z["ok"].astype(bool)
    assert np.array_equal(out[ok], z["out"][ok])                      # bit-exact where the reference returns
    for b in np.nonzero(~ok)[0]:                                      # the reference raised: the documented semantics
        ref, hyp = z["refs"][b, :z["ref_lens"][b]], z["hyps"][b, :z["hyp_lens"][b]]
        assert tuple(out[b]) == eval_ref.compute_wer(list(ref), list(hyp))


def test_edit_distance_long_sequences_and_ragged_batch():
    rng = np.random.RandomState(61)
    refs, hyps = [], []
    for (R, H, V) in [(400, 380, 30), (1, 300, 5), (300, 1, 5), (0, 17, 4), (23, 0, 4), (0, 0, 2), (257, 256, 2),
                      (600, 650, 40)]:
        refs.append(rng.randint(0, V, size=R)); hyps.append(rng.randint(0, V, size=H))
    refs_t, ref_lens = evaluation._padded_i32(refs, torch.device("cuda"))
    hyps_t, hyp_lens = evaluation._padded_i32(hyps, torch.device("cuda"))
    out = evaluation.edit_distance(refs_t, ref_lens, hyps_t, hyp_lens).cpu().numpy()
    for b in range(len(refs)):
        assert tuple(out[b]) == eval_ref.compute_wer(list(refs[b]), list(hyps[b])), b


def test_compute_wer_drop_in():
    ref = "the cat sat on the mat".split()
    hyp = "the cat sat sat on mat today".split()
    wer, sub, ins, dele = evaluation.compute_wer(ref, hyp)
    assert (wer, sub, ins, dele) == eval_ref.compute_wer(ref, hyp)
    wer_n, _, _, _ = evaluation.compute_wer(ref, hyp, normalize=True)
    assert abs(wer_n - wer / len(ref)) < 1e-12
    assert evaluation.compute_wer(list("abc"), list("abc")) == (0, 0, 0, 0)


@pytest.mark.parametrize("B,T,V,temp", [(3, 50, 30, 1.0), (2, 40, 62, 2.0), (2, 9, 3386, 0.7), (1, 1, 1, 1.0)])
def test_posteriors_softmax_with_temperature(B, T, V, temp):
    rng = np.random.RandomState(B + T + V)
    logits = (rng.randn(B, T, V) * 3).astype(np.float32)
    got = evaluation.posteriors(torch.from_numpy(logits).cuda(), temperature=temp).cpu().numpy()
    ref = eval_ref.posteriors(logits, temp)
    assert np.max(np.abs(got - ref)) < 2e-6 and np.max(np.abs(got.sum(-1) - 1)) < 1e-5
    view = torch.from_numpy(logits.transpose(1, 0, 2).copy()).cuda().transpose(0, 1)     # strided [B,T,V] view
    assert torch.equal(evaluation.posteriors(view, temperature=temp).cpu(), torch.from_numpy(got))


def test_evaluate_batch_decodes_and_scores_on_the_device():
    """greedy decode -> hypotheses - 1 (ctc.py:444) -> edit distance against the references, for a whole
    mini-batch, equals the reference's per-utterance pipeline (its GreedyDecoder restated in the oracle +
    compute_wer)."""
    rng = np.random.RandomState(62)
    B, T, V, Lmax = 7, 120, 30, 40
    logits = (rng.randn(B, T, V) * 2).astype(np.float32)
    logits[:, :, 0] += 1.5                                   # frequent blanks: hypotheses of realistic length
    x_lens = np.sort(rng.randint(T // 2, T + 1, size=B))[::-1].copy()
    y_lens = rng.randint(5, Lmax + 1, size=B)
    ys = np.zeros((B, Lmax), np.int64)
    for b in range(B):
        ys[b, :y_lens[b]] = rng.randint(0, V - 1, size=y_lens[b])
    errors, hyps, hyp_lens = evaluation.evaluate_batch(torch.from_numpy(logits).cuda(), x_lens, ys, y_lens)
    errors, hyps, hyp_lens = errors.cpu().numpy(), hyps.cpu().numpy(), hyp_lens.cpu().numpy()
    ref_hyps = ctc_ref.greedy_decode(logits, x_lens)
    for b in range(B):
        h = np.asarray(ref_hyps[b]) - 1
        assert np.array_equal(hyps[b, :hyp_lens[b]], h)
        assert tuple(errors[b]) == eval_ref.compute_wer(list(ys[b, :y_lens[b]]), list(h))
