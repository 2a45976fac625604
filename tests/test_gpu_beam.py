"""GPU tests of the CTC prefix beam search: bit-exact hypotheses against golden vectors
produced by the reference's own BeamSearchDecoder, and against the oracle on further seeded inputs."""
import os

import numpy as np
import pytest
import torch

import pytorch_end2end_speech_recognition_b200 as b200
from oracle import beam_ref, ctc_ref

pytestmark = pytest.mark.gpu


def test_beam_search_matches_the_reference_decoder(golden_dir):
    z = np.load(os.path.join(golden_dir, "beam_golden.npz"))
    for i in range(int(z["n_cases"])):
        lp, x_lens, W = z["log_probs_%d" % i], z["x_lens_%d" % i], int(z["beam_%d" % i])
        tokens, lens = b200.beam_search_decode(torch.from_numpy(lp).cuda(), x_lens, W)
        tokens, lens = tokens.cpu().numpy(), lens.cpu().numpy()
        assert np.array_equal(lens, z["hyp_lens_%d" % i]), i
        offs = np.concatenate([[0], np.cumsum(lens)])
        for b in range(lp.shape[0]):
            assert np.array_equal(tokens[b, :lens[b]], z["hyp_flat_%d" % i][offs[b]:offs[b + 1]]), (i, b)
            assert np.all(tokens[b, lens[b]:] == -1)


@pytest.mark.parametrize("B,T,V,W,blank", [(4, 45, 30, 10, 0), (3, 30, 8, 4, 3), (2, 60, 30, 32, 0), (2, 12, 200, 5, 0),
                                           (3, 20, 4, 64, 0)])
def test_beam_search_against_the_oracle(B, T, V, W, blank):
    rng = np.random.RandomState(B * 100 + T + V + W)
    logits = (rng.randn(B, T, V) * 2.0).astype(np.float32)
    logits[:, :, blank] += 1.0
    lp = torch.log_softmax(torch.from_numpy(logits), dim=-1)
    x_lens = rng.randint(T // 2, T + 1, size=B)
    tokens, lens, scores = b200.beam_search_decode(lp.cuda(), x_lens, W, blank=blank, return_scores=True)
    tokens, lens, scores = tokens.cpu().numpy(), lens.cpu().numpy(), scores.cpu().numpy()
    for b in range(B):
        hyp, score = beam_ref.beam_search(lp[b].numpy(), x_lens[b], W, blank=blank)
        assert list(tokens[b, :lens[b]]) == hyp, b
        assert abs(scores[b] - score) < 1e-5 * max(1.0, abs(score))


def test_beam_width_one_and_the_drop_in_class():
    rng = np.random.RandomState(5)
    logits = (rng.randn(3, 40, 30) * 3).astype(np.float32)
    lp = torch.log_softmax(torch.from_numpy(logits), dim=-1).numpy()
    x_lens = np.array([40, 31, 0])
    dec = b200.BeamSearchDecoder(blank_index=0)
    hyps = dec(lp, x_lens, beam_width=1)
    for b in range(3):
        hyp, _ = beam_ref.beam_search(lp[b], x_lens[b], 1)
        assert list(hyps[b]) == hyp
    assert len(hyps[2]) == 0
    # a strided (time-major) view is consumed without a copy
    tm = torch.from_numpy(lp.transpose(1, 0, 2).copy()).cuda().transpose(0, 1)
    t2, l2 = b200.beam_search_decode(tm, x_lens, 7)
    t1, l1 = b200.beam_search_decode(torch.from_numpy(lp).cuda(), x_lens, 7)
    assert torch.equal(t1, t2) and torch.equal(l1, l2)


def test_evaluate_batch_with_beam_search():
    from oracle import eval_ref
    from pytorch_end2end_speech_recognition_b200 import evaluation
    rng = np.random.RandomState(9)
    B, T, V = 4, 50, 12
    logits = (rng.randn(B, T, V) * 2).astype(np.float32)
    logits[:, :, 0] += 1.0
    x_lens = np.array([50, 44, 40, 33])
    y_lens = np.array([12, 9, 10, 7])
    ys = np.zeros((B, 12), np.int64)
    for b in range(B):
        ys[b, :y_lens[b]] = rng.randint(0, V - 1, size=y_lens[b])
    errors, hyps, hyp_lens = evaluation.evaluate_batch(torch.from_numpy(logits).cuda(), x_lens, ys, y_lens, beam_width=5)
    lp = torch.log_softmax(torch.from_numpy(logits).cuda(), dim=-1).cpu().numpy()
    errors, hyps, hyp_lens = errors.cpu().numpy(), hyps.cpu().numpy(), hyp_lens.cpu().numpy()
    for b in range(B):
        hyp, _ = beam_ref.beam_search(lp[b], x_lens[b], 5)
        h = np.asarray(hyp, dtype=np.int64) - 1
        assert np.array_equal(hyps[b, :hyp_lens[b]], h)
        assert tuple(errors[b]) == eval_ref.compute_wer(list(ys[b, :y_lens[b]]), list(h))
