"""GPU greedy decoder: bit-exact against the reference's own GreedyDecoder outputs (golden
fixtures) and against the oracle restatement on random inputs."""
import os

import numpy as np
import pytest
import torch

import pytorch_end2end_speech_recognition_b200 as b200
from oracle import ctc_ref

pytestmark = pytest.mark.gpu


def decode(logits, x_lens, blank=0):
    tokens, lens = b200.greedy_decode(torch.from_numpy(logits).cuda(), x_lens, blank)
    tokens, lens = tokens.cpu().numpy(), lens.cpu().numpy()
    assert all(np.all(tokens[b, lens[b]:] == -1) for b in range(len(lens)))
    return [tokens[b, :lens[b]].astype(np.int64) for b in range(len(lens))]


def test_reference_golden_vectors(golden_dir):
    z = np.load(os.path.join(golden_dir, "greedy_golden.npz"))
    for i in range(int(z["n_cases"])):
        hyps = decode(z["logits_%d" % i], z["x_lens_%d" % i])
        assert [len(h) for h in hyps] == z["hyp_lens_%d" % i].tolist(), i
        flat = np.concatenate(hyps) if hyps else np.zeros(0, np.int64)
        assert np.array_equal(flat, z["hyp_flat_%d" % i]), i


@pytest.mark.parametrize("B,T,V", [(3, 50, 5), (8, 300, 30), (2, 2500, 33), (4, 100, 3386), (1, 1, 2), (2, 40, 513)])
def test_random_vs_oracle(B, T, V):
    rng = np.random.RandomState(B * 1000 + T + V)
    logits = rng.randn(B, T, V).astype(np.float32)
    logits[:, ::7, 0] += 6.0                      # plenty of blanks
    logits[:, 1::4] = logits[:, 0::4][:, :logits[:, 1::4].shape[1]]   # repeated frames
    q = np.round(logits * 2) / 2                  # quantise -> many exact ties
    x_lens = rng.randint(0, T + 1, size=B); x_lens[0] = T
    for arr in (logits, q.astype(np.float32)):
        got = decode(arr, x_lens)
        ref = ctc_ref.greedy_decode(arr, x_lens)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r)


def test_special_values_and_nonzero_blank():
    logits = np.zeros((1, 6, 4), np.float32)
    logits[0, 0, 2] = np.inf
    logits[0, 1, :] = -np.inf                     # all -inf: argmax 0
    logits[0, 2, 1] = np.nan                      # NaN wins (numpy argmax semantics)
    logits[0, 3, [1, 3]] = 5.0                    # tie -> first index
    logits[0, 4, 0] = -0.0; logits[0, 4, 1] = 0.0  # -0.0 == +0.0 -> index 0
    logits[0, 5, 3] = 1.0
    for blank in (0, 3):
        got = decode(logits, [6], blank)
        ref = ctc_ref.greedy_decode(logits, [6], blank)
        assert np.array_equal(got[0], ref[0])


def test_numpy_interface_like_the_reference():
    rng = np.random.RandomState(2)
    logits = rng.randn(3, 20, 6).astype(np.float32)
    x_lens = np.array([20, 15, 9])
    out = b200.GreedyDecoder(blank_index=0)(logits, x_lens)
    ref = ctc_ref.greedy_decode(logits, x_lens)
    assert len(out) == 3
    for g, r in zip(out, ref):
        assert np.array_equal(np.asarray(g), r)
    # strided batch-major view of a time-major tensor
    tm = torch.from_numpy(logits).cuda().transpose(0, 1).contiguous()        # [T,B,V]
    tokens, lens = b200.greedy_decode(tm.transpose(0, 1), x_lens)
    for b in range(3):
        assert np.array_equal(tokens[b, :lens[b]].cpu().numpy(), ref[b])
