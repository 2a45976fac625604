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


@pytest.mark.parametrize("V", [515, 1003, 4100])
def test_large_vocabulary_special_values_at_every_position(V):
    """The streaming arg-max kernel (V > 512): scalar head up to 16-byte alignment, unrolled and single 128-bit body
    loads, scalar tail; ties, NaN, infinities and signed zeros in each of those regions, rows at every alignment."""
    rng = np.random.RandomState(V)
    T = 24
    logits = np.round(rng.randn(2, T, V) * 2).astype(np.float32) / 2     # many exact ties
    pos = [v for v in (0, 1, 2, 3, 5, 130, 511, 512, 700) if v < V - 4] + [V - 4, V - 3, V - 2, V - 1]
    for t, v in enumerate(pos):
        logits[0, t, v] = 9.0                                             # unique maximum at a chosen position
        logits[1, t, v] = np.nan if t % 2 else np.inf
        logits[1, t, (v + 7) % V] = np.inf                                # inf loses to NaN, ties with inf: first index
    logits[0, 20, :] = -0.0; logits[0, 20, 3] = 0.0                      # -0.0 == +0.0: index 0
    logits[0, 21, :] = -np.inf                                            # all -inf: index 0
    logits[0, 22, [V - 1, 4]] = 50.0                                      # tie across tail and head: 4
    for blank in (0, 5):
        for off in (0, 1, 3):                                             # base pointer at different 16-byte phases
            buf = torch.empty(logits.size + 4, device="cuda")
            view = buf[off:off + logits.size].view(2, T, V)
            view.copy_(torch.from_numpy(logits))
            tokens, lens = b200.greedy_decode(view, np.array([T, T - 3]), blank)
            ref = ctc_ref.greedy_decode(logits, [T, T - 3], blank)
            for b in range(2):
                assert np.array_equal(tokens[b, :lens[b]].cpu().numpy(), ref[b]), (blank, off, b)


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
