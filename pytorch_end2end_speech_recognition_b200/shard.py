"""Length-balanced utterance sharding across GPUs + the scalar loss all-reduce.

Replaces the reference's unused multi-process stubs (``utils/parallel.py:14-33`` -- an
``mp.Pool.map`` helper -- and ``utils/dataset/base.py:260-264`` ``split_per_device`` =
``np.array_split``).  Utterances are independent in CTC (cost_b and grad[:, b, :] depend on
column b only), so the batch is partitioned by utterance with no data-path collective; the one
exchange step is the scalar loss, summed with an NCCL all-reduce over NVLink.
"""

import numpy as np


def lattice_work(act_lens, label_lens, V):
    """Per-utterance cost estimate = the algorithmic byte model of SURVEY 8(d):
    8*T_b*V (softmax rows read + gradient rows written) + 8*T_b*(2L_b+1) (lattice gathers)."""
    T_b = np.asarray(act_lens, dtype=np.int64)
    L_b = np.asarray(label_lens, dtype=np.int64)
    return 8 * T_b * int(V) + 8 * T_b * (2 * L_b + 1)


def balance_shards(act_lens, label_lens, V, n_shards):
    """Longest-processing-time-first assignment of utterances to ``n_shards`` GPUs.

    Returns a list of ``n_shards`` int64 index arrays; each shard keeps its utterances sorted by
    decreasing input length (the order the reference's encoder produces, encoders/rnn.py:318-321).
    Deterministic: ties are broken by utterance index, then by shard index.
    """
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    act_lens = np.asarray(act_lens, dtype=np.int64)
    work = lattice_work(act_lens, label_lens, V)
    order = np.lexsort((np.arange(len(work)), -work))   # by work descending, index ascending
    loads = np.zeros(n_shards, dtype=np.int64)
    shards = [[] for _ in range(n_shards)]
    for b in order:
        g = int(np.argmin(loads))                        # first least-loaded shard
        shards[g].append(int(b))
        loads[g] += work[b]
    out = []
    for g in range(n_shards):
        idx = np.asarray(shards[g], dtype=np.int64)
        if len(idx):
            idx = idx[np.lexsort((idx, -act_lens[idx]))]
        out.append(idx)
    return out


def shard_batch(labels, label_lens, act_lens, index):
    """Host-side slice of the flat label vector and the length vectors for one shard."""
    label_lens = np.asarray(label_lens, dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(label_lens)])
    labels = np.asarray(labels)
    parts = [labels[offs[b]:offs[b + 1]] for b in index]
    flat = np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
    return flat, label_lens[index].astype(np.int32), np.asarray(act_lens)[index].astype(np.int32)


def allreduce_loss(loss, group=None):
    """Sum the per-rank scalar loss over all ranks (NCCL over NVLink for CUDA tensors, gloo on CPU).
    No-op when torch.distributed is not initialised (single GPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    return loss


def sharded_ctc_loss(acts, labels, act_lens, label_lens, blank=0, group=None, rank=None, world_size=None):
    """Data-parallel CTC over the utterances of ONE global batch that every rank holds:
    rank r evaluates its length-balanced shard, the scalar losses are all-reduced.

    Returns (global_loss[1] device tensor, local_index, local_costs, local_grads) where
    local_grads is [T_local, B_local, V] for the utterances ``local_index`` (T_local = the longest
    utterance of the shard; gradients never cross GPUs)."""
    import torch
    import torch.distributed as dist
    from .ctc import ctc_loss_and_grad
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    V = acts.size(2)
    index = balance_shards(act_lens, label_lens, V, world_size)[rank]
    flat, ll, al = shard_batch(labels, label_lens, act_lens, index)
    sel = torch.as_tensor(index, device=acts.device)
    t_local = int(al.max()) if len(al) else 0
    local_acts = acts[:t_local].index_select(1, sel)      # the shard's padded length, not the batch's
    costs, loss, grads = ctc_loss_and_grad(local_acts, flat, al, ll, blank=blank)
    loss = allreduce_loss(loss, group)
    return loss, index, costs, grads
