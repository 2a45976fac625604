"""world_size-2 gloo test (CPU) of the multi-GPU host logic: every rank takes its
length-balanced shard, computes per-utterance costs (oracle as a stand-in for the kernel, which
needs a GPU), and the scalar loss is all-reduced.  The sum must equal the single-process loss."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ctc_ref
    from pytorch_end2end_speech_recognition_b200 import shard, workloads
    wl = workloads.make_lengths_and_labels(None, B=7, T=40, V=9, Lmax=10, kind="var", seed=11)
    acts = workloads.make_acts(wl).numpy()
    index = shard.balance_shards(wl.act_lens, wl.label_lens, wl.V, world)[rank]
    flat, ll, al = shard.shard_batch(wl.labels, wl.label_lens, wl.act_lens, index)
    costs, _ = ctc_ref.ctc_cost_and_grad(acts[:, index], flat, al, ll)
    loss = torch.tensor([costs.sum()], dtype=torch.float64)
    loss = shard.allreduce_loss(loss)
    full, _ = ctc_ref.ctc_cost_and_grad(acts, wl.labels, wl.act_lens, wl.label_lens)
    q.put((rank, float(loss[0]), float(full.sum()), index.tolist()))
    dist.destroy_process_group()


def test_sharded_loss_allreduce_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = sorted(sum((r[3] for r in res), []))
    assert seen == list(range(7))
    for _, loss, full, _ in res:
        assert abs(loss - full) < 1e-9 * abs(full)
