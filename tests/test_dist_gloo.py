"""Data-parallel host logic with world_size 2 on CPU (gloo).

Slots are sharded across ranks; each rank produces the UNNORMALISED gradient arena of its slots
(here: the oracle's hand-written backward, standing in for wn_train_backward) plus its loss
statistics; the bucket plan all-reduces them; dividing by the global n_valid must reproduce the
single-process gradient of the reference's loss (tmodel.py:244-249) -- the invariant of SURVEY.md
4.5 -- and the loader shards must deal exactly the slots of the global batch.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wavenet_oracle as O
from tests import util

ARCH = util.TINY_GC
B, T = 4, 48


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(reg, grads):
    flat = torch.zeros(reg.n_param_elems, dtype=torch.float64)
    for name, info in reg.params.items():
        flat[info.offset:info.offset + info.numel] = torch.as_tensor(np.asarray(grads[name])).reshape(-1)
    return flat


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from lb_wavenet_b200.data import SlotDealer
    from lb_wavenet_b200.dist import DistContext, bucket_plan, reduce_plan_sync
    from lb_wavenet_b200.engine import Registry
    ctx = DistContext.from_env("gloo")
    assert (ctx.rank, ctx.world) == (rank, world)
    lo, hi = ctx.slot_range(B)
    a = util.oracle_arch(ARCH)
    p = util.scaled_params(a, B, 3)
    rng = np.random.default_rng(0)
    cat = [(int(rng.integers(1, 11)), rng.integers(0, 256, int(rng.integers(60, 300))).astype(np.int32))
           for _ in range(9)]
    wav, ids = SlotDealer(cat, B, T, a.recep_field(), 1, 5, 0, lo, hi, quiet=True).next_batch()[1:]
    # this rank's slots, their SAVE rows, full weight replica
    p_local = {k: (v[lo:hi] if k.startswith("SAVE") else v) for k, v in p.items()}
    pt, save, _ = O.to_torch_params(a, p_local, hi - lo, torch.float64, False)
    g, info = O.train_backward_manual(a, pt, save, torch.as_tensor(wav).long(), torch.as_tensor(ids).long())
    reg = Registry(ARCH, hi - lo)
    flat = _pack(reg, g)
    stats = torch.tensor([info["xent_sum"], float(info["n_valid"]), 0.0], dtype=torch.float64)
    ctx.all_reduce_sum_(stats)
    plan = bucket_plan([(n, i.offset, i.numel) for n, i in reg.params.items()], reg.n_layers,
                       ARCH["n_block_layers"], reg.n_param_elems, ARCH["n_gc_embed"] > 0)
    reduce_plan_sync(ctx, flat, plan)
    gathered = ctx.all_gather_cat(torch.as_tensor(wav), dim=0)
    if rank == 0:
        torch.save(dict(flat=flat, stats=stats, wav=gathered), os.path.join(out_dir, "r0.pt"))
    ctx.barrier()
    dist.destroy_process_group()


def test_two_rank_slot_sharding_reproduces_single_process_gradients(tmp_path, lib):
    from lb_wavenet_b200.data import SlotDealer
    from lb_wavenet_b200.engine import Registry
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "r0.pt"))
    a = util.oracle_arch(ARCH)
    p = util.scaled_params(a, B, 3)
    rng = np.random.default_rng(0)
    cat = [(int(rng.integers(1, 11)), rng.integers(0, 256, int(rng.integers(60, 300))).astype(np.int32))
           for _ in range(9)]
    wav, ids = SlotDealer(cat, B, T, a.recep_field(), 1, 5, 0, quiet=True).next_batch()[1:]
    assert np.array_equal(got["wav"].numpy(), wav)  # shards == slices of the global dealing
    grads, L, _ = O.train_step_autograd(a, p, wav, ids, 0.0, torch.float64)
    assert int(got["stats"][1]) == L.n_valid and abs(float(got["stats"][0]) - float(L.xent_sum)) < 1e-9
    reg = Registry(ARCH, B)
    ref = _pack(reg, grads)
    assert torch.allclose(got["flat"] / L.n_valid, ref, rtol=1e-9, atol=1e-12)
