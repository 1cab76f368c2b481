"""Data parallelism over the B independent slots (new work: the reference is single-device).

One process per GPU.  Rank r owns slots [r*B/N, (r+1)*B/N) together with their SAVE rows and
loader state; every rank holds a full weight replica.  One exchange per optimiser step:
all-reduce(sum) of (xent_sum, n_valid, diff_sum) right after the loss kernel and of the
UNNORMALISED gradient arena in a few contiguous buckets, each issued on a side stream as soon as
the backward phases that produce it have been queued (wn_train_backward_phases), so NCCL overlaps
the rest of backward.  The 1/n_valid_global scale and the L2 term are applied afterwards by
wn_adam_step, identically on every rank.

The helpers here are device-agnostic so that the bucket plan and reduction semantics are tested
with gloo on CPU tensors (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple


@dataclass
class DistContext:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    group: object = None

    @staticmethod
    def from_env(backend: Optional[str] = None) -> "DistContext":
        """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* when launched
        by torchrun; a plain single process otherwise."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1:
            return DistContext()
        import torch
        import torch.distributed as dist
        rank = int(os.environ["RANK"])
        local_rank = int(os.environ.get("LOCAL_RANK", rank))
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
        return DistContext(rank, world, local_rank, None)

    def slot_range(self, batch_sz: int) -> Tuple[int, int]:
        if batch_sz % self.world != 0:
            raise ValueError("batch_sz {} must be divisible by the number of ranks {}".format(batch_sz, self.world))
        per = batch_sz // self.world
        return self.rank * per, (self.rank + 1) * per

    def all_reduce_sum_(self, t) -> None:
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def barrier(self) -> None:
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)

    def all_gather_cat(self, t, dim: int = 0):
        """Gather equally-shaped shards from every rank and concatenate along dim (SAVE rows on save)."""
        if self.world == 1:
            return t
        import torch
        import torch.distributed as dist
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t.contiguous(), group=self.group)
        return torch.cat(parts, dim=dim)


def bucket_plan(param_offsets: Sequence[Tuple[str, int, int]], n_layers: int, n_block_layers: int,
                total_elems: int, has_gc: bool, n_buckets: int = 3) -> List[Tuple[int, List[Tuple[int, int]]]]:
    """Plan the overlap of gradient all-reduce with backward.

    param_offsets: (serial name, offset, numel) in arena order.  Returns a list of
    (phase_end, [(lo, hi) arena ranges that are final once phases < phase_end are queued]).
    Arena order is [GC_EMBED, PRE, PRE_BIAS | layer 0 .. layer L-1 | POST1.., POST2..]; phases as in
    wn_train_backward_phases (0 = post-net, p = layer L-p, L+1 = PRE/GC).  With global
    conditioning the per-layer GC projections are only written by the last phase, so the layer
    region is reduced at the end in one bucket.
    """
    L = n_layers

    def layer_of(name: str) -> Optional[int]:
        parts = name.split("_")
        if len(parts) >= 3 and parts[-1].isdigit() and parts[-2].isdigit():
            return int(parts[-2]) * n_block_layers + int(parts[-1])
        return None

    layer_lo = [None] * L
    post_lo = None
    for name, off, _ in param_offsets:
        l = layer_of(name)
        if l is not None:
            if layer_lo[l] is None:
                layer_lo[l] = off
        elif name.startswith("POST") and post_lo is None:
            post_lo = off
    assert post_lo is not None and all(x is not None for x in layer_lo)
    plan: List[Tuple[int, List[Tuple[int, int]]]] = []
    plan.append((1, [(post_lo, total_elems)]))  # after phase 0
    if has_gc or L < 2 * n_buckets:
        plan.append((L + 2, [(0, post_lo)]))
        return plan
    hi_layer = L
    for b in range(n_buckets):
        lo_layer = (L * (n_buckets - 1 - b)) // n_buckets
        hi_off = post_lo if hi_layer == L else layer_lo[hi_layer]
        plan.append((L - lo_layer + 1, [(layer_lo[lo_layer], hi_off)]))
        hi_layer = lo_layer
    plan.append((L + 2, [(0, layer_lo[0])]))
    return plan


def reduce_plan_sync(ctx: DistContext, grads, plan) -> None:
    """Reference (non-overlapped) execution of a bucket plan: same ranges, same order."""
    for _, ranges in plan:
        for lo, hi in ranges:
            if hi > lo:
                ctx.all_reduce_sum_(grads[lo:hi])
