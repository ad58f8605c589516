"""Multi-GPU plumbing for the hot path.

The path shards over independent units -- images of a batch, NoC evaluation samples -- one
process per GPU, with NO data-path collective.  The only collectives are the ones the
reference has: the all-reduce (mean) of the trainable gradients (head + click embedding,
11.5 MB; core/utils/distributed.py:66-78, core/training/trainer.py:144-149) and the final
gather of per-sample results for the NoC table (our addition: the reference evaluates on one
GPU, core/inference/utils.py:270-274).  Backend: NCCL on GPUs, gloo in the CPU tests.
"""
from typing import List, Sequence

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_indices(n: int, world_size: int = None, rank_: int = None) -> List[int]:
    """Round-robin assignment of n units (images / evaluation samples) to ranks: unit i goes to
    rank i % world.  Round-robin (not contiguous blocks) spreads early-exit samples of the NoC
    loop evenly.  Every unit is owned by exactly one rank."""
    w = world() if world_size is None else world_size
    r = rank() if rank_ is None else rank_
    return list(range(r, n, w))


def shard_batch(t: torch.Tensor, world_size: int = None, rank_: int = None) -> torch.Tensor:
    """This rank's images of a global batch (per-GPU batch = global // world, like the reference's
    `batch_size // ngpus`, core/training/trainer.py:66-68)."""
    idx = shard_indices(t.shape[0], world_size, rank_)
    return t[idx]


class FlatGradArena:
    """All trainable gradients in ONE contiguous buffer so the step needs a single all-reduce
    (latency-bound at 11.5 MB: one launch instead of DDP's per-bucket calls).  Parameters keep
    their own `.grad` views into the arena."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        """Mean over ranks, in place (DDP semantics)."""
        if world() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(world())
        return self.flat


def gather_sample_results(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Collect per-sample rows (e.g. [n_local, 20] IoU-per-click) computed on round-robin shards
    back into global sample order on every rank.  `local[j]` belongs to sample rank + j*world."""
    w, r = world(), rank()
    if w == 1:
        return local
    n_max = (n_total + w - 1) // w
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(parts, pad)
    out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for rr in range(w):
        idx = list(range(rr, n_total, w))
        out[idx] = parts[rr][: len(idx)]
    return out
