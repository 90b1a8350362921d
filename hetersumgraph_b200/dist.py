"""Data parallelism by graph.

Document graphs are disconnected components of the batched graph (dgl.batch, module/dataloader.py:480) and
the loss is a mean over graphs (train.py:118-119), so a global batch shards across ranks with NO data-path
collective: every rank builds and processes its own HeteroBatch.  The only exchange is one all-reduce (sum)
of the contiguous fp32 gradient arena per step; each rank scales its loss by 1/B_global so the summed
gradient equals the single-process gradient.
"""
from typing import List, Sequence

import numpy as np
import torch


def shard_indices(n_sent: Sequence[int], weights: Sequence[float], world_size: int) -> List[List[int]]:
    """Deal graphs to ranks: global stable sort by #sentences descending (module/dataloader.py:479), then a
    snake (boustrophedon) deal so every rank gets a similar weight (edge count) and its own shard stays sorted
    descending (what pack_padded_sequence in HiGraph.py:137 needs).  Deterministic; every graph lands on exactly
    one rank."""
    order = np.argsort(-np.asarray(n_sent, np.int64), kind="stable")
    shards: List[List[int]] = [[] for _ in range(world_size)]
    load = np.zeros(world_size)
    for pos, idx in enumerate(order.tolist()):
        rnd, k = divmod(pos, world_size)
        r = k if rnd % 2 == 0 else world_size - 1 - k
        shards[r].append(idx)
        load[r] += weights[idx]
    return shards


class FlatGradArena:
    """All trainable gradients of a module as views into ONE contiguous fp32 buffer, so the data-parallel
    exchange is a single in-place all-reduce (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests)."""

    def __init__(self, params, flatten_params=False):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.flat_param = None
        if flatten_params:
            # parameters become views of one buffer as well: the optimizer then updates ONE tensor with ONE fused
            # kernel (Adam is element-wise, so this is the same arithmetic as per-tensor Adam)
            with torch.no_grad():
                buf = torch.empty(n, dtype=torch.float32, device=dev)
                off = 0
                for p in self.params:
                    buf[off:off + p.numel()].copy_(p.reshape(-1))
                    p.data = buf[off:off + p.numel()].view_as(p)
                    off += p.numel()
            self.flat_param = torch.nn.Parameter(buf)
            self.flat_param.grad = self.flat

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat
