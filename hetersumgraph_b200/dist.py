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
        n_used = sum(p.numel() for p in self.params)
        n = (n_used + 3) // 4 * 4          # 16-byte multiple: the peer all-reduce (PeerAllReduceAdam) moves 128-bit words
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
                buf = torch.zeros(n, dtype=torch.float32, device=dev)
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


# fp32 bound of `gradient_parity`: the all-reduced sum of N shard gradients and the gradient of the whole batch on one
# GPU add the same per-row terms in a different association (per-shard weight-gradient products and LayerNorm / bias
# column sums over ~48 k rows each, then one NCCL sum, against one pass over ~95 k rows).  Measured on 2 x B200:
# 1.45e-6 of the largest gradient entry (gpurun r02m); SURVEY 8-e's estimate was 1e-6.  Every kernel is run-to-run
# bitwise deterministic, and both sides hold 1e-5 against the fp64 closed form (tests/test_gpu_parity.py).
GRAD_PARITY_BOUND = 3e-6


def gradient_parity(rank, world, dev, all_reduce=None, n_global=256, seed=3):
    """"1-GPU vs G-GPU gradients".  ONE seeded global batch of `n_global` CNN/DM-shaped graphs is dealt to
    the ranks by `shard_indices` (module/dataloader.py:479-480 order + snake deal); every rank runs the fused train
    step (loss scaled by 1/n_global) on its shard, the flat gradient arenas are summed by `all_reduce` (NCCL), and
    rank 0 compares the sum with its own run of the WHOLE batch.  Returns the report dict on rank 0, None elsewhere."""
    from . import _lib
    from . import synthetic as syn
    from .graph import HeteroBatch
    from .path_model import FusedTrainStep, HSGPath
    # Kernel selection is pinned for the comparison: every product goes to the tensor-core path.  A shard and the whole
    # batch otherwise fall on different sides of the size thresholds (the exact-fp32 small-product / one-launch FFN
    # kernels below 3e8 flops, 3xTF32 above; the recomputing edge prep from 65 536 rows on): each is inside the fp32
    # class, but their ~1e-7 differences flip single ReLU units at their kink - measured at 256 graphs on 2 ranks:
    # 1.3e-5 on 48 of 433 k gradient elements, 1e-7 with the selection pinned (profiles/shard_parity.py).
    lib = _lib.load()
    lib.hsg_set_gemm_small_flops(0.0)
    lib.hsg_set_edge_recompute(0)
    try:
        return _gradient_parity(rank, world, dev, all_reduce, n_global, seed)
    finally:
        lib.hsg_set_gemm_small_flops(3e8)
        lib.hsg_set_edge_recompute(-1)


def _gradient_parity(rank, world, dev, all_reduce, n_global, seed):
    from . import synthetic as syn
    from .graph import HeteroBatch
    from .path_model import FusedTrainStep, HSGPath
    exs_all = syn.make_examples(n_global, "cnndm", seed=seed)
    sf_all = torch.randn(sum(e.n_sent for e in exs_all) + 8, 64, generator=torch.Generator().manual_seed(11))
    # sentence rows follow the examples: every example owns a slice, so a shard sees the same features as the full batch
    offs = np.concatenate([[0], np.cumsum([e.n_sent for e in exs_all])])

    def run(exs_sub, idxs):
        tbs = syn.pack_token_batch(exs_sub)
        order = list(tbs.order) if getattr(tbs, "order", None) is not None else list(range(len(exs_sub)))
        rows = np.concatenate([np.arange(offs[idxs[j]], offs[idxs[j]] + exs_sub[j].n_sent) for j in order]) \
            if len(exs_sub) else np.zeros(0, np.int64)
        batch = HeteroBatch.from_token_batch(tbs, dev)
        torch.manual_seed(1234)
        m = HSGPath(n_iter=1).to(dev)
        ar = FlatGradArena(m.parameters(), flatten_params=True)
        m.loop.fuse_grad_accumulation = True
        FusedTrainStep(m, n_global)(batch, sf_all[rows].to(dev))
        return ar.flat.clone()

    sh = shard_indices([e.n_sent for e in exs_all], [float(sum(len(x) for x in e.w2s)) for e in exs_all], world)
    g_shard = run([exs_all[i] for i in sh[rank]], sh[rank])
    if all_reduce is not None:
        all_reduce(g_shard)
    if rank != 0:
        return None
    g_full = run(exs_all, list(range(n_global)))
    err = float((g_shard - g_full).abs().max() / g_full.abs().max())
    return {"global_graphs": n_global, "ranks": world, "normalised_max_error": err, "bound": GRAD_PARITY_BOUND,
            "survey_estimate": 1e-6, "ok": bool(err <= GRAD_PARITY_BOUND),
            "what": "all-reduced flat gradient arena of the N shards (dist.shard_indices) vs rank 0 running the whole "
                    "global batch alone, same parameters and sent_feature rows, kernel selection pinned (every product "
                    "on the tensor-core path: see gradient_parity); bound = fp32 reassociation of the row sums"}


class PeerAllReduceAdam:
    """Gradient all-reduce + Adam + zero_grad of a `functional.FusedAdam` in ONE kernel over NVLink peer memory
    (`hsg_allreduce_adam_step`, csrc/hsg_head.cu) instead of `dist.all_reduce` (NCCL) followed by `step_dev`: the
    433 k-float arena of the path is latency-bound, NCCL costs ~75 us per step for it at any rank count, the fused
    kernel a fraction of that.  The receive buffers come from torch's symmetric-memory allocator (peer mappings of one
    box); `available()` tells whether that works here, callers fall back to NCCL + `step_dev` otherwise.

        fused = PeerAllReduceAdam(opt)          # collective: every rank of the default group
        ...backward...; fused.step()            # instead of all_reduce(opt.g); opt.step_dev(zero_grad=True)
    """

    def __init__(self, opt, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        if opt.max_grad_norm > 0.0:
            raise ValueError("PeerAllReduceAdam: gradient clipping needs the global norm after the reduce - use NCCL + step_dev")
        self.opt = opt
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        lib = _lib.load()
        n = opt.g.numel()
        if n % 4 != 0:
            raise ValueError("PeerAllReduceAdam: arena of %d floats is not a 16-byte multiple (use dist.FlatGradArena)" % n)
        dev = opt.g.device
        self.buf = symm_mem.empty(int(lib.hsg_allreduce_adam_buffer_floats(n, self.world)), dtype=torch.float32, device=dev)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.peers = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=dev)
        self.state = opt.device_step_counter()              # [0] completed steps (shared with step_dev), [1..3] tickets
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)                            # every rank's buffer is zeroed before anybody pushes

    @staticmethod
    def available():
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory  # noqa: F401
            return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        except Exception:
            return False

    def step(self):
        from . import _lib
        from .functional import _p, _st
        o = self.opt
        _lib.check(_lib.load().hsg_allreduce_adam_step(o.p.numel(), _p(o.p), _p(o.g), _p(o.m), _p(o.v), o.lr, o.betas[0],
                                                       o.betas[1], o.eps, self.state.data_ptr(), self.peers.data_ptr(),
                                                       self.rank, self.world, _st()))

