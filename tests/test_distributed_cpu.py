"""CPU tests (gloo, world_size 2) of the N>1 host logic: sharding by graph, per-rank loss scaling 1/B_global,
one all-reduce of the flat gradient arena.  The model math on CPU is the oracle (the product path has no CPU
fallback); what is under test is the data-parallel plumbing of hetersumgraph_b200.dist."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hetersumgraph_b200 import synthetic as syn  # noqa: E402
from hetersumgraph_b200.dist import FlatGradArena, shard_indices  # noqa: E402
from oracle import closed_form as cf  # noqa: E402
from oracle import graph_builder_ref as gb  # noqa: E402


def test_shard_indices_partition_and_order():
    rng = np.random.default_rng(0)
    n_sent = rng.integers(3, 51, size=37).tolist()
    w = rng.random(37).tolist()
    for world in (1, 2, 4, 8):
        shards = shard_indices(n_sent, w, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(37))                                   # a partition
        for s in shards:
            lens = [n_sent[i] for i in s]
            assert lens == sorted(lens, reverse=True)                    # each shard sorted descending
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    assert shard_indices(n_sent, w, 2) == shard_indices(n_sent, w, 2)    # deterministic


def _params(seed):
    gen = torch.Generator().manual_seed(seed)
    p = {"_TFembed.weight": torch.randn(10, 10, generator=gen)}
    for pre, (i, o, H, bias) in {"word2sent.": (48, 16, 4, False), "sent2word.": (16, 48, 6, True)}.items():
        d = o // H
        for k in range(H):
            p[pre + "layer.heads.%d.fc.weight" % k] = torch.randn(d, i, generator=gen) * 0.2
            p[pre + "layer.heads.%d.feat_fc.weight" % k] = torch.randn(d, 10, generator=gen) * 0.2
            if bias:
                p[pre + "layer.heads.%d.feat_fc.bias" % k] = torch.randn(d, generator=gen) * 0.2
            p[pre + "layer.heads.%d.attn_fc.weight" % k] = torch.randn(1, 3 * d, generator=gen) * 0.5
        p[pre + "ffn.w_1.weight"] = torch.randn(32, o, 1, generator=gen) * 0.2
        p[pre + "ffn.w_1.bias"] = torch.randn(32, generator=gen) * 0.1
        p[pre + "ffn.w_2.weight"] = torch.randn(o, 32, 1, generator=gen) * 0.2
        p[pre + "ffn.w_2.bias"] = torch.randn(o, generator=gen) * 0.1
        p[pre + "ffn.layer_norm.weight"] = torch.rand(o, generator=gen) + 0.5
        p[pre + "ffn.layer_norm.bias"] = torch.randn(o, generator=gen) * 0.1
    p["wh.weight"] = torch.randn(2, 16, generator=gen) * 0.3
    p["wh.bias"] = torch.zeros(2)
    return p


def _loss_on(examples, idxs, params, embed, n_graphs_global):
    """sum over the shard's graphs of the per-graph CE sum, divided by the GLOBAL batch size (train.py:118-119)."""
    filt = set(syn.filter_ids().tolist())
    exs = [examples[i] for i in idxs]
    graphs = [gb.create_graph_hsg(e.sents.tolist(), e.w2s, filt) for e in exs]
    bg, order = gb.collate(graphs)
    csc = gb.derive_csc(bg)
    wfeat = embed[torch.from_numpy(bg.wid[csc["wnode_id"]])]
    gen = torch.Generator().manual_seed(1000 + sum(idxs))
    labels = torch.cat([torch.from_numpy(exs[i].labels) for i in order])
    sfeat = torch.cat([torch.randn(exs[i].n_sent, 16, generator=torch.Generator().manual_seed(77 + idxs[i]))
                       for i in order])
    _, ss = cf.update_loop_cf(csc, wfeat, sfeat, params, 1)
    logits = ss @ params["wh.weight"].t() + params["wh.bias"]
    return torch.nn.functional.cross_entropy(logits, labels, reduction="sum") / n_graphs_global


def _worker(rank, world, port, n_graphs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    examples = syn.make_examples(n_graphs, "tiny", seed=5)
    params = {k: v.clone().requires_grad_(True) for k, v in _params(3).items()}
    embed = torch.randn(50000, 48, generator=torch.Generator().manual_seed(9))
    names = sorted(params)
    arena = FlatGradArena([params[k] for k in names])
    shards = shard_indices([e.n_sent for e in examples], [e.n_sent for e in examples], world)
    arena.zero()
    loss = _loss_on(examples, shards[rank], params, embed, n_graphs)
    loss.backward()
    arena.all_reduce()
    torch.save({"flat": arena.flat.clone(), "loss": float(loss)}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_equal_single_process(tmp_path):
    n_graphs, world = 6, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_graphs, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert torch.equal(r0["flat"], r1["flat"])                            # identical after the all-reduce
    # single process, whole batch
    examples = syn.make_examples(n_graphs, "tiny", seed=5)
    params = {k: v.clone().requires_grad_(True) for k, v in _params(3).items()}
    embed = torch.randn(50000, 48, generator=torch.Generator().manual_seed(9))
    arena = FlatGradArena([params[k] for k in sorted(params)])
    arena.zero()
    full = shard_indices([e.n_sent for e in examples], [1.0] * n_graphs, 1)[0]
    loss = _loss_on(examples, full, params, embed, n_graphs)
    loss.backward()
    err = float((arena.flat - r0["flat"]).abs().max() / arena.flat.abs().max())
    assert err <= 1e-6, err                                               # fp32 reduction-order noise only
    assert abs(float(loss) - (r0["loss"] + r1["loss"])) <= 1e-5 * abs(float(loss))
