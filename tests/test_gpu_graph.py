"""CUDA-graph replay of the whole training step (hetersumgraph_b200.step_graph): a captured + replayed step must be
bit-equal to the same step enqueued eagerly - loss, logits, d_sent_feature, every parameter and both Adam moments - over
several steps, with batches of different shapes alternating (graph reuse, key misses, re-capture), for HSG and HDSG,
with and without training-mode dropout.  Also the device-step Adam and the embedding gather against their stock
counterparts (train.py:90,131-135; HiGraph.py:147-148)."""
import numpy as np
import pytest
import torch

import hetersumgraph_b200 as hb
from hetersumgraph_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _make(hdsg, p_drop, seed=9, n_iter=1):
    from hetersumgraph_b200.dist import FlatGradArena
    from hetersumgraph_b200.functional import FusedAdam
    from hetersumgraph_b200.path_model import HSGPath
    torch.manual_seed(seed)
    model = HSGPath(n_iter=n_iter, hdsg=hdsg, atten_dropout_prob=p_drop, ffn_dropout_prob=p_drop).cuda()
    model.train()
    arena = FlatGradArena(model.parameters(), flatten_params=True)
    model.loop.fuse_grad_accumulation = True
    opt = FusedAdam(arena.flat_param.data, arena.flat, lr=5e-4)
    return model, arena, opt


def _hosts(hdsg, shapes):
    from hetersumgraph_b200.graph import DeviceTokenBatch
    out = []
    for n, seed in shapes:
        exs = syn.make_examples(n, "multinews" if hdsg else "tiny", seed=seed, hdsg=hdsg)
        tb = syn.pack_token_batch(exs, hdsg=hdsg)
        host, _ = DeviceTokenBatch.host_buffers(tb)
        gen = torch.Generator().manual_seed(100 + seed)
        sf = torch.randn(int(tb.tokens.shape[0]), 64, generator=gen)
        out.append((host, tb, sf))
    return out


@pytest.mark.parametrize("hdsg,p_drop,sf_on_host", [(False, 0.0, True), (False, 0.0, False), (True, 0.0, True),
                                                    (False, 0.1, True), (False, 0.0, "prefetch")])
def test_replayed_step_is_bit_equal_to_eager(hdsg, p_drop, sf_on_host):
    """sf_on_host == "prefetch": the host sent_feature of step i+1 is handed to step i (next_sent_feature) and uploaded
    by its side branch; the eager run of that case uses the plain host path, so the comparison also shows that the
    prefetched rows are the right ones."""
    from hetersumgraph_b200.step_graph import GraphedTrainStep
    hosts = _hosts(hdsg, [(6, 51), (6, 52), (6, 53)])
    tb0 = hosts[0][1]
    bitmap = torch.from_numpy(tb0.filter_bitmap.view(np.int32).copy()).cuda()
    seq = [0, 1, 0, 1, 0, 1, 2, 0, 1, 0, 1]                 # A B A B A B C A B A B
    runs = []
    base_seed = None
    for capture in (False, True):
        model, arena, opt = _make(hdsg, p_drop)
        if base_seed is not None:
            model.loop.__dict__["_base_seed"] = base_seed    # same host-side dropout seed in both runs
        gs = GraphedTrainStep(model, opt, bitmap, n_graphs_global=6, capture=capture)
        gs.prime(hosts[seq[0]][0])
        rec = []
        for i, k in enumerate(seq):
            nxt = hosts[seq[i + 1]][0] if i + 1 < len(seq) else None
            sf = hosts[k][2] if sf_on_host else hosts[k][2].cuda()
            nsf = hosts[seq[i + 1]][2] if (sf_on_host == "prefetch" and capture and i + 1 < len(seq)) else None
            loss_h, logits, d_sf = gs.step(nxt, sf, next_sent_feature=nsf)
            loss = gs.sync_loss()
            rec.append((loss, logits.clone(), d_sf.clone(), arena.flat_param.data.clone(), opt.m.clone(), opt.v.clone()))
        base_seed = model.loop.__dict__.get("_base_seed")
        runs.append((rec, gs))
    (eager, gs_e), (graph, gs_g) = runs
    assert gs_e.replays == 0 and gs_g.replays >= 5, (gs_e.replays, gs_g.replays)
    assert float(opt.device_step_counter()[0]) == len(seq)
    for i, (a, b) in enumerate(zip(eager, graph)):
        assert a[0] == b[0], "loss differs at step %d: %r vs %r" % (i, a[0], b[0])
        for x, y in zip(a[1:], b[1:]):
            assert torch.equal(x, y), "step %d" % i
    assert np.isfinite(eager[-1][0]) and float(eager[-1][3].abs().max()) > 0
    # the arena is left zeroed by the optimizer (zero_grad folded into Adam)
    assert float(arena.flat.abs().max()) == 0.0
    if p_drop > 0:            # masks change from step to step: same batch, same parameters-ish, different loss path
        assert eager[0][0] != eager[2][0]


def test_graphed_step_equals_plain_fused_step():
    """GraphedTrainStep (static slots, capacity-sized builder outputs, device-step Adam) against the plain path
    (HeteroBatch.from_token_batch + FusedTrainStep + FusedAdam.step): same loss / logits / d_sent_feature bitwise on the
    first step, parameters equal to fp32 rounding of the bias corrections afterwards."""
    from hetersumgraph_b200.path_model import FusedTrainStep
    from hetersumgraph_b200.step_graph import GraphedTrainStep
    hosts = _hosts(False, [(8, 61), (8, 62)])
    bitmap = torch.from_numpy(hosts[0][1].filter_bitmap.view(np.int32).copy()).cuda()
    m1, a1, o1 = _make(False, 0.0)
    m2, a2, o2 = _make(False, 0.0)
    gs = GraphedTrainStep(m1, o1, bitmap, n_graphs_global=8, capture=True)
    gs.prime(hosts[0][0])
    plain = FusedTrainStep(m2, 8)
    for i in range(6):
        host, tb, sf = hosts[i % 2]
        nxt = hosts[(i + 1) % 2][0]
        _, logits1, dsf1 = gs.step(nxt, sf)
        loss1 = gs.sync_loss()
        batch = hb.HeteroBatch.from_token_batch(tb)
        a2.flat.zero_()
        loss2, logits2, dsf2 = plain(batch, sf.cuda())
        o2.step()
        if i == 0:
            assert loss1 == float(loss2) and torch.equal(logits1, logits2) and torch.equal(dsf1, dsf2)
        else:
            assert abs(loss1 - float(loss2)) <= 1e-5 * abs(float(loss2))
        err = float((a1.flat_param.data - a2.flat_param.data).abs().max())
        assert err <= 2e-6, (i, err)


def test_adam_device_step_matches_torch():
    from hetersumgraph_b200.functional import FusedAdam
    torch.manual_seed(0)
    n = 100003
    p0 = torch.randn(n)
    ref = torch.nn.Parameter(p0.clone().cuda())
    opt = torch.optim.Adam([ref], lr=5e-4)
    mine = p0.clone().cuda()
    g = torch.zeros(n, device="cuda")
    fa = FusedAdam(mine, g, lr=5e-4)
    for step in range(6):
        grad = torch.randn(n, device="cuda") * (0.1 + step)
        ref.grad = grad.clone()
        opt.step()
        g.copy_(grad)
        fa.step_dev(zero_grad=True)
        assert float(g.abs().max()) == 0.0
        assert float((mine - ref).abs().max() / ref.abs().max()) <= 1e-6
    st = fa.device_step_counter().tolist()
    assert st[:2] == [6, 0]      # [0] completed steps, [1] ticket back at 0; [2..3] belong to the peer-reduce kernel


def test_embed_gather_bit_exact():
    from hetersumgraph_b200.functional import embed_gather
    torch.manual_seed(1)
    table = torch.randn(5000, 300, device="cuda")
    ids = torch.randint(0, 5000, (12345,), device="cuda", dtype=torch.int32)
    assert torch.equal(embed_gather(ids, table), table[ids.long()])
    assert embed_gather(ids[:0], table).shape == (0, 300)
