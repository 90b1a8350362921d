"""CPU tests: pin the oracle (oracle/) against the golden vectors generated from the
UNMODIFIED reference (tests/golden/make_golden.py), and, when /root/reference is
present (build container), against the live reference modules on the DGL-0.4 shim."""
import os
import sys

import numpy as np
import pytest
import torch

from hetersumgraph_b200 import synthetic as syn
from oracle import closed_form as cf
from oracle import fixtures as fx
from oracle import graph_builder_ref as gb
from oracle import wswgat_ref as wr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["wswgat_hsg_default.npz", "wswgat_hdsg_small.npz", "wswgat_hsg_small.npz"]
TOL = 1e-5   # BASELINE.json: fp32 normalised max error


def nerr(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def load_fixture(name):
    z = dict(np.load(os.path.join(GOLD, name)))
    g = fx.graph_from_arrays(z, "g_")
    params = {k[2:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("p:")}
    grads = {k[3:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("gp:")}
    return z, g, params, grads


def build_graphs(z, prefix, hdsg):
    exs = fx.examples_from_arrays(z, prefix)
    filt = set(syn.filter_ids().tolist())
    graphs = []
    for e in exs:
        if hdsg:
            graphs.append(gb.create_graph_hdsg(e.doc_len, e.sents.tolist(), e.doc_tokens, e.w2s, e.w2d, filt))
        else:
            graphs.append(gb.create_graph_hsg(e.sents.tolist(), e.w2s, filt))
    return exs, graphs


@pytest.mark.parametrize("prefix,hdsg", [("hsg", False), ("hdsg", True)])
def test_builder_restatement_matches_reference_golden(prefix, hdsg):
    z = dict(np.load(os.path.join(GOLD, "builder.npz")))
    exs, graphs = build_graphs(z, prefix, hdsg)
    bg, order = gb.collate(graphs)
    assert order == z[prefix + "_order"].tolist()
    ref = fx.graph_from_arrays(z, prefix + "_g_")
    for k in fx.GRAPH_KEYS:
        assert np.array_equal(getattr(bg, k), getattr(ref, k)), k
    assert bg.batch_num_nodes == ref.batch_num_nodes and bg.batch_num_edges == ref.batch_num_edges
    if not hdsg:
        csc = gb.derive_csc(bg)
        assert np.array_equal(csc["wnode_id"], z["hsg_wnode_id"])
        assert np.array_equal(csc["snode_id"], z["hsg_snode_id"])
        assert np.array_equal(np.sort(csc["super_eid"]), z["hsg_wsedge_id"])
        assert np.array_equal(np.sort(csc["word_eid"]), z["hsg_swedge_id"])


def test_edge_id_formulas():
    """closed-form DGL edge ids of the HSG builder vs the literal restatement."""
    z = dict(np.load(os.path.join(GOLD, "builder.npz")))
    exs, graphs = build_graphs(z, "hsg", False)
    for g in graphs:
        N = g.n_sent
        ws = np.nonzero((g.unit[g.src] == 0) & (g.unit[g.dst] == 1))[0]
        sent_of = g.dst[ws] - (g.n_nodes - N)
        k = np.bincount(sent_of, minlength=N)
        base = np.concatenate([[0], np.cumsum(2 * k + 2 * N)[:-1]])
        t = np.concatenate([np.arange(x) for x in k]) if len(ws) else np.zeros(0, np.int64)
        assert np.array_equal(ws, base[sent_of] + 2 * t)
        # s_i -> s_j at base_i + 2k_i + j ; s_j -> s_i at base_i + 2k_i + N + j
        for i in range(N):
            e0 = base[i] + 2 * k[i]
            assert np.array_equal(g.dst[e0:e0 + N], np.arange(N) + g.n_nodes - N)
            assert np.array_equal(g.src[e0 + N:e0 + 2 * N], np.arange(N) + g.n_nodes - N)


def test_rounding_half_even_in_packer():
    z = dict(np.load(os.path.join(GOLD, "builder.npz")))
    exs = fx.examples_from_arrays(z, "hsg")
    e = exs[-1]                                     # the hand-made edge-case example
    tb = syn.pack_token_batch([e])
    row0 = {int(w): int(b) for w, b in zip(e.sents[0], tb.sent_bin[0]) if w != 0}
    assert row0[1000] == 0 and row0[1001] == 2 and row0[1002] == 2 and row0[1003] == 9   # 0.5->0, 1.5->2, 2.5->2
    assert row0[1016] == -1 and row0[1] == -1       # not TF-IDF keys


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_reference_golden(name):
    z, g, params, gold_grads = load_fixture(name)
    n_iter = int(z["dims"][5])
    for p in params.values():
        p.requires_grad_(True)
    w = torch.from_numpy(z["in_w"]).requires_grad_(True)
    s = torch.from_numpy(z["in_s"]).requires_grad_(True)
    ws, ss = wr.update_loop(g, w, s, params, n_iter)
    assert nerr(ws, z["out_w"]) <= TOL and nerr(ss, z["out_s"]) <= TOL
    loss = (ws * torch.from_numpy(z["cw"])).sum() + (ss * torch.from_numpy(z["cs"])).sum()
    loss.backward()
    assert nerr(w.grad, z["grad_in_w"]) <= TOL and nerr(s.grad, z["grad_in_s"]) <= TOL
    for k, p in params.items():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert nerr(got, gold_grads[k]) <= TOL, k
    # (ii) attn_fc.weight[:, d:2d].grad == 0 exactly: the destination z is DGL's zero fill
    for k, gr in gold_grads.items():
        if k.endswith("attn_fc.weight"):
            d = gr.shape[1] // 3
            assert float(gr[:, d:2 * d].abs().max()) == 0.0


@pytest.mark.parametrize("name", FIXTURES)
def test_closed_form_matches_reference_golden(name):
    z, g, params, gold_grads = load_fixture(name)
    n_iter = int(z["dims"][5])
    for p in params.values():
        p.requires_grad_(True)
    csc = gb.derive_csc(g)
    w = torch.from_numpy(z["in_w"]).requires_grad_(True)
    s = torch.from_numpy(z["in_s"]).requires_grad_(True)
    ws, ss = cf.update_loop_cf(csc, w, s, params, n_iter)
    assert nerr(ws, z["out_w"]) <= TOL and nerr(ss, z["out_s"]) <= TOL
    loss = (ws * torch.from_numpy(z["cw"])).sum() + (ss * torch.from_numpy(z["cs"])).sum()
    loss.backward()
    assert nerr(w.grad, z["grad_in_w"]) <= TOL and nerr(s.grad, z["grad_in_s"]) <= TOL
    for k, p in params.items():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        assert nerr(got, gold_grads[k]) <= TOL, k


def _tiny_problem(seed=0, hdsg=False):
    exs = syn.make_examples(3, "tiny", seed=seed, hdsg=hdsg)
    filt = set(syn.filter_ids().tolist())
    if hdsg:
        graphs = [gb.create_graph_hdsg(e.doc_len, e.sents.tolist(), e.doc_tokens, e.w2s, e.w2d, filt) for e in exs]
    else:
        graphs = [gb.create_graph_hsg(e.sents.tolist(), e.w2s, filt) for e in exs]
    bg, _ = gb.collate(graphs)
    return bg


def _rand_params(seed, emb=48, hid=16, nh=4, ffn=32, fe=10):
    gen = torch.Generator().manual_seed(seed)
    p = {"_TFembed.weight": torch.randn(10, fe, generator=gen)}
    for pre, (i, o, H, bias) in {"word2sent.": (emb, hid, nh, False), "sent2word.": (hid, emb, 6, True)}.items():
        d = o // H
        for k in range(H):
            p[pre + "layer.heads.%d.fc.weight" % k] = torch.randn(d, i, generator=gen) * 0.2
            p[pre + "layer.heads.%d.feat_fc.weight" % k] = torch.randn(d, fe, generator=gen) * 0.2
            if bias:
                p[pre + "layer.heads.%d.feat_fc.bias" % k] = torch.randn(d, generator=gen) * 0.2
            p[pre + "layer.heads.%d.attn_fc.weight" % k] = torch.randn(1, 3 * d, generator=gen) * 0.5
        p[pre + "ffn.w_1.weight"] = torch.randn(ffn, o, 1, generator=gen) * 0.2
        p[pre + "ffn.w_1.bias"] = torch.randn(ffn, generator=gen) * 0.1
        p[pre + "ffn.w_2.weight"] = torch.randn(o, ffn, 1, generator=gen) * 0.2
        p[pre + "ffn.w_2.bias"] = torch.randn(o, generator=gen) * 0.1
        p[pre + "ffn.layer_norm.weight"] = torch.rand(o, generator=gen) + 0.5
        p[pre + "ffn.layer_norm.bias"] = torch.randn(o, generator=gen) * 0.1
    return p


@pytest.mark.parametrize("hdsg", [False, True])
def test_bucketed_oracle_equals_closed_form(hdsg):
    bg = _tiny_problem(5, hdsg)
    params = _rand_params(1)
    csc = gb.derive_csc(bg)
    nw, ns = int((bg.unit == 0).sum()), int((bg.unit == 1).sum())
    w, s = torch.randn(nw, 48), torch.randn(ns, 16)
    a = wr.update_loop(bg, w, s, params, 2)
    b = cf.update_loop_cf(csc, w, s, params, 2)
    assert nerr(a[0], b[0]) <= TOL and nerr(a[1], b[1]) <= TOL


def test_known_answers_zero_attention_and_isolated_nodes():
    """§8-c (iii) all-zero attention => sh = mean_act(z) * deg/(deg+x); (iv) isolated word => S2W row == FFN(origin);
    (v) sentence with zero word edges => W2S aggregation 0."""
    bg = _tiny_problem(7, False)
    params = _rand_params(2)
    for k in list(params):
        if k.endswith("attn_fc.weight"):
            params[k] = torch.zeros_like(params[k])
    csc = gb.derive_csc(bg)
    nw, ns = int((bg.unit == 0).sum()), int((bg.unit == 1).sum())
    w, s = torch.randn(nw, 48), torch.randn(ns, 16)
    te = wr.tfidf_embed(bg, params["_TFembed.weight"])
    sh = wr.multi_head(bg, w, params, "word2sent.layer.", "W2S", te)
    W, _, _, _ = cf.pack_layer(params, "word2sent.layer.", 4)
    z = w @ W.t()
    ip, src, ex = csc["super_indptr"], csc["super_src"], csc["extra_cnt"]
    for v in range(ns):
        deg = ip[v + 1] - ip[v]
        if deg == 0:
            assert float(sh[v].abs().max()) == 0.0
            continue
        want = z[src[ip[v]:ip[v + 1]]].mean(0) * deg / (deg + ex[v])
        assert torch.allclose(sh[v], want, atol=1e-5)
    # isolated word nodes (no in-edges): S2W output row == FFN(origin row)
    wip = csc["word_indptr"]
    iso = np.nonzero(wip[1:] == wip[:-1])[0]
    assert len(iso) > 0, "generator should produce isolated word nodes (UNK / non-key tokens)"
    out = wr.wswgat(bg, w, s, params, "sent2word.", "S2W", te)
    want = wr.ffn(w[iso].unsqueeze(0), params, "sent2word.ffn.").squeeze(0)
    assert torch.allclose(out[iso], want, atol=1e-6)


def test_edge_permutation_invariance():
    """§8-c (vi): permuting the word pairs inside a sentence changes nothing beyond 1e-6."""
    bg = _tiny_problem(9, False)
    params = _rand_params(3)
    csc = gb.derive_csc(bg)
    nw, ns = int((bg.unit == 0).sum()), int((bg.unit == 1).sum())
    w, s = torch.randn(nw, 48), torch.randn(ns, 16)
    a = cf.update_loop_cf(csc, w, s, params, 1)
    rng = np.random.default_rng(0)
    csc2 = dict(csc)
    for name in ("super", "word"):
        ip = csc[name + "_indptr"]
        perm = np.concatenate([ip[v] + rng.permutation(ip[v + 1] - ip[v]) for v in range(len(ip) - 1)]).astype(np.int64)
        csc2[name + "_src"], csc2[name + "_bin"] = csc[name + "_src"][perm], csc[name + "_bin"][perm]
    b = cf.update_loop_cf(csc2, w, s, params, 1)
    assert nerr(a[0], b[0]) <= 1e-6 and nerr(a[1], b[1]) <= 1e-6


def test_topm_and_loss_restatement():
    bg = _tiny_problem(11, False)
    ns = int((bg.ndtype == 1).sum())
    logits = torch.randn(ns, 2)
    labels = torch.randint(0, 2, (ns,))
    loss = wr.graph_loss(bg, logits, labels)
    ce = torch.nn.functional.cross_entropy(logits, labels, reduction="sum") / len(bg.batch_num_nodes)
    assert torch.allclose(loss, ce, atol=1e-6)
    idx = wr.topm_indices(bg, logits, 3)
    assert len(idx) == len(bg.batch_num_nodes) and all(len(i) <= 3 for i in idx)


@pytest.mark.skipif(not os.path.isdir("/root/reference/module"), reason="live reference only in the build container")
def test_oracle_vs_live_reference():
    """Re-run the UNMODIFIED reference WSWGAT on the shim and compare with the restatement."""
    sys.path.insert(0, "/root/reference")
    from oracle import dgl04_shim as shim
    shim.install()
    from module.GAT import WSWGAT
    sys.path.insert(0, os.path.join(GOLD))
    import make_golden as mg
    exs = syn.make_examples(3, "tiny", seed=31)
    filt = set(syn.filter_ids().tolist())
    graphs = [mg.ref_graph_hsg(e, filt) for e in exs]
    order = gb.stable_desc_order([e.n_sent for e in exs]).tolist()
    BG = shim.batch([graphs[i] for i in order])
    ga = mg.shim_to_arrays(BG)
    torch.manual_seed(5)
    w2s = WSWGAT(48, 16, 4, 0.1, 32, 0.1, 10, "W2S").eval()
    s2w = WSWGAT(16, 48, 6, 0.1, 32, 0.1, 10, "S2W").eval()
    T = torch.nn.Embedding(10, 10)
    nw, ns = int((ga.unit == 0).sum()), int((ga.unit == 1).sum())
    w, s = torch.randn(nw, 48), torch.randn(ns, 16)
    with torch.no_grad():
        rw, rs = mg.run_reference_loop(BG, w2s, s2w, T, w, s, 2)
        params = {"word2sent." + k: v for k, v in w2s.state_dict().items()}
        params.update({"sent2word." + k: v for k, v in s2w.state_dict().items()})
        params["_TFembed.weight"] = T.weight
        ow, os_ = wr.update_loop(ga, w, s, params, 2)
    assert nerr(ow, rw) <= 1e-6 and nerr(os_, rs) <= 1e-6


@pytest.mark.parametrize("name", ["s2s_hsg.npz", "s2s_hdsg.npz"])
def test_s2s_closed_form_matches_reference_golden(name):
    """S2S layer type (SGATLayer, GATLayer.py:49-78; never instantiated by the reference's models): the closed form
    of oracle/closed_form.py against the reference's own classes run on the shim (tests/golden/make_golden_s2s.py)."""
    import os
    from oracle import closed_form as cf
    from oracle import fixtures as fx
    from oracle import graph_builder_ref as gb
    z = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)))
    g = fx.graph_from_arrays(z, "g_")
    csc = gb.derive_csc(g)
    params = {"x." + k[2:]: torch.from_numpy(v).clone().requires_grad_(True) for k, v in z.items() if k.startswith("p:")}
    s = torch.from_numpy(z["in_s"]).clone().requires_grad_(True)
    out = cf.s2s_cf(g, csc, s, params, "x.")
    (out * torch.from_numpy(z["cs"])).sum().backward()

    def err(a, b):
        return float((a - b).abs().max() / (b.abs().max() + 1e-30))
    assert err(out.detach(), torch.from_numpy(z["out_s"])) <= 1e-6
    assert err(s.grad, torch.from_numpy(z["grad_in_s"])) <= 1e-6
    for k, v in z.items():
        if k.startswith("gp:"):
            ref = torch.from_numpy(v)
            got = params["x." + k[3:]].grad
            got = torch.zeros_like(ref) if got is None else got
            if float(ref.abs().max()) == 0.0:
                assert float(got.abs().max()) == 0.0, k
            else:
                assert err(got, ref) <= 1e-5, k
