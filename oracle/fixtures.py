"""(De)serialisation of examples / graphs for tests/golden (TEST INFRASTRUCTURE)."""
import numpy as np

from hetersumgraph_b200.synthetic import DocExample


def examples_to_arrays(examples, prefix="ex"):
    out = {}
    out[prefix + "_n"] = np.asarray([len(examples)], np.int64)
    for i, e in enumerate(examples):
        p = "%s%d_" % (prefix, i)
        out[p + "sents"] = e.sents.astype(np.int32)
        out[p + "labels"] = e.labels.astype(np.int64)
        si, wi, tv = [], [], []
        for s, d in enumerate(e.w2s):
            for w, v in d.items():
                si.append(s), wi.append(w), tv.append(v)
        out[p + "w2s_s"] = np.asarray(si, np.int64)
        out[p + "w2s_w"] = np.asarray(wi, np.int64)
        out[p + "w2s_v"] = np.asarray(tv, np.float64)
        if e.doc_len is not None:
            out[p + "doc_len"] = np.asarray(e.doc_len, np.int64)
            out[p + "doc_tok"] = np.asarray([t for d in e.doc_tokens for t in d], np.int64)
            out[p + "doc_tok_ptr"] = np.cumsum([0] + [len(d) for d in e.doc_tokens]).astype(np.int64)
            di, wi, tv = [], [], []
            for j, d in enumerate(e.w2d):
                for w, v in d.items():
                    di.append(j), wi.append(w), tv.append(v)
            out[p + "w2d_d"] = np.asarray(di, np.int64)
            out[p + "w2d_w"] = np.asarray(wi, np.int64)
            out[p + "w2d_v"] = np.asarray(tv, np.float64)
    return out


def examples_from_arrays(z, prefix="ex"):
    n = int(z[prefix + "_n"][0])
    out = []
    for i in range(n):
        p = "%s%d_" % (prefix, i)
        sents = z[p + "sents"]
        w2s = [dict() for _ in range(sents.shape[0])]
        for s, w, v in zip(z[p + "w2s_s"].tolist(), z[p + "w2s_w"].tolist(), z[p + "w2s_v"].tolist()):
            w2s[s][w] = v
        e = DocExample(sents=sents, w2s=w2s, labels=z[p + "labels"])
        if (p + "doc_len") in z:
            e.doc_len = z[p + "doc_len"].tolist()
            ptr = z[p + "doc_tok_ptr"].tolist()
            tok = z[p + "doc_tok"].tolist()
            e.doc_tokens = [tok[ptr[j]:ptr[j + 1]] for j in range(len(ptr) - 1)]
            e.w2d = [dict() for _ in range(len(e.doc_len))]
            for j, w, v in zip(z[p + "w2d_d"].tolist(), z[p + "w2d_w"].tolist(), z[p + "w2d_v"].tolist()):
                e.w2d[j][w] = v
        out.append(e)
    return out


GRAPH_KEYS = ("unit", "ndtype", "wid", "src", "dst", "tffrac", "etype")


def graph_to_arrays(g, prefix="g_"):
    out = {prefix + k: getattr(g, k) for k in GRAPH_KEYS}
    out[prefix + "batch_num_nodes"] = np.asarray(g.batch_num_nodes, np.int64)
    out[prefix + "batch_num_edges"] = np.asarray(g.batch_num_edges, np.int64)
    return out


def graph_from_arrays(z, prefix="g_"):
    from oracle.graph_builder_ref import GraphArrays
    g = GraphArrays()
    for k in GRAPH_KEYS:
        setattr(g, k, z[prefix + k])
    g.batch_num_nodes = z[prefix + "batch_num_nodes"].tolist()
    g.batch_num_edges = z[prefix + "batch_num_edges"].tolist()
    return g


# --------------------------------------------------------------------------------------------
# sentence-encoder fixtures (tests/golden/make_golden_encoder.py and the tests share these)
# --------------------------------------------------------------------------------------------
def encoder_param_shapes(vocab, emb, sent_max_len, doc_max, n_feature, hidden, lstm_hidden, lstm_layers=2):
    """state_dict keys / shapes of the sentence-encoder part of the reference's HSumGraph (HiGraph.py:112-125,
    Encoder.py:41-54, HiGraph.py:53), bidirectional LSTM."""
    shapes = {"ngram_enc.embed.weight": (vocab, emb), "sent_pos_embed.weight": (doc_max + 1, emb),
              "ngram_enc.position_embedding.weight": (sent_max_len + 1, emb),
              "cnn_proj.weight": (n_feature, emb), "cnn_proj.bias": (n_feature,),
              "lstm_proj.weight": (n_feature, 2 * lstm_hidden), "lstm_proj.bias": (n_feature,),
              "n_feature_proj.weight": (hidden, 2 * n_feature)}
    for i, h in enumerate(range(2, 8)):
        shapes["ngram_enc.convs.%d.weight" % i] = (50, 1, h, emb)
        shapes["ngram_enc.convs.%d.bias" % i] = (50,)
    for layer in range(lstm_layers):
        for sfx in ("", "_reverse"):
            n_in = emb if layer == 0 else 2 * lstm_hidden
            shapes["lstm.weight_ih_l%d%s" % (layer, sfx)] = (4 * lstm_hidden, n_in)
            shapes["lstm.weight_hh_l%d%s" % (layer, sfx)] = (4 * lstm_hidden, lstm_hidden)
            shapes["lstm.bias_ih_l%d%s" % (layer, sfx)] = (4 * lstm_hidden,)
            shapes["lstm.bias_hh_l%d%s" % (layer, sfx)] = (4 * lstm_hidden,)
    return shapes


def seeded_encoder_params(shapes, seed, zero_pad_row=True):
    """Deterministic parameter values from the shapes alone (so that a fixture need not store megabytes of weights):
    tensor number i is drawn from its own torch.Generator(seed * 1000 + i).  The two sinusoid tables are the
    reference's (frozen, PositionEmbedding.py); weights ~ N(0, 1/fan_in), biases ~ N(0, 0.1), embedding ~ N(0, 1)."""
    import torch
    from oracle import encoder_ref as er
    out = {}
    for i, (k, shp) in enumerate(sorted(shapes.items())):
        g = torch.Generator().manual_seed(seed * 1000 + i)
        if k in ("sent_pos_embed.weight", "ngram_enc.position_embedding.weight"):
            out[k] = er.sinusoid_table(shp[0], shp[1], padding_idx=0)
        elif k == "ngram_enc.embed.weight":
            w = torch.randn(shp, generator=g)
            if zero_pad_row:
                w[0] = 0.0
            out[k] = w
        elif len(shp) == 1:
            out[k] = 0.1 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            out[k] = torch.randn(shp, generator=g) / fan_in ** 0.5
    return out


def encoder_tokens(n_sent_per_graph, L, vocab, seed):
    """Token matrix [S, L] (trailing PAD = 0) with the edge cases of the n-gram encoder: an empty sentence, a full one,
    lengths within 7 of L, length 1; graph_sent_ptr for graphs given in batch order."""
    import numpy as np
    rng = np.random.default_rng(seed)
    S = int(sum(n_sent_per_graph))
    lens = rng.integers(1, max(2, L // 2), size=S)
    special = [0, L, L - 1, L - 6, L - 7, 1, 2, 7]
    for i, v in enumerate(special[:S]):
        lens[(3 * i + 1) % S] = max(0, min(L, v))
    tokens = np.zeros((S, L), np.int32)
    for s in range(S):
        tokens[s, :lens[s]] = rng.integers(1, vocab, size=lens[s])
    ptr = np.concatenate([[0], np.cumsum(n_sent_per_graph)]).astype(np.int32)
    return tokens, ptr


def load_encoder_fixture(gold_dir, name):
    """(arrays, seeded parameters) of tests/golden/encoder_*.npz"""
    import os
    import numpy as np
    z = dict(np.load(os.path.join(gold_dir, name)))
    vocab, emb, L, doc_max, n_feature, hidden, lstm_hidden = [int(v) for v in z["dims"]]
    shapes = encoder_param_shapes(vocab, emb, L, doc_max, n_feature, hidden, lstm_hidden)
    params = seeded_encoder_params(shapes, int(z["seed"]), bool(z["zero_pad_row"]))
    return z, params


def golden_grad(z, key, g):
    """(gradient as stored, golden gradient): the encoder fixtures store large gradients sub-sampled"""
    import torch
    ref = z["gp:" + key]
    g = g.detach()
    return (g if ref.shape == tuple(g.shape) else g.flatten()[::int(z["stride"])]), torch.from_numpy(ref)


def seeded_state_dict(shapes, seed, keep=()):
    """Deterministic values for a whole-model state_dict from (key -> shape) alone: tensor number i (keys sorted) is
    drawn from torch.Generator(seed * 100003 + i).  Matrices ~ N(0, 1/fan_in), vectors ~ N(0, 0.1), LayerNorm gains
    1 + N(0, 0.1), word embedding ~ N(0, 1) with a zero PAD row; keys in `keep` (frozen sinusoid tables) are skipped."""
    import torch
    out = {}
    for i, (k, shp) in enumerate(sorted(shapes.items())):
        if k in keep:
            continue
        g = torch.Generator().manual_seed(seed * 100003 + i)
        if k.endswith("embed.weight") and "TFembed" not in k and "pos" not in k and "position" not in k:
            w = torch.randn(shp, generator=g)
            w[0] = 0.0
        elif k.endswith("layer_norm.weight"):
            w = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 1:
            w = 0.1 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            w = torch.randn(shp, generator=g) / fan_in ** 0.5
        out[k] = w
    if "_embed.weight" in out and "ngram_enc.embed.weight" in out:
        out["ngram_enc.embed.weight"] = out["_embed.weight"]          # one shared table (HiGraph.py:125)
    return out


FROZEN_MODEL_KEYS = ("sent_pos_embed.weight", "ngram_enc.position_embedding.weight")


def time_verbatim_reference(exs, tb, hdsg, n_iter, steps=1):
    """B-ref of BASELINE.md 3 (test infrastructure; build container only - needs /root/reference): the reference's OWN
    WSWGAT modules (module/GAT.py:45-59 over GATStackLayer.py / GATLayer.py) run UNMODIFIED on the DGL-0.4-semantics
    shim, fwd+bwd of the update loop (HiGraph.py:98-106) on the given batch, dropout 0, at 1 thread and at all host
    threads.  Graph construction (reference CreateGraph on the shim) is timed separately, once."""
    import importlib.util
    import os
    import time

    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("hsg_make_golden", os.path.join(root, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)                     # installs the shim, imports the reference from /root/reference
    from hetersumgraph_b200 import synthetic as syn
    filt = set(syn.filter_ids().tolist())
    t0 = time.time()
    graphs = [(mg.ref_graph_hdsg if hdsg else mg.ref_graph_hsg)(e, filt) for e in exs]
    BG = mg.shim.batch([graphs[i] for i in tb.order])
    build_s = time.time() - t0
    ga = mg.shim_to_arrays(BG)
    torch.manual_seed(1234)
    w2s = mg.WSWGAT(300, 64, 8, 0.0, 512, 0.0, 50, "W2S").train()
    s2w = mg.WSWGAT(64, 300, 6, 0.0, 512, 0.0, 50, "S2W").train()
    T = torch.nn.Embedding(10, 50)
    nw, ns = int((ga.unit == 0).sum()), int((ga.unit == 1).sum())
    w = torch.randn(nw, 300, requires_grad=True)
    s = torch.randn(ns, 64, requires_grad=True)
    cw, cs = torch.randn(nw, 300) / len(exs), torch.randn(ns, 64) / len(exs)
    params = list(w2s.parameters()) + list(s2w.parameters()) + list(T.parameters())

    base_n, base_e = set(BG.ndata.keys()), set(BG.edata.keys())

    def step():
        for p in params + [w, s]:
            p.grad = None
        # a fresh step starts from the graph as the data loader delivers it: fields written by the previous step
        # (tfidfembed, e, ...) carry that step's autograd history
        for view, base in ((BG.ndata, base_n), (BG.edata, base_e)):
            for key in [k for k in view.keys() if k not in base]:
                view.pop(key)
        ws_, ss_ = mg.run_reference_loop(BG, w2s, s2w, T, w, s, n_iter)
        ((ws_ * cw).sum() + (ss_ * cs).sum()).backward()

    out = {"what": "reference WSWGAT modules verbatim on the DGL-0.4-semantics shim (real DGL not installable), update "
                   "loop fwd+bwd, dropout 0", "graphs": len(exs), "word_nodes": nw, "supernodes": ns,
           "graph_build_s_once": build_s, "host_cores": os.cpu_count()}
    all_threads = torch.get_num_threads()
    try:
        for label, nt in (("threads_1", 1), ("threads_all", os.cpu_count() or all_threads)):
            torch.set_num_threads(nt)
            step()
            t0 = time.time()
            for _ in range(steps):
                step()
            sec = (time.time() - t0) / steps
            out[label] = {"threads": nt, "s_per_step": sec, "graphs_per_s": len(exs) / sec}
    finally:
        torch.set_num_threads(all_threads)
    return out
