"""Generate tests/golden/*.npz by running the UNMODIFIED reference on the DGL-0.4 shim.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference's own classes are executed verbatim:
  * module/dataloader.py  ExampleSet.CreateGraph / MultiExampleSet.CreateGraph / graph_collate_fn order
  * module/GAT.py         WSWGAT (W2S / S2W), incl. GATStackLayer.py / GATLayer.py
  * HiGraph.py            the update loop order (W2S, then n_iter x (S2W, W2S)) and set_wnfeature's
                          tfidfembed write (reproduced here call by call on the shim graph)
Outputs (inputs + reference results) are committed so that the CPU and GPU test
suites can run where /root/reference does not exist.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import dgl04_shim as shim  # noqa: E402

shim.install()
from module import dataloader as refdl  # noqa: E402
from module.GAT import WSWGAT  # noqa: E402

from hetersumgraph_b200 import synthetic as syn  # noqa: E402
from oracle import fixtures as fx  # noqa: E402
from oracle import graph_builder_ref as gb  # noqa: E402


class _Vocab:
    def id2word(self, i):
        return "w%d" % i

    def word2id(self, w):
        return int(w[1:])


def _str_keys(dicts):
    return {str(i): {"w%d" % k: v for k, v in d.items()} for i, d in enumerate(dicts)}


def ref_graph_hsg(e, filt):
    ds = object.__new__(refdl.ExampleSet)
    ds.vocab, ds.filterids = _Vocab(), list(filt)
    lab = np.zeros((e.n_sent, 50), np.int64)
    return ds.CreateGraph(e.sents.tolist(), lab, _str_keys(e.w2s))


def ref_graph_hdsg(e, filt):
    ds = object.__new__(refdl.MultiExampleSet)
    ds.vocab, ds.filterids = _Vocab(), list(filt)
    lab = np.zeros((e.n_sent, 50), np.int64)
    return ds.CreateGraph(e.doc_len, e.sents.tolist(), e.doc_tokens, lab, _str_keys(e.w2s), _str_keys(e.w2d))


def shim_to_arrays(G):
    g = gb.GraphArrays()
    s, d = G.edges_arrays()
    g.src, g.dst = s.numpy().copy(), d.numpy().copy()
    g.unit = G.ndata["unit"].numpy().astype(np.int64)
    g.ndtype = G.ndata["dtype"].numpy().astype(np.int64)
    g.wid = G.ndata["id"].numpy().astype(np.int64)
    g.tffrac = G.edata["tffrac"].numpy().astype(np.int64)
    g.etype = G.edata["dtype"].numpy().astype(np.int64)
    g.batch_num_nodes = list(G.batch_num_nodes) if G.batch_num_nodes else [G.number_of_nodes()]
    g.batch_num_edges = list(G.batch_num_edges) if G.batch_num_edges else [G.number_of_edges()]
    return g


def edge_case_examples(L=100):
    """Hand-made examples covering the §8-c(vii) cases."""
    def pad(x):
        return x + [0] * (L - len(x))
    # ids: 0 PAD (filtered), 1 UNK (node, never a key), 4..203 stop words (filtered), 1000+ content words
    s0 = [1000, 1001, 1000, 1002, 1, 5, 1001, 1003, 1016]      # duplicates, UNK, stop word 5, non-key 1016
    s1 = [5, 6, 7]                                            # only filtered words -> zero word edges
    s2 = [1002, 1004, 1005, 1006, 1002, 1007]
    s3 = [1, 1, 1]                                            # only UNK -> node exists, no edges
    sents = np.asarray([pad(s0), pad(s1), pad(s2), pad(s3)], np.int32)
    w2s = [
        {1000: 0.5 / 9, 1001: 1.5 / 9, 1002: 2.5 / 9, 1003: 1.0, 5: 0.3},   # .5 roundings; 1016, 1 missing
        {5: 0.5, 6: 0.5, 7: 0.7},
        {1002: 0.449, 1004: 3.5 / 9, 1005: 0.0001, 1006: 8.5 / 9, 1007: 0.999},
        {},
    ]
    labels = np.asarray([1, 0, 1, 0], np.int64)
    e = syn.DocExample(sents=sents, w2s=w2s, labels=labels)
    e2 = syn.DocExample(sents=sents.copy(), w2s=w2s, labels=labels)
    e2.doc_len = [3, 1]
    e2.doc_tokens = [s0 + s1 + s2, s3]
    e2.w2d = [{1000: 0.2, 1002: 4.5 / 9, 1004: 0.61, 5: 0.1, 1007: 0.05}, {}]
    return e, e2


def builder_golden():
    filt = set(syn.filter_ids().tolist())
    out = {}
    # HSG: 5 synthetic + edge-case example, ties in #sentences to exercise the stable order
    exs = syn.make_examples(5, "cnndm", seed=11)
    exs[3].sents = exs[3].sents[:exs[1].n_sent] if exs[3].n_sent > exs[1].n_sent else exs[3].sents
    exs[3].w2s = exs[3].w2s[:exs[3].n_sent]
    exs[3].labels = exs[3].labels[:exs[3].n_sent]
    edge_hsg, edge_hdsg = edge_case_examples()
    exs.append(edge_hsg)
    graphs = [ref_graph_hsg(e, filt) for e in exs]
    order = gb.stable_desc_order([e.n_sent for e in exs]).tolist()
    BG = shim.batch([graphs[i] for i in order])
    out.update(fx.examples_to_arrays(exs, "hsg"))
    out.update(fx.graph_to_arrays(shim_to_arrays(BG), "hsg_g_"))
    out["hsg_order"] = np.asarray(order, np.int64)
    # filter ids as seen by WSGATLayer etc.
    out["hsg_wnode_id"] = BG.filter_nodes(lambda n: n.data["unit"] == 0).numpy()
    out["hsg_snode_id"] = BG.filter_nodes(lambda n: n.data["unit"] == 1).numpy()
    out["hsg_wsedge_id"] = BG.filter_edges(lambda e: (e.src["unit"] == 0) & (e.dst["unit"] == 1)).numpy()
    out["hsg_swedge_id"] = BG.filter_edges(lambda e: (e.src["unit"] == 1) & (e.dst["unit"] == 0)).numpy()
    # HDSG
    exd = syn.make_examples(3, "multinews", seed=12, hdsg=True)
    exd.append(edge_hdsg)
    graphs = [ref_graph_hdsg(e, filt) for e in exd]
    order = gb.stable_desc_order([e.n_sent for e in exd]).tolist()
    BG = shim.batch([graphs[i] for i in order])
    out.update(fx.examples_to_arrays(exd, "hdsg"))
    out.update(fx.graph_to_arrays(shim_to_arrays(BG), "hdsg_g_"))
    out["hdsg_order"] = np.asarray(order, np.int64)
    np.savez_compressed(os.path.join(HERE, "builder.npz"), **out)
    print("builder.npz", {k: v.shape for k, v in out.items() if k.endswith("g_src")})


def run_reference_loop(BG, w2s, s2w, T, w, s, n_iter):
    """HiGraph.py:144-152 (tfidfembed write) + :98-106 (update loop), on the shim graph."""
    eid = BG.filter_edges(lambda edges: edges.data["dtype"] == 0)
    BG.edges[eid].data["tfidfembed"] = T(BG.edges[eid].data["tffrac"])
    word_state = w
    sent_state = w2s(BG, w, s)
    for _ in range(n_iter):
        word_state = s2w(BG, word_state, sent_state)
        sent_state = w2s(BG, word_state, sent_state)
    return word_state, sent_state


def wswgat_golden(name, exs, hdsg, dims, n_iter, seed):
    emb, hid, nh, ffn_h, fe = dims
    filt = set(syn.filter_ids().tolist())
    graphs = [(ref_graph_hdsg if hdsg else ref_graph_hsg)(e, filt) for e in exs]
    order = gb.stable_desc_order([e.n_sent for e in exs]).tolist()
    BG = shim.batch([graphs[i] for i in order])
    ga = shim_to_arrays(BG)
    torch.manual_seed(seed)
    w2s = WSWGAT(emb, hid, nh, 0.1, ffn_h, 0.1, fe, "W2S").eval()
    s2w = WSWGAT(hid, emb, 6, 0.1, ffn_h, 0.1, fe, "S2W").eval()
    T = torch.nn.Embedding(10, fe)
    # make LayerNorm affine / biases non-trivial
    with torch.no_grad():
        for m in (w2s, s2w):
            m.ffn.layer_norm.weight.uniform_(0.5, 1.5)
            m.ffn.layer_norm.bias.uniform_(-0.2, 0.2)
    nw, ns = int((ga.unit == 0).sum()), int((ga.unit == 1).sum())
    w = torch.randn(nw, emb, requires_grad=True)
    s = torch.randn(ns, hid, requires_grad=True)
    cw = torch.randn(nw, emb)
    cs = torch.randn(ns, hid)
    word_state, sent_state = run_reference_loop(BG, w2s, s2w, T, w, s, n_iter)
    loss = (word_state * cw).sum() + (sent_state * cs).sum()
    loss.backward()
    out = {}
    out.update(fx.examples_to_arrays(exs, "ex"))
    out.update(fx.graph_to_arrays(ga, "g_"))
    out["order"] = np.asarray(order, np.int64)
    out["dims"] = np.asarray(list(dims) + [n_iter, int(hdsg)], np.int64)
    out["in_w"], out["in_s"] = w.detach().numpy(), s.detach().numpy()
    out["cw"], out["cs"] = cw.numpy(), cs.numpy()
    out["out_w"], out["out_s"] = word_state.detach().numpy(), sent_state.detach().numpy()
    out["grad_in_w"], out["grad_in_s"] = w.grad.numpy(), s.grad.numpy()
    params = {"word2sent." + k: v for k, v in w2s.state_dict(keep_vars=True).items()}
    params.update({"sent2word." + k: v for k, v in s2w.state_dict(keep_vars=True).items()})
    params["_TFembed.weight"] = T.weight
    for k, v in params.items():
        out["p:" + k] = v.detach().numpy()
        out["gp:" + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "Nw", nw, "Ns", ns, "E_all", ga.n_edges, "loss", float(loss))


def main():
    builder_golden()
    # default dims (train.py:279-309): emb 300, hidden 64, 8 heads, ffn 512, feat_embed 50; HSG, n_iter=1
    exs = syn.make_examples(3, "cnndm", seed=21)
    for e in exs:                                  # keep the fixture small: <= 8 sentences per doc
        n = min(e.n_sent, 8)
        e.sents, e.w2s, e.labels = e.sents[:n], e.w2s[:n], e.labels[:n]
    exs.append(edge_case_examples()[0])
    wswgat_golden("wswgat_hsg_default.npz", exs, False, (300, 64, 8, 512, 50), 1, 1234)
    # small dims, HDSG graph (doc nodes, sent->doc extras), n_iter=2
    exd = syn.make_examples(4, "tiny", seed=22, hdsg=True)
    exd.append(edge_case_examples()[1])
    wswgat_golden("wswgat_hdsg_small.npz", exd, True, (48, 16, 4, 32, 10), 2, 4321)
    # small dims, HSG graph, n_iter=3
    exs2 = syn.make_examples(5, "tiny", seed=23)
    wswgat_golden("wswgat_hsg_small.npz", exs2, False, (48, 16, 4, 32, 10), 3, 99)


if __name__ == "__main__":
    main()
