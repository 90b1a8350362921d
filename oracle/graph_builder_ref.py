"""CPU restatement of the reference's graph construction (TEST INFRASTRUCTURE).

Follows /root/reference/module/dataloader.py literally, but emits plain numpy
arrays (DGL node ids / edge ids in insertion order) instead of a DGLGraph:

  * add_word_nodes      <- ExampleSet.AddWordNode        dataloader.py:201-220
  * create_graph_hsg    <- ExampleSet.CreateGraph        dataloader.py:222-268
  * create_graph_hdsg   <- MultiExampleSet.CreateGraph   dataloader.py:328-406
  * map_sent2doc        <- MultiExampleSet.MapSent2Doc   dataloader.py:314-326
  * collate             <- graph_collate_fn + dgl.batch  dataloader.py:472-481
  * derive_csc          <- what WSGATLayer/SWGATLayer's filter_* + pull see
                           (GATLayer.py:105-107,113 / 143-145,149)

Pinned against the reference's own classes executed on oracle/dgl04_shim.py
(tests/test_oracle_pinning.py, tests/golden/make_golden.py).  Words are keyed by
vocabulary id instead of the word string (the reference's Vocab is a bijection,
vocabulary.py:30-89).  graph_collate_fn's torch.sort is not stable
(dataloader.py:479); this restatement and the product both use a STABLE
descending sort, and the pinning test feeds that order to the reference.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from collections import Counter

import numpy as np


class GraphArrays:
    """One (possibly batched) heterogeneous graph as flat arrays."""

    def __init__(self):
        self.unit = np.zeros(0, np.int64)      # 0 word, 1 supernode
        self.ndtype = np.zeros(0, np.int64)    # 0 word, 1 sentence, 2 document
        self.wid = np.zeros(0, np.int64)       # ndata["id"] (0 for supernodes)
        self.src = np.zeros(0, np.int64)
        self.dst = np.zeros(0, np.int64)
        self.tffrac = np.zeros(0, np.int64)
        self.etype = np.zeros(0, np.int64)
        self.batch_num_nodes = []
        self.batch_num_edges = []

    @property
    def n_nodes(self):
        return len(self.unit)

    @property
    def n_edges(self):
        return len(self.src)


def add_word_nodes(inputid, filterids):
    """dataloader.py:201-212 - word node per distinct, unfiltered id, first-occurrence order."""
    wid2nid = {}
    nid = 0
    for sentid in inputid:
        for wid in sentid:
            if wid not in filterids and wid not in wid2nid:
                wid2nid[wid] = nid
                nid += 1
    return wid2nid


def _finish(unit, ndtype, wid, src, dst, tf, et):
    g = GraphArrays()
    g.unit = np.asarray(unit, np.int64)
    g.ndtype = np.asarray(ndtype, np.int64)
    g.wid = np.asarray(wid, np.int64)
    g.src = np.asarray(src, np.int64)
    g.dst = np.asarray(dst, np.int64)
    g.tffrac = np.asarray(tf, np.int64)
    g.etype = np.asarray(et, np.int64)
    g.batch_num_nodes = [g.n_nodes]
    g.batch_num_edges = [g.n_edges]
    return g


def create_graph_hsg(input_pad, w2s, filterids):
    """dataloader.py:222-268.  w2s[i] : {wid: tfidf} for sentence i."""
    wid2nid = add_word_nodes(input_pad, filterids)
    w_nodes = len(wid2nid)
    N = len(input_pad)
    unit = [0] * w_nodes + [1] * N
    ndtype = [0] * w_nodes + [1] * N
    wid = list(wid2nid.keys()) + [0] * N
    sentid2nid = [i + w_nodes for i in range(N)]
    src, dst, tf, et = [], [], [], []
    for i in range(N):
        c = Counter(input_pad[i])                          # :247  insertion order = first occurrence
        sent_nid = sentid2nid[i]
        sent_tfw = w2s[i]
        for w in c.keys():
            if w in wid2nid and w in sent_tfw:             # :251
                box = int(np.round(sent_tfw[w] * 9))       # :253  half-to-even
                src += [wid2nid[w], sent_nid]              # :254-257  w->s then s->w
                dst += [sent_nid, wid2nid[w]]
                tf += [box, box]
                et += [0, 0]
        src += [sent_nid] * N                              # :262  s_i -> every sentence
        dst += sentid2nid
        src += sentid2nid                                  # :263  every sentence -> s_i
        dst += [sent_nid] * N
        tf += [0] * (2 * N)
        et += [1] * (2 * N)
    g = _finish(unit, ndtype, wid, src, dst, tf, et)
    g.n_sent = N
    return g


def map_sent2doc(article_len, sent_num):
    """dataloader.py:314-326 (including the early-return quirk)."""
    sent2doc = {}
    sent_no = 0
    for i in range(len(article_len)):
        for _ in range(article_len[i]):
            sent2doc[sent_no] = i
            sent_no += 1
            if sent_no > sent_num:
                return sent2doc
    return sent2doc


def create_graph_hdsg(doc_len, sent_pad, doc_pad, w2s, w2d, filterids):
    """dataloader.py:328-406.  w2d[j] : {wid: tfidf} for document j."""
    wid2nid = add_word_nodes(sent_pad, filterids)
    w_nodes = len(wid2nid)
    N = len(sent_pad)
    sentid2nid = [i + w_nodes for i in range(N)]
    ws_nodes = w_nodes + N
    sent2doc = map_sent2doc(doc_len, N)
    article_num = len(set(sent2doc.values()))
    docid2nid = [i + ws_nodes for i in range(article_num)]
    unit = [0] * w_nodes + [1] * N + [1] * article_num
    ndtype = [0] * w_nodes + [1] * N + [2] * article_num
    wid = list(wid2nid.keys()) + [0] * (N + article_num)
    src, dst, tf, et = [], [], [], []
    for i in range(N):
        c = Counter(sent_pad[i])
        sent_nid = sentid2nid[i]
        sent_tfw = w2s[i]
        for w in c.keys():
            if w in wid2nid and w in sent_tfw:
                box = int(np.round(sent_tfw[w] * 9))
                src += [wid2nid[w], sent_nid]
                dst += [sent_nid, wid2nid[w]]
                tf += [box, box]
                et += [0, 0]
        src.append(sent_nid)                               # :383-385  s -> d, dtype 2
        dst.append(docid2nid[sent2doc[i]])
        tf.append(0)
        et.append(2)
    for j in range(article_num):                           # :388-400
        c = Counter(doc_pad[j])
        doc_nid = docid2nid[j]
        doc_tfw = w2d[j]
        for w in c.keys():
            if w in wid2nid and w in doc_tfw:
                box = int(np.round(doc_tfw[w] * 9))
                src += [wid2nid[w], doc_nid]
                dst += [doc_nid, wid2nid[w]]
                tf += [box, box]
                et += [0, 0]
    g = _finish(unit, ndtype, wid, src, dst, tf, et)
    g.n_sent = N
    return g


def stable_desc_order(graph_len):
    """graph_collate_fn's sort (dataloader.py:479) made stable."""
    return np.argsort(-np.asarray(graph_len, np.int64), kind="stable")


def collate(graphs, order=None):
    """dgl.batch semantics (dataloader.py:480): concat frames, offset ids."""
    if order is None:
        order = stable_desc_order([int((g.ndtype == 1).sum()) for g in graphs])
    bg = GraphArrays()
    n_off = 0
    parts = {k: [] for k in ("unit", "ndtype", "wid", "src", "dst", "tffrac", "etype")}
    for idx in order:
        g = graphs[int(idx)]
        for k in ("unit", "ndtype", "wid", "tffrac", "etype"):
            parts[k].append(getattr(g, k))
        parts["src"].append(g.src + n_off)
        parts["dst"].append(g.dst + n_off)
        n_off += g.n_nodes
        bg.batch_num_nodes.append(g.n_nodes)
        bg.batch_num_edges.append(g.n_edges)
    for k, v in parts.items():
        setattr(bg, k, np.concatenate(v) if v else np.zeros(0, np.int64))
    return bg, [int(i) for i in order]


def derive_csc(g):
    """Known answer for the device-side CSR/CSC builder.

    Row numbering: word row = rank of the node among unit==0 nodes (ascending
    node id) = row order of `w` in WSWGAT.forward; supernode row likewise over
    unit==1 (GATLayer.py:105-106, HiGraph.py:145,193).  In-edges of each
    destination are listed in ascending DGL edge id (what `pull` sees).
    """
    unit = g.unit
    wnode = np.nonzero(unit == 0)[0]
    snode = np.nonzero(unit == 1)[0]
    row = np.zeros(g.n_nodes, np.int64)
    row[wnode] = np.arange(len(wnode))
    row[snode] = np.arange(len(snode))
    eid = np.arange(g.n_edges)
    ws = (unit[g.src] == 0) & (unit[g.dst] == 1)          # GATLayer.py:107
    sw = (unit[g.src] == 1) & (unit[g.dst] == 0)          # GATLayer.py:145
    out = {"wnode_id": wnode, "snode_id": snode,
           "sent_id": np.nonzero(g.ndtype == 1)[0], "doc_id": np.nonzero(g.ndtype == 2)[0]}

    def _csc(mask, n_dst):
        e = eid[mask]
        d = row[g.dst[e]]
        o = np.argsort(d, kind="stable")
        e = e[o]
        indptr = np.zeros(n_dst + 1, np.int64)
        np.add.at(indptr, d + 1, 1)
        indptr = np.cumsum(indptr)
        return indptr, row[g.src[e]], g.tffrac[e], e

    ip, s, b, e = _csc(ws, len(snode))
    out.update(super_indptr=ip, super_src=s, super_bin=b, super_eid=e)
    ip, s, b, e = _csc(sw, len(wnode))
    out.update(word_indptr=ip, word_src=s, word_bin=b, word_eid=e)
    # extra in-edges of supernodes (sent->sent dtype 1, sent->doc dtype 2): e = 0, z_src = 0
    extra_mask = (unit[g.dst] == 1) & ~ws
    extra = np.zeros(len(snode), np.int64)
    np.add.at(extra, row[g.dst[extra_mask]], 1)
    out["extra_cnt"] = extra
    # extra in-edges of word nodes never exist in HSG/HDSG graphs, but the contract allows them
    extra_w_mask = (unit[g.dst] == 0) & ~sw
    extra_w = np.zeros(len(wnode), np.int64)
    np.add.at(extra_w, row[g.dst[extra_w_mask]], 1)
    out["extra_cnt_word"] = extra_w
    return out
