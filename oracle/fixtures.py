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
