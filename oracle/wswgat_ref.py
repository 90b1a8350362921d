"""CPU restatement of the reference's WSWGAT message-passing path (TEST INFRASTRUCTURE).

This is the ORACLE the sm_100a kernels are judged against, and the "port" CPU
baseline bench.py times.  It is never imported by hetersumgraph_b200/.

It restates, on flat edge arrays and per head exactly as DGL 0.4 executes it,
what the following reference code computes (paths relative to /root/reference):

  * gat_head            <- WSGATLayer / SWGATLayer / SGATLayer   module/GATLayer.py:49-152
  * multi_head          <- MultiHeadLayer / MultiHeadSGATLayer   module/GATStackLayer.py:27-63
  * ffn                 <- PositionwiseFeedForward               module/GATLayer.py:25-44
  * wswgat              <- WSWGAT.forward                        module/GAT.py:45-59
  * update_loop         <- HSumGraph.forward / HSumDocGraph.forward update loop
                                                                 HiGraph.py:98-106, 205-214
  * tfidf_embed         <- HSumGraph.set_wnfeature               HiGraph.py:144-152

DGL is a third-party dependency that is absent from /root/reference (README.md:15
pins "dgl 0.4", no lockfile).  The DGL-0.4 behaviour reproduced here (zero-filled
frames for rows never written, `pull` over ALL in-edges with degree bucketing,
in-degree-0 destinations skipped) is listed rule by rule in SURVEY.md Appendix A.

Pinning: the reference has no tests or golden vectors of its own.
This restatement is pinned against the reference's OWN modules executed verbatim
on oracle/dgl04_shim.py in the build container: tests/golden/make_golden.py wrote
tests/golden/*.npz from those runs, tests/test_oracle_pinning.py re-checks them
(and re-runs the live reference when /root/reference is present).

Parameters are addressed by the reference's state_dict keys
(e.g. "layer.heads.3.attn_fc.weight", "ffn.w_1.weight").
"""
import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.01  # F.leaky_relu default, GATLayer.py:58,92,131


def tfidf_embed(g, tfembed_weight):
    """HiGraph.py:146,150-151 - edata['tfidfembed'] on dtype==0 edges, zero-fill elsewhere."""
    etype = torch.as_tensor(g.etype)
    tffrac = torch.as_tensor(g.tffrac)
    ids = torch.nonzero(etype == 0).reshape(-1)
    col = torch.zeros(len(etype), tfembed_weight.shape[1], dtype=tfembed_weight.dtype)
    return col.index_copy(0, ids, tfembed_weight[tffrac[ids]])


def _pull_degree_bucketed(n_all, dst_all, pull_nodes, z_src_edges, e_edges, out_dim):
    """DGL-0.4 `pull` with the reference's message/reduce UDFs (GATLayer.py:94-102).

    dst_all: destination of every edge; z_src_edges/e_edges: per-edge message
    fields for ALL edges.  Returns the full 'sh' column [n_all, out_dim] with
    initializer (zero) rows for nodes that were not reduced.
    """
    is_pull = torch.zeros(n_all, dtype=torch.bool)
    is_pull[pull_nodes] = True
    eids = torch.nonzero(is_pull[dst_all]).reshape(-1)
    sh = torch.zeros(n_all, out_dim, dtype=z_src_edges.dtype)
    if len(eids) == 0:
        return sh
    d = dst_all[eids]
    order = torch.argsort(d, stable=True)
    d_sorted = d[order]
    nodes, counts = torch.unique_consecutive(d_sorted, return_counts=True)
    starts = torch.cumsum(counts, 0) - counts
    out_nodes, out_vals = [], []
    for deg in torch.unique(counts).tolist():
        sel = torch.nonzero(counts == deg).reshape(-1)
        pos = (starts[sel].reshape(-1, 1) + torch.arange(deg).reshape(1, -1)).reshape(-1)
        rows = eids[order[pos]]
        mz = z_src_edges[rows].reshape(len(sel), deg, out_dim)      # mailbox['z']
        me = e_edges[rows].reshape(len(sel), deg, 1)                # mailbox['e']
        alpha = F.softmax(me, dim=1)                                # GATLayer.py:100
        out_vals.append(torch.sum(alpha * mz, dim=1))               # GATLayer.py:101
        out_nodes.append(nodes[sel])
    return sh.index_copy(0, torch.cat(out_nodes), torch.cat(out_vals, 0))


def gat_head(g, h, params, prefix, kind, tfidfembed):
    """One attention head.  kind in {"W2S","S2W","S2S"}.

    W2S: GATLayer.py:104-116, S2W: :142-152, S2S: :69-78.
    """
    unit = torch.as_tensor(g.unit)
    src = torch.as_tensor(g.src)
    dst = torch.as_tensor(g.dst)
    etype = torch.as_tensor(g.etype)
    n_all = len(unit)
    wnode = torch.nonzero(unit == 0).reshape(-1)
    snode = torch.nonzero(unit == 1).reshape(-1)
    fc_w = params[prefix + "fc.weight"]
    attn_w = params[prefix + "attn_fc.weight"]
    out_dim = fc_w.shape[0]
    z = h @ fc_w.t()
    if kind == "W2S":
        z_nodes, pull_nodes = wnode, snode
        act = torch.nonzero((unit[src] == 0) & (unit[dst] == 1)).reshape(-1)
    elif kind == "S2W":
        z_nodes, pull_nodes = snode, wnode
        act = torch.nonzero((unit[src] == 1) & (unit[dst] == 0)).reshape(-1)
    else:
        z_nodes, pull_nodes = snode, snode
        act = torch.nonzero(etype == 0).reshape(-1)
    # g.nodes[ids].data['z'] = z : new column, zero-filled elsewhere (GATLayer.py:73,111,147)
    z_all = torch.zeros(n_all, out_dim, dtype=z.dtype).index_copy(0, z_nodes, z)
    # apply_edges(edge_attention, act)
    if kind == "S2S":
        z2 = torch.cat([z_all[src[act]], z_all[dst[act]]], dim=1)
    else:
        feat_w = params[prefix + "feat_fc.weight"]
        dfeat = tfidfembed[act] @ feat_w.t()
        if (prefix + "feat_fc.bias") in params:
            dfeat = dfeat + params[prefix + "feat_fc.bias"]
        z2 = torch.cat([z_all[src[act]], z_all[dst[act]], dfeat], dim=1)
    wa = F.leaky_relu(z2 @ attn_w.t())
    e_all = torch.zeros(len(src), 1, dtype=z.dtype).index_copy(0, act, wa)   # never-written rows stay 0
    # pull(pull_nodes, message_func, reduce_func): message = (src z, e) of EVERY in-edge
    sh = _pull_degree_bucketed(n_all, dst, pull_nodes, z_all[src], e_all, out_dim)
    return sh[pull_nodes]


def multi_head(g, h, params, prefix, kind, tfidfembed, dropout_p=0.0, training=False):
    """GATStackLayer.py:55-63 (merge='cat'); per-head independent dropout of the input."""
    n_heads = 0
    while (prefix + "heads.%d.fc.weight" % n_heads) in params:
        n_heads += 1
    outs = [gat_head(g, F.dropout(h, dropout_p, training), params, prefix + "heads.%d." % k, kind, tfidfembed)
            for k in range(n_heads)]
    return torch.cat(outs, dim=1)


def ffn(x, params, prefix, dropout_p=0.0, training=False):
    """GATLayer.py:35-44 on x [1, N, d_in]; Conv1d(k=1) weights [d_out, d_in, 1]."""
    residual = x
    out = x.transpose(1, 2)
    out = F.conv1d(F.relu(F.conv1d(out, params[prefix + "w_1.weight"], params[prefix + "w_1.bias"])),
                   params[prefix + "w_2.weight"], params[prefix + "w_2.bias"])
    out = out.transpose(1, 2)
    out = F.dropout(out, dropout_p, training)
    d_in = x.shape[-1]
    return F.layer_norm(out + residual, (d_in,), params[prefix + "layer_norm.weight"],
                        params[prefix + "layer_norm.bias"], 1e-5)


def wswgat(g, w, s, params, prefix, kind, tfidfembed, attn_drop=0.0, ffn_drop=0.0, training=False):
    """GAT.py:45-59."""
    if kind == "W2S":
        origin, neighbor = s, w
    else:
        origin, neighbor = w, s
    h = F.elu(multi_head(g, neighbor, params, prefix + "layer.", kind, tfidfembed, attn_drop, training))
    h = h + origin
    return ffn(h.unsqueeze(0), params, prefix + "ffn.", ffn_drop, training).squeeze(0)


def update_loop(g, word_feature, super_feature, params, n_iter, training=False, attn_drop=0.0, ffn_drop=0.0):
    """HiGraph.py:98-106 (HSG) / :205-214 (HDSG; super_feature rows = sentence+doc nodes).

    params uses the model-level keys ("word2sent.…", "sent2word.…", "_TFembed.weight").
    Returns (word_state, sent_state).
    """
    te = tfidf_embed(g, params["_TFembed.weight"])
    word_state = word_feature
    sent_state = wswgat(g, word_feature, super_feature, params, "word2sent.", "W2S", te, attn_drop, ffn_drop, training)
    for _ in range(n_iter):
        word_state = wswgat(g, word_state, sent_state, params, "sent2word.", "S2W", te, attn_drop, ffn_drop, training)
        sent_state = wswgat(g, word_state, sent_state, params, "word2sent.", "W2S", te, attn_drop, ffn_drop, training)
    return word_state, sent_state


def graph_loss(g, logits, labels):
    """train.py:114-119 - CE per sentence node, summed per graph, mean over graphs."""
    ndtype = torch.as_tensor(g.ndtype)
    sent = torch.nonzero(ndtype == 1).reshape(-1)
    seg_all = torch.repeat_interleave(torch.arange(len(g.batch_num_nodes)), torch.as_tensor(g.batch_num_nodes))
    ce = F.cross_entropy(logits, labels, reduction="none")
    per_graph = torch.zeros(len(g.batch_num_nodes), dtype=logits.dtype).index_add(0, seg_all[sent], ce)
    return per_graph.mean()


def topm_indices(g, logits, m):
    """Tester.py:105-131 - per graph torch.topk on the raw class-1 logit."""
    ndtype = torch.as_tensor(g.ndtype)
    seg_all = torch.repeat_interleave(torch.arange(len(g.batch_num_nodes)), torch.as_tensor(g.batch_num_nodes))
    seg = seg_all[ndtype == 1]
    out = []
    for b in range(len(g.batch_num_nodes)):
        p = logits[seg == b][:, 1]
        out.append(torch.topk(p, min(m, len(p)))[1].tolist())
    return out
