"""Vectorised closed-form restatement of the WSWGAT path (TEST INFRASTRUCTURE).

Independent of oracle/wswgat_ref.py: instead of replaying DGL's per-head,
degree-bucketed execution it evaluates the formula of SURVEY.md §8-a directly
on CSC arrays with scatter ops, all heads at once:

    e      = leaky_relu(p_src + q_bin)                 (dst z is DGL's zero fill)
    sh_v   = sum_act exp(e-m) z_u / (sum_act exp(e-m) + x_v exp(-m)),
    m      = max(max_act e, 0 if x_v > 0)              (x_v extra in-edges, e = 0, z = 0)

Used (i) as a cross-check of wswgat_ref (tests), (ii) as the fast oracle for
large inputs, (iii) as the stronger CPU baseline "B-cf-cpu" in bench.py.
Follows GATLayer.py:88-102,127-140, GATStackLayer.py:55-59, GAT.py:56-58,
HiGraph.py:98-106 of the reference.
"""
import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.01
# test hook: when set to a list, every multi_head_cf call appends the smallest |pre-activation| of leaky_relu over its
# edge logits.  leaky_relu has a kink at 0: a logit within rounding distance of 0 takes slope 1 in one arithmetic and
# 0.01 in another, which no tolerance on the OUTPUT can absorb in the gradients; tests use this to know whether the
# batch at hand contains such a logit.
KINK_LOG = None


def pack_layer(params, prefix, n_heads):
    """per-head state_dict keys -> packed (W [F,in], Wf [F,fe], bf [F], a [H,3d])."""
    W = torch.cat([params[prefix + "heads.%d.fc.weight" % k] for k in range(n_heads)], 0)
    Wf = torch.cat([params[prefix + "heads.%d.feat_fc.weight" % k] for k in range(n_heads)], 0)
    key = prefix + "heads.0.feat_fc.bias"
    if key in params:
        bf = torch.cat([params[prefix + "heads.%d.feat_fc.bias" % k] for k in range(n_heads)], 0)
    else:
        bf = torch.zeros(W.shape[0], dtype=W.dtype)
    a = torch.cat([params[prefix + "heads.%d.attn_fc.weight" % k] for k in range(n_heads)], 0)
    return W, Wf, bf, a


def n_heads_of(params, prefix):
    n = 0
    while (prefix + "heads.%d.fc.weight" % n) in params:
        n += 1
    return n


def multi_head_cf(h_src, n_dst, indptr, src, bins, extra_cnt, W, Wf, bf, a, T, attn_mask=None):
    """All heads of one MultiHeadLayer application on a CSC (dst-major) edge list.
    attn_mask: optional [H, N_src, in] multiplier (0 or 1/(1-p)): head k sees its own dropped-out copy of the
    input, `attn_head(g, self.dropout(h))` (GATStackLayer.py:56)."""
    H = a.shape[0]
    d = a.shape[1] // 3
    indptr = torch.as_tensor(indptr, dtype=torch.int64)
    src = torch.as_tensor(src, dtype=torch.int64)
    bins = torch.as_tensor(bins, dtype=torch.int64)
    extra = torch.as_tensor(extra_cnt, dtype=h_src.dtype).reshape(-1, 1)
    deg = indptr[1:] - indptr[:-1]
    dst = torch.repeat_interleave(torch.arange(n_dst), deg)
    if attn_mask is None:
        z = (h_src @ W.t()).reshape(-1, H, d)
    else:
        z = torch.stack([(h_src * attn_mask[k]) @ W[k * d:(k + 1) * d].t() for k in range(H)], 1)
    p = (z * a[:, :d].unsqueeze(0)).sum(-1)                                   # [N_src, H]
    dfeat = (T @ Wf.t() + bf).reshape(-1, H, d)                               # [10, H, d]
    q = (dfeat * a[:, 2 * d:].unsqueeze(0)).sum(-1)                           # [10, H]
    pre_e = p[src] + q[bins]
    if KINK_LOG is not None and pre_e.numel():
        KINK_LOG.append(float(pre_e.detach().abs().min()))
    e = F.leaky_relu(pre_e, LEAKY_SLOPE)                                      # [E, H]
    neg = torch.full((n_dst, H), float("-inf"), dtype=e.dtype)
    m = neg.scatter_reduce(0, dst.reshape(-1, 1).expand(-1, H), e, "amax", include_self=True)
    m = torch.where(extra > 0, torch.clamp(m, min=0.0), m)
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    w = torch.exp(e - m[dst])
    den = torch.zeros(n_dst, H, dtype=e.dtype).index_add(0, dst, w) + extra * torch.exp(-m)
    den_safe = torch.where(den > 0, den, torch.ones_like(den))
    alpha = w / den_safe[dst]
    sh = torch.zeros(n_dst, H, d, dtype=e.dtype).index_add(0, dst, alpha.unsqueeze(-1) * z[src])
    return sh.reshape(n_dst, H * d)


def ffn_cf(x, w1, b1, w2, b2, gamma, beta, mask=None, flips=None, ffn_mask=None):
    """mask: optional bool [N, d_hid] - the ReLU active set to use instead of (pre > 0).  ReLU is the one
    discontinuous function on the path: a unit whose pre-activation is within rounding distance of 0 may take
    either branch depending on the arithmetic; injecting the device path's own active set lets the tests compare
    everything else at full tolerance.  flips (list) receives (#units whose branch differs, #units, max |pre| there)."""
    pre = x @ w1.reshape(w1.shape[0], -1).t() + b1
    if mask is None:
        h = F.relu(pre)
    else:
        if flips is not None:
            diff = (pre.detach() > 0) != mask
            flips.append((int(diff.sum()), diff.numel(), float(pre.detach().abs()[diff].max()) if diff.any() else 0.0))
        h = pre * mask.to(pre.dtype)
    y = h @ w2.reshape(w2.shape[0], -1).t() + b2
    if ffn_mask is not None:                       # nn.Dropout on the FFN output before the residual (GATLayer.py:41-42)
        y = y * ffn_mask
    return F.layer_norm(y + x, (x.shape[-1],), gamma, beta, 1e-5)


def wswgat_cf(csc, w, s, params, prefix, kind, T, mask=None, flips=None, drop=None):
    """drop: optional dict(attn=[H, N_src, in] multiplier, ffn=[N_dst, F] multiplier) - explicit dropout masks."""
    attn_mask = drop.get("attn") if drop else None
    ffn_mask = drop.get("ffn") if drop else None
    H = n_heads_of(params, prefix + "layer.")
    W, Wf, bf, a = pack_layer(params, prefix + "layer.", H)
    if kind == "W2S":
        origin, neighbor = s, w
        sh = multi_head_cf(neighbor, s.shape[0], csc["super_indptr"], csc["super_src"], csc["super_bin"],
                           csc["extra_cnt"], W, Wf, bf, a, T, attn_mask)
    else:
        origin, neighbor = w, s
        sh = multi_head_cf(neighbor, w.shape[0], csc["word_indptr"], csc["word_src"], csc["word_bin"],
                           csc["extra_cnt_word"], W, Wf, bf, a, T, attn_mask)
    h = F.elu(sh) + origin
    return ffn_cf(h, params[prefix + "ffn.w_1.weight"], params[prefix + "ffn.w_1.bias"],
                  params[prefix + "ffn.w_2.weight"], params[prefix + "ffn.w_2.bias"],
                  params[prefix + "ffn.layer_norm.weight"], params[prefix + "ffn.layer_norm.bias"], mask, flips,
                  ffn_mask)


def update_loop_cf(csc, word_feature, super_feature, params, n_iter, masks=None, flips=None, drops=None):
    """masks: optional list of ReLU active sets, one per WSWGAT application in execution order.
    drops: optional list of explicit dropout masks (see wswgat_cf), one per application."""
    T = params["_TFembed.weight"]
    it = iter(masks) if masks is not None else None
    nxt = (lambda: next(it)) if it is not None else (lambda: None)
    itd = iter(drops) if drops is not None else None
    nxd = (lambda: next(itd)) if itd is not None else (lambda: None)
    word_state = word_feature
    sent_state = wswgat_cf(csc, word_feature, super_feature, params, "word2sent.", "W2S", T, nxt(), flips, nxd())
    for _ in range(n_iter):
        word_state = wswgat_cf(csc, word_state, sent_state, params, "sent2word.", "S2W", T, nxt(), flips, nxd())
        sent_state = wswgat_cf(csc, word_state, sent_state, params, "word2sent.", "W2S", T, nxt(), flips, nxd())
    return word_state, sent_state


# ---------------------------------------------------------------------------------------------------------
# S2S layer type (module/GAT.py:38-39,50-52; SGATLayer module/GATLayer.py:49-78) - never instantiated by the
# reference's models, restated for completeness.
#   z = fc(h) on the supernodes only; attention logits are computed on the dtype == 0 edges (:71 - the
#   word<->supernode edges, NOT the sent<->sent ones) from [z_src, z_dst]: for a word->supernode edge z_src is DGL's
#   zero fill, so e_v = leaky_relu(a[d:2d] . z_v) for every word in-edge of v; pull(snode) reduces over ALL in-edges:
#   deg_v word edges (logit e_v, message 0) and the extra edges (sent->sent twice per ordered pair in HSG,
#   dataloader.py:262-263; sent->doc in HDSG, :385) with the never-written logit 0 and message z_src.  Hence
#       sh_v = sum_{extra in-edges j->v} z_j / (deg_v exp(e_v) + x_v)        (0 when v has no extra in-edge).
# ---------------------------------------------------------------------------------------------------------
def extra_edges_of(g, csc):
    """(src supernode row, dst supernode row) of every in-edge of a supernode that is not a word->supernode edge,
    with multiplicity, from the literal graph arrays."""
    import numpy as np
    row = np.full(g.n_nodes, -1, np.int64)
    row[csc["snode_id"]] = np.arange(len(csc["snode_id"]))
    m = (g.unit[g.dst] == 1) & (g.unit[g.src] == 1)
    return row[g.src[m]], row[g.dst[m]]


def s2s_multi_head_cf(h, deg, xsrc, xdst, W, a):
    """h [Ns, in]; deg [Ns] word in-degree; (xsrc, xdst) extra edges; W [H*d, in]; a [H, 2d]."""
    H = a.shape[0]
    d = a.shape[1] // 2
    n = h.shape[0]
    z = (h @ W.t()).reshape(n, H, d)
    t = (z * a[:, d:].unsqueeze(0)).sum(-1)                                  # [Ns, H]
    e = F.leaky_relu(t, LEAKY_SLOPE)
    xsrc = torch.as_tensor(xsrc, dtype=torch.int64)
    xdst = torch.as_tensor(xdst, dtype=torch.int64)
    xcnt = torch.zeros(n, dtype=h.dtype).index_add(0, xdst, torch.ones(len(xdst), dtype=h.dtype))
    A = torch.zeros(n, H, d, dtype=h.dtype).index_add(0, xdst, z[xsrc])
    den = torch.as_tensor(deg, dtype=h.dtype).reshape(-1, 1) * torch.exp(e) + xcnt.reshape(-1, 1)
    den = torch.where(xcnt.reshape(-1, 1) > 0, den, torch.ones_like(den))
    return (A / den.unsqueeze(-1)).reshape(n, H * d)


def s2s_cf(g, csc, s, params, prefix, mask=None, flips=None):
    """WSWGAT(..., "S2S").forward(g, s, s) (GAT.py:45-59)."""
    import numpy as np
    H = n_heads_of(params, prefix + "layer.")
    W = torch.cat([params[prefix + "layer.heads.%d.fc.weight" % k] for k in range(H)], 0)
    a = torch.cat([params[prefix + "layer.heads.%d.attn_fc.weight" % k] for k in range(H)], 0)
    xsrc, xdst = extra_edges_of(g, csc)
    deg = np.diff(csc["super_indptr"])
    sh = s2s_multi_head_cf(s, deg, xsrc, xdst, W, a)
    h = F.elu(sh) + s
    return ffn_cf(h, params[prefix + "ffn.w_1.weight"], params[prefix + "ffn.w_1.bias"],
                  params[prefix + "ffn.w_2.weight"], params[prefix + "ffn.w_2.bias"],
                  params[prefix + "ffn.layer_norm.weight"], params[prefix + "ffn.layer_norm.bias"], mask, flips)


def doc_init_ref(sent_feature, sent_rows, doc_rows, doc_of_sent_row, W, n_super):
    """HSumDocGraph.forward / set_dnfeature (HiGraph.py:196-203, 231-244): supernode init features =
    sentence rows <- sent_feature, document rows <- dn_feature_proj(mean of the document's sentences)."""
    dmap = torch.full((n_super,), -1, dtype=torch.long)
    dmap[doc_rows] = torch.arange(len(doc_rows))
    didx = dmap[doc_of_sent_row]
    sums = torch.zeros(len(doc_rows), sent_feature.shape[1], dtype=sent_feature.dtype).index_add(0, didx, sent_feature)
    cnt = torch.zeros(len(doc_rows), dtype=sent_feature.dtype).index_add(0, didx, torch.ones(len(didx), dtype=sent_feature.dtype))
    doc_feature = (sums / cnt.unsqueeze(1)) @ W.t()
    out = torch.zeros(n_super, sent_feature.shape[1], dtype=sent_feature.dtype)
    return out.index_copy(0, sent_rows, sent_feature).index_copy(0, doc_rows, doc_feature)
