"""HSG / HDSG forward from the point where the WSWGAT path starts.

Mirrors HSumGraph.forward (HiGraph.py:82-110) / HSumDocGraph.forward (:177-228) with the
sentence encoder's output (`sent_feature`, HiGraph.py:96) given as an input: the CNN+BiLSTM
sentence encoder (module/Encoder.py, HiGraph.py:112-161) is out of the hot-path scope
 and stays in stock PyTorch.

  word_feature = embed(word ids)                      set_wnfeature   HiGraph.py:144-152
  word/sent states = WSWGAT update loop               HiGraph.py:98-106          <- sm_100a kernels
  logits = wh(sent_state[sentence rows])              HiGraph.py:108 (HDSG: cat doc state, :216-228)
  loss   = mean_graphs sum_sentences CE               train.py:114-119
  top-m  = per graph topk(logit[:,1], m)              Tester.py:128
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .graph import HeteroBatch
from .modules import WSWGATUpdateLoop


class HSGPath(nn.Module):
    def __init__(self, vocab_size=50000, word_emb_dim=300, hidden_size=64, n_head=8, atten_dropout_prob=0.0,
                 ffn_inner_hidden_size=512, ffn_dropout_prob=0.0, feat_embed_size=50, n_iter=1, hdsg=False,
                 embed=None):
        super().__init__()
        self.hdsg = hdsg
        self._embed = embed if embed is not None else nn.Embedding(vocab_size, word_emb_dim, padding_idx=0)
        self._embed.weight.requires_grad_(False)          # frozen unless --embed_train (train.py:340-342)
        self.loop = WSWGATUpdateLoop(word_emb_dim, hidden_size, n_head, atten_dropout_prob, ffn_inner_hidden_size,
                                     ffn_dropout_prob, feat_embed_size, n_iter)
        self.wh = nn.Linear(hidden_size * (2 if hdsg else 1), 2)
        if hdsg:
            self.dn_feature_proj = nn.Linear(hidden_size, hidden_size, bias=False)   # HiGraph.py:174

    def doc_of_sentence(self, g: HeteroBatch):
        """supernode row of the document of every sentence row (HDSG), from the builder's sent->doc map."""
        return g.sent_doc_row

    def states(self, g: HeteroBatch, sent_feature: torch.Tensor):
        """(word_state, supernode_state) after the update loop."""
        word_feature = self._embed(g.word_wid.long())
        if not self.hdsg:
            super_feature = sent_feature
        else:
            srows, drows = g.sentence_rows(), g.doc_rows()
            d_of_s = g.sent_doc_row                                   # [n_sent] supernode row of each sentence's doc
            # doc init = mean of its sentences' init features (set_dnfeature, HiGraph.py:231-244)
            dmap = torch.full((g.n_super,), -1, dtype=torch.long, device=sent_feature.device)
            dmap[drows] = torch.arange(len(drows), device=sent_feature.device)
            didx = dmap[d_of_s]
            sums = torch.zeros(len(drows), sent_feature.shape[1], device=sent_feature.device).index_add(0, didx, sent_feature)
            cnt = torch.zeros(len(drows), device=sent_feature.device).index_add(0, didx, torch.ones_like(didx, dtype=torch.float32))
            doc_feature = self.dn_feature_proj(sums / cnt.unsqueeze(1))
            super_feature = torch.zeros(g.n_super, sent_feature.shape[1], device=sent_feature.device)
            super_feature = super_feature.index_copy(0, srows, sent_feature).index_copy(0, drows, doc_feature)
        return self.loop(g, word_feature, super_feature)

    def forward(self, g: HeteroBatch, sent_feature: torch.Tensor):
        """sent_feature: [n sentence rows, hidden] (HSG) - the encoder output in batched-graph sentence order."""
        _, sent_state = self.states(g, sent_feature)
        if not self.hdsg:
            return self.wh(sent_state)
        s_state = torch.cat([sent_state[g.sentence_rows()], sent_state[g.sent_doc_row]], dim=-1)   # HiGraph.py:216-228
        return self.wh(s_state)


def fused_loss(model: "HSGPath", g: HeteroBatch, sent_feature: torch.Tensor, n_graphs_global=None,
               fuse_grad_accumulation=False):
    """(loss, logits) with the classifier, the per-graph cross-entropy sums and their mean in one device kernel pair
    (hsg_head_fwd/bwd) instead of ~20 stock launches.  Same value as graph_loss(g, model(g, sf), g.labels, ...)."""
    from .functional import SentenceLossFn
    state = model.states(g, sent_feature)[1]
    targets = None
    if fuse_grad_accumulation and torch.is_grad_enabled():
        targets = (model.wh.weight.grad, model.wh.bias.grad)
    n = n_graphs_global if n_graphs_global is not None else g.n_graphs
    return SentenceLossFn.apply(g, n, targets, state, model.wh.weight, model.wh.bias, g.labels)


def graph_loss(g: HeteroBatch, logits: torch.Tensor, labels: torch.Tensor, n_graphs_global=None):
    """train.py:114-119: CE per sentence node, dgl.sum_nodes per graph, mean over graphs."""
    ce = F.cross_entropy(logits, labels, reduction="sum")
    return ce / float(n_graphs_global if n_graphs_global is not None else g.n_graphs)


def topm_indices(g: HeteroBatch, logits: torch.Tensor, m: int):
    """Tester.py:105-131: per graph torch.topk on the raw class-1 logit; returns a list of index lists."""
    srows = g.sentence_rows()
    seg = g.super_graph[srows].long()
    p = logits[:, 1]
    counts = torch.bincount(seg, minlength=g.n_graphs).tolist()
    out, off = [], 0
    for c in counts:
        out.append(torch.topk(p[off:off + c], min(m, c))[1].tolist() if c else [])
        off += c
    return out
