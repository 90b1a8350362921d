"""Drop-in counterparts of the reference's top-level models, HiGraph.HSumGraph (HiGraph.py:34-161) and
HiGraph.HSumDocGraph (:164-244): same constructor `(hps, embed)`, same attribute names and creation order - hence the
SAME state_dict keys (a reference checkpoint loads with strict=True) and, under the same seed, the same initial
weights - and `forward(graph) -> [n sentences, 2]` logits.  `graph` is a hetersumgraph_b200.HeteroBatch built from a
TokenBatch instead of a DGL graph.

Everything between the token ids and the classifier input runs in the library's sm_100a kernels: sentence encoder
(encoder.SentenceEncoder), word embedding lookup, HDSG document-node init, the WSWGAT update loop.  The classifier `wh`
is a stock nn.Linear here (path_model.fused_loss / FusedTrainStep use the library's fused readout + loss instead).
"""
import torch
import torch.nn as nn

from .encoder import SentenceEncoder
from .graph import HeteroBatch
from .modules import WSWGATUpdateLoop


class HSumGraph(SentenceEncoder, WSWGATUpdateLoop):
    """HiGraph.HSumGraph.  hps fields read (train.py:279-309): n_iter, word_emb_dim, sent_max_len, doc_max_timesteps,
    n_feature_size, hidden_size, lstm_hidden_state, lstm_layers, bidirectional, n_head, atten_dropout_prob,
    ffn_inner_hidden_size, ffn_dropout_prob, feat_embed_size."""

    def __init__(self, hps, embed):
        nn.Module.__init__(self)
        self._hps = hps
        self._embed = embed
        self.embed_size = hps.word_emb_dim
        # creation order of HiGraph.py:49-79: _init_sn_param, _TFembed, n_feature_proj, word2sent, sent2word, wh
        self._build_sn_param(embed, hps.word_emb_dim, hps.sent_max_len, hps.doc_max_timesteps, hps.n_feature_size,
                             hps.lstm_hidden_state, hps.lstm_layers, hps.bidirectional, 0.1)
        self._build_tfembed(hps.feat_embed_size, hps.n_iter)
        self._build_n_feature_proj(hps.n_feature_size, hps.hidden_size)
        self._build_layers(hps.word_emb_dim, hps.hidden_size, hps.n_head, hps.atten_dropout_prob,
                           hps.ffn_inner_hidden_size, hps.ffn_dropout_prob, hps.feat_embed_size)
        self.n_feature = hps.hidden_size
        self.wh = nn.Linear(self.n_feature, 2)

    def set_wnfeature(self, graph: HeteroBatch):
        """HiGraph.py:144-152: word rows are already in filter_nodes(unit == 0) order; the tfidfembed edge write is the
        10 x feat_embed table handed to the edge kernels."""
        graph.set_tfidf_embedding(self._TFembed.weight)
        return self._embed(graph.word_wid)

    def set_snfeature_proj(self, graph: HeteroBatch, plan=None):
        """n_feature_proj(set_snfeature(graph)) (HiGraph.py:96,154-161)."""
        return self.encode(plan if plan is not None else graph.encoder_plan)

    def supernode_state(self, graph: HeteroBatch, plan=None):
        """state of the supernodes after the update loop (HiGraph.py:98-106)."""
        word_feature = self.set_wnfeature(graph)
        sent_feature = self.set_snfeature_proj(graph, plan)
        return self.update(graph, word_feature, sent_feature)[1]

    def forward(self, graph: HeteroBatch, plan=None):
        return self.wh(self.supernode_state(graph, plan))                       # HiGraph.py:108

    def loss(self, graph: HeteroBatch, plan=None, n_graphs_global=None):
        """(loss, logits): `outputs = model.forward(G)` followed by the reference's training loss - per-sentence
        cross-entropy, dgl.sum_nodes per graph, mean over graphs (train.py:113-119) - with the classifier, the loss and
        their backward in the library's head kernels (hsg_head_fwd/bwd) instead of ~20 stock launches.  Same values as
        path_model.graph_loss(graph, self(graph), graph.labels); `logits` is returned for extraction only.
        n_graphs_global: the divisor of the mean when the batch is one shard of a data-parallel step."""
        from .functional import SentenceLossFn
        state = self.supernode_state(graph, plan)
        n = n_graphs_global if n_graphs_global is not None else graph.n_graphs
        targets = None
        if self.fuse_grad_accumulation and torch.is_grad_enabled():
            targets = (self.wh.weight.grad, self.wh.bias.grad)
        return SentenceLossFn.apply(graph, n, targets, state, self.wh.weight, self.wh.bias, graph.labels)


class HSumDocGraph(HSumGraph):
    """HiGraph.HSumDocGraph (document nodes, word-document edges)."""

    def __init__(self, hps, embed):
        super().__init__(hps, embed)
        self.dn_feature_proj = nn.Linear(hps.hidden_size, hps.hidden_size, bias=False)     # HiGraph.py:173
        self.wh = nn.Linear(self.n_feature * 2, 2)

    def supernode_state(self, graph: HeteroBatch, plan=None):
        from .functional import DocInitFn
        word_feature = self.set_wnfeature(graph)
        sent_feature = self.set_snfeature_proj(graph, plan)
        super_feature = DocInitFn.apply(graph, sent_feature, self.dn_feature_proj.weight)   # HiGraph.py:196-203,231-244
        return self.update(graph, word_feature, super_feature)[1]                           # :205-214

    def forward(self, graph: HeteroBatch, plan=None):
        state = self.supernode_state(graph, plan)
        s_state = torch.cat([state[graph.sentence_rows()], state[graph.sent_doc_row]], dim=-1)   # :216-228
        return self.wh(s_state)
