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
import os

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
        word_feature = self._embed(g.word_wid)              # F.embedding takes the int32 ids directly
        if not self.hdsg:
            super_feature = sent_feature
        else:
            # doc init = mean of its sentences' init features (set_dnfeature, HiGraph.py:231-244), projected, and the
            # sentence / document rows interleaved per graph: hsg_doc_mean + hsg_gemm_nt + hsg_super_assemble
            from .functional import DocInitFn
            super_feature = DocInitFn.apply(g, sent_feature, self.dn_feature_proj.weight)
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


class _Ctx:
    """Minimal stand-in for the autograd context of the Function classes (FusedTrainStep calls their static
    forward / backward directly)."""

    def __init__(self, needs_input_grad):
        self.needs_input_grad = needs_input_grad
        self.saved_tensors = ()

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors

    def mark_non_differentiable(self, *tensors):
        pass

    def set_materialize_grads(self, value):
        pass


class FusedTrainStep:
    """loss, logits, d_sent_feature = step(batch, sent_feature): forward AND backward of the HSG path from the
    encoder output to the loss, with every parameter gradient accumulated straight into the existing `.grad` buffers
    (a dist.FlatGradArena) - the same C entry points as `fused_loss(...); loss.backward()` (embedding gather,
    hsg_update_loop_fwd, hsg_head_fwd, hsg_head_bwd, hsg_update_loop_bwd), driven without the autograd engine,
    whose CUDA worker-thread hand-off costs about as much host time per step as all the kernel launches together
    at batch 32.  d_sent_feature is what the sentence encoder's backward consumes (HiGraph.py:96).

    HSG and HDSG (document-node init through DocInitFn's kernels, dn_feature_proj gradient added into its .grad) with
    a frozen embedding (the reference default, train.py:340-342); a trainable embedding uses the autograd path.
    Gradient parity with the autograd path is tested."""

    def __init__(self, model: "HSGPath", n_graphs_global=None):
        if model._embed.weight.requires_grad:
            raise NotImplementedError("FusedTrainStep assumes the frozen word embedding of the reference default")
        self.model, self.n_graphs_global = model, n_graphs_global
        # measured (gpurun r02x): as its own root of the step graph the gather starts ~9 us AFTER the attention prep, the
        # first projection begins at the same time either way (0.6031 ms per step on and off) - off by default
        self.gather_overlap = os.environ.get("HSG_GATHER_OVERLAP", "0") == "1"
        self._gstream = self._gevent = None
        self.fused_head = os.environ.get("HSG_FUSED_HEAD", "1") != "0"

    def __call__(self, g: HeteroBatch, sent_feature: torch.Tensor, hooks=None):
        """hooks: optional dict of callables invoked between the enqueue phases of the step ("after_forward": every
        forward kernel of the update loop is enqueued; "after_head": loss forward + backward enqueued) - the step graph
        forks its build-of-the-next-batch branch there instead of at the very start."""
        from .functional import SentenceLossFn, UpdateLoopFn
        hooks = hooks or {}
        m = self.model
        loop = m.loop
        prev = loop.fuse_grad_accumulation
        loop.fuse_grad_accumulation = True
        try:
            cfg, tensors = loop.loop_call(g)                       # needs grad mode on to pick up the .grad targets
        finally:
            loop.fuse_grad_accumulation = prev
        if m.wh.weight.grad is None or m.wh.bias.grad is None:
            raise RuntimeError("FusedTrainStep needs .grad buffers on every parameter (dist.FlatGradArena)")
        n = self.n_graphs_global if self.n_graphs_global is not None else g.n_graphs
        with torch.no_grad():
            from .functional import embed_gather
            # own kernel (no stock ATen launch on the step), on a forked stream: the update loop's parameter-only
            # attention prep runs next to it and application 0 waits for the event (hsg_loop_args.input_ready)
            if self.gather_overlap:
                cur = torch.cuda.current_stream(g.word_wid.device)
                if self._gstream is None:
                    self._gstream = torch.cuda.Stream(device=g.word_wid.device)
                    self._gevent = torch.cuda.Event()
                word_feature = torch.empty(g.word_wid.shape[0], m._embed.weight.shape[1], dtype=torch.float32,
                                           device=g.word_wid.device)
                self._gstream.wait_stream(cur)
                with torch.cuda.stream(self._gstream):
                    embed_gather(g.word_wid, m._embed.weight, out=word_feature)
                self._gevent.record(self._gstream)
                cfg = dict(cfg, input_ready=self._gevent)
            else:
                word_feature = embed_gather(g.word_wid, m._embed.weight)
            super_feature = sent_feature
            if m.hdsg:                                   # HiGraph.py:196-203,231-244
                from .functional import DocInitFn
                if m.dn_feature_proj.weight.grad is None:
                    raise RuntimeError("FusedTrainStep needs .grad buffers on every parameter (dist.FlatGradArena)")
                c0 = _Ctx([False, True, True])
                super_feature = DocInitFn.forward(c0, g, sent_feature, m.dn_feature_proj.weight)
            c1 = _Ctx([False, False, False, True] + [False] * len(tensors))
            _, super_state = UpdateLoopFn.forward(c1, g, cfg, word_feature, super_feature, *tensors)
            if "after_forward" in hooks:
                hooks["after_forward"]()
            if self.fused_head:                           # loss forward + backward in one launch (bit-identical)
                from .functional import head_fwd_bwd
                loss, logits, d_state = head_fwd_bwd(g, n, (m.wh.weight.grad, m.wh.bias.grad), super_state,
                                                     m.wh.weight, m.wh.bias, g.labels)
            else:
                c2 = _Ctx([False] * 7)
                loss, logits = SentenceLossFn.forward(c2, g, n, (m.wh.weight.grad, m.wh.bias.grad), super_state,
                                                      m.wh.weight, m.wh.bias, g.labels)
                d_state = SentenceLossFn.backward(c2, None, None)[3]
            if "after_head" in hooks:
                hooks["after_head"]()
            d_sent_feature = UpdateLoopFn.backward(c1, None, d_state)[3]
            if m.hdsg:
                _, d_sent_feature, dW = DocInitFn.backward(c0, d_sent_feature)
                m.dn_feature_proj.weight.grad.add_(dW)
        return loss, logits, d_sent_feature


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
