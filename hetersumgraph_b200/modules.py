"""Drop-in replacements for the reference's WSWGAT modules.

Same constructors, forward signatures and state_dict keys as
  module/GAT.py:30-59            WSWGAT
  module/GATStackLayer.py:46-63  MultiHeadLayer
  module/GATLayer.py:25-44       PositionwiseFeedForward
  module/GATLayer.py:81-152      WSGATLayer / SWGATLayer (used as the `layer=` selector)
but the `g` argument is a hetersumgraph_b200.graph.HeteroBatch instead of a DGL
graph and every forward/backward runs in the sm_100a kernels of libhsg_b200.so.

Parameters are stored packed over heads (W [H*d, in], Wf [H*d, fe], bf [H*d],
a [H, 3d]); state_dict()/load_state_dict() speak the reference's per-head keys
("layer.heads.3.fc.weight", ...), so the released checkpoints load unchanged.
"""
import math

import torch
import torch.nn as nn

from .functional import (AttnPrepFn, FFNDropFn, FFNFn, MultiHeadDropFn, MultiHeadFn, S2SFn, UpdateLoopFn,
                         WSWGATCoreFn)


class WSGATLayer:
    """Selector for word->supernode heads (reference: module/GATLayer.py:81-116; feat_fc has NO bias)."""
    kind = "W2S"
    feat_bias = False


class SWGATLayer:
    """Selector for supernode->word heads (reference: module/GATLayer.py:120-152; feat_fc HAS a bias)."""
    kind = "S2W"
    feat_bias = True


def _check_dropout(module, p, what):
    if module.training and p > 0.0:
        raise NotImplementedError(
            "hetersumgraph_b200: %s dropout p=%g in training mode is implemented inside WSWGAT / WSWGATUpdateLoop "
            "(hsg_update_loop_fwd), not for this sub-layer on its own; call it through WSWGAT, construct it with "
            "p=0 or use .eval()" % (what, p))


def _next_dropout_seed():
    """A fresh 63-bit seed from torch's CPU generator (so torch.manual_seed makes runs reproducible); drawn on the
    host, no device synchronisation."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


# test hook: when set to a list, every dropout-enabled forward appends its (seed, n_apps, start_kind)
DROPOUT_SEED_LOG = None


def _packed_params(mod):
    lay = mod.layer
    return (lay.fc_weight, lay.feat_fc_weight, lay.feat_fc_bias, lay.attn_fc_weight) + mod.ffn.packed()


def _grad_buffers(mod):
    """.grad of the ten packed parameters, viewed in the packed shapes (Conv1d weights are [out, in, 1])."""
    lay, ffn = mod.layer, mod.ffn
    leaves = (lay.fc_weight, lay.feat_fc_weight, lay.feat_fc_bias, lay.attn_fc_weight, ffn.w_1.weight, ffn.w_1.bias,
              ffn.w_2.weight, ffn.w_2.bias, ffn.layer_norm.weight, ffn.layer_norm.bias)
    out = []
    for p in leaves:
        if p is None:
            out.append(None)
            continue
        if p.grad is None or not p.grad.is_contiguous():
            raise RuntimeError("fuse_grad_accumulation needs a contiguous .grad buffer on every parameter "
                               "(see hetersumgraph_b200.dist.FlatGradArena)")
        out.append(p.grad.view(p.shape[0], -1) if p.dim() == 3 else p.grad)
    return out


def _dropout_cfg(mods, fixed_seed=None):
    """(attn_p, ffn_p, seed) of a chain over `mods` (training mode only; the C loop takes one pair of rates).
    fixed_seed: callable returning the seed to use instead of drawing a fresh one (device-step-keyed masks)."""
    attn = {m.layer.dropout.p if m.layer.training else 0.0 for m in mods}
    ffn = {m.ffn.dropout.p if m.ffn.training else 0.0 for m in mods}
    if len(attn) > 1 or len(ffn) > 1:
        raise NotImplementedError("word2sent and sent2word must share their dropout rates (the reference builds both "
                                  "from hps.atten_dropout_prob / hps.ffn_dropout_prob, HiGraph.py:57-76)")
    attn_p, ffn_p = attn.pop(), ffn.pop()
    seed = 0
    if attn_p > 0.0 or ffn_p > 0.0:
        seed = fixed_seed() if fixed_seed is not None else _next_dropout_seed()
    return attn_p, ffn_p, seed


class PositionwiseFeedForward(nn.Module):
    """A two-feed-forward-layer module (reference: module/GATLayer.py:25-44)."""

    def __init__(self, d_in, d_hid, dropout=0.1):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, 1)   # parameter containers only; the math runs in hsg_gemm_nt
        self.w_2 = nn.Conv1d(d_hid, d_in, 1)
        self.layer_norm = nn.LayerNorm(d_in)
        self.dropout = nn.Dropout(dropout)
        self.d_in, self.d_hid = d_in, d_hid

    def packed(self):
        return (self.w_1.weight.view(self.d_hid, self.d_in), self.w_1.bias,
                self.w_2.weight.view(self.d_in, self.d_hid), self.w_2.bias,
                self.layer_norm.weight, self.layer_norm.bias)

    def forward(self, x):
        squeeze = x.dim() == 3
        x2 = x.reshape(-1, x.shape[-1])
        if self.training and self.dropout.p > 0.0:     # on its own (inside WSWGAT the kernels apply the mask)
            out = FFNDropFn.apply(x2, *self.packed(), float(self.dropout.p), _next_dropout_seed())
        else:
            out = FFNFn.apply(x2, *self.packed())
        return out.reshape(x.shape) if squeeze else out


class MultiHeadLayer(nn.Module):
    """num_heads attention heads, outputs concatenated (reference: module/GATStackLayer.py:46-63)."""

    def __init__(self, in_dim, out_dim, num_heads, attn_drop_out, feat_embed_size, layer, merge='cat'):
        super().__init__()
        if merge != 'cat':
            raise NotImplementedError("merge != 'cat' is a scalar-mean bug in the reference (GATStackLayer.py:62) "
                                      "and is never used")
        self.kind = layer.kind
        self.in_dim, self.out_dim, self.num_heads, self.feat_embed_size = in_dim, out_dim, num_heads, feat_embed_size
        F = out_dim * num_heads
        self.fc_weight = nn.Parameter(torch.empty(F, in_dim))
        self.feat_fc_weight = nn.Parameter(torch.empty(F, feat_embed_size))
        self.feat_fc_bias = nn.Parameter(torch.empty(F)) if layer.feat_bias else None
        self.attn_fc_weight = nn.Parameter(torch.empty(num_heads, 3 * out_dim))
        self.merge = merge
        self.dropout = nn.Dropout(attn_drop_out)
        self.reset_parameters()

    def reset_parameters(self):
        # nn.Linear default init per head (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)))
        for w, fan_in in ((self.fc_weight, self.in_dim), (self.feat_fc_weight, self.feat_embed_size),
                          (self.attn_fc_weight, 3 * self.out_dim)):
            bound = 1.0 / math.sqrt(fan_in)
            nn.init.uniform_(w, -bound, bound)
        if self.feat_fc_bias is not None:
            bound = 1.0 / math.sqrt(self.feat_embed_size)
            nn.init.uniform_(self.feat_fc_bias, -bound, bound)

    # ---- reference state_dict layout: heads.{k}.fc.weight / feat_fc.weight / feat_fc.bias / attn_fc.weight
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        d = self.out_dim
        for k in range(self.num_heads):
            sl = slice(k * d, (k + 1) * d)
            items = [("fc.weight", self.fc_weight[sl]), ("feat_fc.weight", self.feat_fc_weight[sl])]
            if self.feat_fc_bias is not None:
                items.append(("feat_fc.bias", self.feat_fc_bias[sl]))
            items.append(("attn_fc.weight", self.attn_fc_weight[k:k + 1]))
            for name, t in items:
                destination["%sheads.%d.%s" % (prefix, k, name)] = t if keep_vars else t.detach()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        d = self.out_dim
        names = ["fc.weight", "feat_fc.weight", "attn_fc.weight"] + (["feat_fc.bias"] if self.feat_fc_bias is not None else [])
        target = {"fc.weight": self.fc_weight, "feat_fc.weight": self.feat_fc_weight,
                  "feat_fc.bias": self.feat_fc_bias, "attn_fc.weight": self.attn_fc_weight}
        expected = set()
        with torch.no_grad():
            for k in range(self.num_heads):
                for name in names:
                    key = "%sheads.%d.%s" % (prefix, k, name)
                    expected.add(key)
                    if key not in state_dict:
                        missing_keys.append(key)
                        continue
                    src = state_dict[key]
                    dst = target[name][k:k + 1] if name == "attn_fc.weight" else target[name][k * d:(k + 1) * d]
                    if src.shape != dst.shape:
                        error_msgs.append("size mismatch for %s: %s vs %s" % (key, tuple(src.shape), tuple(dst.shape)))
                        continue
                    dst.copy_(src)
        if strict:
            for key in state_dict:
                if key.startswith(prefix + "heads.") and key not in expected:
                    unexpected_keys.append(key)

    def forward(self, g, h):
        if g.tfidfembed_weight is None:
            raise RuntimeError("HeteroBatch has no TF-IDF embedding table: call g.set_tfidf_embedding(_TFembed.weight) "
                               "(counterpart of HSumGraph.set_wnfeature, HiGraph.py:150-151)")
        if self.training and self.dropout.p > 0.0:     # on its own (inside WSWGAT the kernels apply the masks)
            return MultiHeadDropFn.apply(g, self.kind, self.num_heads, self.out_dim, h, self.fc_weight,
                                         self.feat_fc_weight, self.feat_fc_bias, self.attn_fc_weight,
                                         g.tfidfembed_weight, float(self.dropout.p), _next_dropout_seed())
        return MultiHeadFn.apply(g, self.kind, self.num_heads, self.out_dim, h, self.fc_weight, self.feat_fc_weight,
                                 self.feat_fc_bias, self.attn_fc_weight, g.tfidfembed_weight)


class MultiHeadSGATLayer(nn.Module):
    """num_heads sentence->sentence heads (reference: module/GATStackLayer.py:27-44 over SGATLayer,
    module/GATLayer.py:49-78).  The reference's models never instantiate it (HiGraph.py:57-76)."""

    def __init__(self, in_dim, out_dim, num_heads, attn_drop_out, merge='cat'):
        super().__init__()
        if merge != 'cat':
            raise NotImplementedError("merge != 'cat' is a scalar-mean bug in the reference (GATStackLayer.py:44)")
        self.in_dim, self.out_dim, self.num_heads, self.merge = in_dim, out_dim, num_heads, merge
        self.fc_weight = nn.Parameter(torch.empty(out_dim * num_heads, in_dim))
        self.attn_fc_weight = nn.Parameter(torch.empty(num_heads, 2 * out_dim))
        self.dropout = nn.Dropout(attn_drop_out)
        for w, fan_in in ((self.fc_weight, in_dim), (self.attn_fc_weight, 2 * out_dim)):
            nn.init.uniform_(w, -1.0 / math.sqrt(fan_in), 1.0 / math.sqrt(fan_in))

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        d = self.out_dim
        for k in range(self.num_heads):
            for name, t in (("fc.weight", self.fc_weight[k * d:(k + 1) * d]),
                            ("attn_fc.weight", self.attn_fc_weight[k:k + 1])):
                destination["%sheads.%d.%s" % (prefix, k, name)] = t if keep_vars else t.detach()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        d = self.out_dim
        with torch.no_grad():
            for k in range(self.num_heads):
                for name, dst in (("fc.weight", self.fc_weight[k * d:(k + 1) * d]),
                                  ("attn_fc.weight", self.attn_fc_weight[k:k + 1])):
                    key = "%sheads.%d.%s" % (prefix, k, name)
                    if key not in state_dict:
                        missing_keys.append(key)
                    elif state_dict[key].shape != dst.shape:
                        error_msgs.append("size mismatch for %s" % key)
                    else:
                        dst.copy_(state_dict[key])

    def forward(self, g, h, origin=None):
        _check_dropout(self, self.dropout.p, "attention-input")
        return S2SFn.apply(g, self.num_heads, self.out_dim, h, origin, self.fc_weight, self.attn_fc_weight)


class WSWGAT(nn.Module):
    """reference: module/GAT.py:30-59."""

    def __init__(self, in_dim, out_dim, num_heads, attn_drop_out, ffn_inner_hidden_size, ffn_drop_out,
                 feat_embed_size, layerType):
        super().__init__()
        self.layerType = layerType
        if layerType == "W2S":
            self.layer = MultiHeadLayer(in_dim, int(out_dim / num_heads), num_heads, attn_drop_out, feat_embed_size,
                                        layer=WSGATLayer)
        elif layerType == "S2W":
            self.layer = MultiHeadLayer(in_dim, int(out_dim / num_heads), num_heads, attn_drop_out, feat_embed_size,
                                        layer=SWGATLayer)
        elif layerType == "S2S":
            # never instantiated by HSumGraph / HSumDocGraph (HiGraph.py:57-76); built for completeness
            self.layer = MultiHeadSGATLayer(in_dim, int(out_dim / num_heads), num_heads, attn_drop_out)
        else:
            raise NotImplementedError("GAT Layer has not been implemented!")
        self.ffn = PositionwiseFeedForward(out_dim, ffn_inner_hidden_size, ffn_drop_out)

    def prepare(self, g):
        """(W_aug, q) of this layer for the TF-IDF table carried by `g`: parameters only, reusable across the
        applications of one step (the update loop of HiGraph.py:98-106 re-applies the same modules)."""
        lay = self.layer
        if g.tfidfembed_weight is None:
            raise RuntimeError("HeteroBatch has no TF-IDF embedding table: call g.set_tfidf_embedding(_TFembed.weight)")
        return AttnPrepFn.apply(lay.num_heads, lay.out_dim, lay.fc_weight, lay.feat_fc_weight, lay.feat_fc_bias,
                                lay.attn_fc_weight, g.tfidfembed_weight)

    def forward(self, g, w, s, prepared=None):
        if self.layerType == "S2S":
            assert torch.equal(w, s)                                           # GAT.py:51
            return self.ffn(self.layer(g, s, origin=w))
        if self.layerType == "W2S":
            origin, neighbor = s, w
        else:
            origin, neighbor = w, s
        lay = self.layer
        attn_p, ffn_p, seed = _dropout_cfg([self])
        if attn_p > 0.0 or ffn_p > 0.0:
            # training-mode dropout lives in the whole-loop entry point: one application of this kind
            if g.tfidfembed_weight is None:
                raise RuntimeError("HeteroBatch has no TF-IDF embedding table: call g.set_tfidf_embedding(_TFembed.weight)")
            kind = 0 if self.layerType == "W2S" else 1
            dims = (lay.num_heads, lay.out_dim, self.ffn.d_hid)
            cfg = dict(n_apps=1, start_kind=kind, w2s=dims, s2w=dims, grad_targets=None, attn_p=attn_p, ffn_p=ffn_p,
                       seed=seed)
            if DROPOUT_SEED_LOG is not None:
                DROPOUT_SEED_LOG.append((seed, 1, kind))
            mine = _packed_params(self)
            none10 = (None,) * 10
            pw, ps = (mine, none10) if kind == 0 else (none10, mine)
            ws_, ss_ = UpdateLoopFn.apply(g, cfg, w, s, g.tfidfembed_weight, *pw, *ps)
            return ss_ if kind == 0 else ws_
        W_aug, q = prepared if prepared is not None else self.prepare(g)
        return WSWGATCoreFn.apply(g, self.layerType, lay.num_heads, lay.out_dim, neighbor, origin, W_aug, q,
                                  *self.ffn.packed())


class WSWGATUpdateLoop(nn.Module):
    """The iterative word<->sentence update of HSumGraph.forward / HSumDocGraph.forward
    (HiGraph.py:98-106, 205-214): W2S, then n_iter x (S2W, W2S), weights shared across iterations."""

    def __init__(self, word_emb_dim=300, hidden_size=64, n_head=8, atten_dropout_prob=0.1,
                 ffn_inner_hidden_size=512, ffn_dropout_prob=0.1, feat_embed_size=50, n_iter=1):
        super().__init__()
        self._build_tfembed(feat_embed_size, n_iter)
        self._build_layers(word_emb_dim, hidden_size, n_head, atten_dropout_prob, ffn_inner_hidden_size,
                           ffn_dropout_prob, feat_embed_size)

    def _build_tfembed(self, feat_embed_size, n_iter):
        self._n_iter = n_iter
        self._TFembed = nn.Embedding(10, feat_embed_size)   # box=10 (HiGraph.py:52)

    def _build_layers(self, word_emb_dim, hidden_size, n_head, atten_dropout_prob, ffn_inner_hidden_size,
                      ffn_dropout_prob, feat_embed_size):
        self.word2sent = WSWGAT(word_emb_dim, hidden_size, n_head, atten_dropout_prob, ffn_inner_hidden_size,
                                ffn_dropout_prob, feat_embed_size, "W2S")
        self.sent2word = WSWGAT(hidden_size, word_emb_dim, 6, atten_dropout_prob, ffn_inner_hidden_size,
                                ffn_dropout_prob, feat_embed_size, "S2W")

    # When True, backward ADDS the parameter gradients straight into the existing `.grad` buffers (e.g. the views
    # of a dist.FlatGradArena) from inside the kernels' last stages instead of returning them to autograd, which
    # would launch one add per parameter.  Same arithmetic; requires every parameter to have a `.grad` tensor.
    fuse_grad_accumulation = False
    # Optional int64 device tensor (FusedAdam.device_step_counter()) mixed into the dropout key at run time, so a
    # CUDA-graph replay of the step draws fresh masks every step although its kernel arguments never change.
    seed_dev = None

    def _grad_targets(self):
        """[dT, 10 x word2sent, 10 x sent2word] .grad buffers in packed shapes; the views are rebuilt only when a
        .grad tensor has been replaced."""
        leaves = self.__dict__.get("_gt_leaves")       # parameter objects never change identity: walk the modules once
        if leaves is None:
            leaves = [self._TFembed.weight] + [p for m in (self.word2sent, self.sent2word) for p in m.parameters()]
            self.__dict__["_gt_leaves"] = leaves
        sig = tuple(-1 if p.grad is None else p.grad.data_ptr() for p in leaves)
        ent = self.__dict__.get("_gt_cache")
        if ent is None or ent[0] != sig:
            if self._TFembed.weight.grad is None:
                raise RuntimeError("fuse_grad_accumulation needs a .grad buffer on every parameter")
            targets = [self._TFembed.weight.grad] + _grad_buffers(self.word2sent) + _grad_buffers(self.sent2word)
            ent = (sig, targets)
            self.__dict__["_gt_cache"] = ent
        return ent[1]

    def loop_call(self, graph):
        """(cfg, parameter tensors) of one UpdateLoopFn call on `graph` - shared by forward() and by
        path_model.FusedTrainStep (which drives the same functions without the autograd engine)."""
        graph.set_tfidf_embedding(self._TFembed.weight)
        mods = (self.word2sent, self.sent2word) if self._n_iter > 0 else (self.word2sent,)
        # with a device step counter the host seed is drawn ONCE (a replayed graph cannot take a new kernel argument
        # per step); the per-step variation comes from the counter
        fixed = None
        if self.seed_dev is not None:
            fixed = lambda: self.__dict__.setdefault("_base_seed", _next_dropout_seed())   # noqa: E731
        attn_p, ffn_p, seed = _dropout_cfg(mods, fixed)
        pw, ps = _packed_params(self.word2sent), _packed_params(self.sent2word)
        targets = None
        if self.fuse_grad_accumulation and torch.is_grad_enabled():
            targets = self._grad_targets()
        lw, ls = self.word2sent.layer, self.sent2word.layer
        n_apps = 1 + 2 * self._n_iter
        cfg = dict(n_apps=n_apps, start_kind=0, w2s=(lw.num_heads, lw.out_dim, self.word2sent.ffn.d_hid),
                   s2w=(ls.num_heads, ls.out_dim, self.sent2word.ffn.d_hid), grad_targets=targets, attn_p=attn_p,
                   ffn_p=ffn_p, seed=seed, seed_dev=self.seed_dev, cache=self.__dict__.setdefault("_arg_cache", {}))
        if DROPOUT_SEED_LOG is not None and seed:
            DROPOUT_SEED_LOG.append((seed, n_apps, 0))
        return cfg, (self._TFembed.weight,) + pw + ps

    def forward(self, graph, word_feature, sent_feature):
        return self.update(graph, word_feature, sent_feature)

    def update(self, graph, word_feature, sent_feature):
        cfg, tensors = self.loop_call(graph)
        return UpdateLoopFn.apply(graph, cfg, word_feature, sent_feature, *tensors)

    def forward_per_application(self, graph, word_feature, sent_feature):
        """Same loop through one autograd node per WSWGAT application (the path a caller gets by invoking the
        WSWGAT modules directly, as HiGraph.py:100-106 does); kept for the parity tests."""
        graph.set_tfidf_embedding(self._TFembed.weight)
        p_w2s = self.word2sent.prepare(graph)                 # once per layer and step (shared weights)
        p_s2w = self.sent2word.prepare(graph) if self._n_iter > 0 else None
        word_state = word_feature
        sent_state = self.word2sent(graph, word_feature, sent_feature, prepared=p_w2s)
        for _ in range(self._n_iter):
            word_state = self.sent2word(graph, word_state, sent_state, prepared=p_s2w)
            sent_state = self.word2sent(graph, word_state, sent_state, prepared=p_w2s)
        return word_state, sent_state
