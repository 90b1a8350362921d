"""Generate tests/golden/model_*.npz: the UNMODIFIED reference models HiGraph.HSumGraph / HSumDocGraph run end to end
(`forward(graph)`, HiGraph.py:82-110 / :175-228) on the DGL-0.4 shim, default hyper-parameters (train.py:279-309).

    python tests/golden/make_golden_model.py        # build container only (needs /root/reference)

Graphs come from the reference's own ExampleSet / MultiExampleSet.CreateGraph; parameters from
oracle.fixtures.seeded_state_dict (the fixture stores only the seed); dropout off (.eval(), p = 0).  Stored: the
examples, the logits, a cotangent, and the gradients of every trainable parameter (large ones sub-sampled).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the shim, puts /root/reference on sys.path)

import HiGraph  # noqa: E402  the reference

from hetersumgraph_b200 import synthetic as syn  # noqa: E402
from oracle import fixtures as fx  # noqa: E402
from oracle import graph_builder_ref as gb  # noqa: E402

STRIDE = 53
VOCAB = 50000


def hps_default(n_iter):
    return types.SimpleNamespace(n_iter=n_iter, word_emb_dim=300, sent_max_len=100, doc_max_timesteps=50,
                                 n_feature_size=128, hidden_size=64, lstm_hidden_state=128, lstm_layers=2,
                                 bidirectional=True, n_head=8, atten_dropout_prob=0.0, ffn_inner_hidden_size=512,
                                 ffn_dropout_prob=0.0, feat_embed_size=50, cuda=False)


def run(name, hdsg, n_iter, seed):
    exs = syn.make_examples(3, "multinews" if hdsg else "cnndm", seed=seed, hdsg=hdsg)
    if not hdsg:
        for e in exs:                                   # keep the run short: <= 7 sentences per document
            n = min(e.n_sent, 7)
            e.sents, e.w2s, e.labels = e.sents[:n], e.w2s[:n], e.labels[:n]
    filt = set(syn.filter_ids().tolist())
    graphs = [(mg.ref_graph_hdsg if hdsg else mg.ref_graph_hsg)(e, filt) for e in exs]
    order = gb.stable_desc_order([e.n_sent for e in exs]).tolist()
    BG = mg.shim.batch([graphs[i] for i in order])
    embed = torch.nn.Embedding(VOCAB, 300, padding_idx=0)
    model = (HiGraph.HSumDocGraph if hdsg else HiGraph.HSumGraph)(hps_default(n_iter), embed).eval()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = fx.seeded_state_dict(shapes, seed, keep=fx.FROZEN_MODEL_KEYS)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert sorted(missing) == sorted(fx.FROZEN_MODEL_KEYS) and not unexpected
    model._embed.weight.requires_grad_(False)          # train.py:340-342 default
    logits = model(BG)
    cot = torch.randn(logits.shape, generator=torch.Generator().manual_seed(seed + 1))
    (logits * cot).sum().backward()
    out = {"seed": np.int64(seed), "n_iter": np.int64(n_iter), "hdsg": np.int64(hdsg), "stride": np.int64(STRIDE),
           "order": np.asarray(order, np.int64), "logits": logits.detach().numpy(), "cot": cot.numpy()}
    out.update(fx.examples_to_arrays(exs, "ex"))

    def sub(g):
        return g.numpy() if g.numel() <= 2048 else g.flatten()[::STRIDE].numpy()

    for k, p in model.named_parameters():
        if p.requires_grad:
            out["gp:" + k] = sub(p.grad if p.grad is not None else torch.zeros_like(p))
    # the SAME reference code evaluated in float64: tells how far the reference's own fp32 run is from the exact value
    # of every gradient (some are sums that cancel to ~1e-7 of their terms - e.g. feat_fc.bias of sent2word, whose
    # gradient is a softmax-shift direction - and carry no significant digit in fp32)
    graphs64 = [(mg.ref_graph_hdsg if hdsg else mg.ref_graph_hsg)(e, filt) for e in exs]
    BG64 = mg.shim.batch([graphs64[i] for i in order])
    model64 = (HiGraph.HSumDocGraph if hdsg else HiGraph.HSumGraph)(hps_default(n_iter),
                                                                    torch.nn.Embedding(VOCAB, 300, padding_idx=0)).eval()
    model64.load_state_dict(sd, strict=False)
    model64 = model64.double()
    model64._embed.weight.requires_grad_(False)
    logits64 = model64(BG64)
    (logits64 * cot.double()).sum().backward()
    out["logits64"] = logits64.detach().numpy()
    for k, p in model64.named_parameters():
        if p.requires_grad:
            out["g64:" + k] = sub(p.grad if p.grad is not None else torch.zeros_like(p))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "sentences", logits.shape[0], "bytes", os.path.getsize(os.path.join(HERE, name)))


if __name__ == "__main__":
    run("model_hsg_default.npz", False, 1, 31)
    run("model_hdsg_default.npz", True, 1, 32)
