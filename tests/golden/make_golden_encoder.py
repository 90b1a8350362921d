"""Generate tests/golden/encoder_*.npz by running the UNMODIFIED reference sentence encoder on the DGL-0.4 shim.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_encoder.py

Executed verbatim: HiGraph.HSumGraph.set_snfeature (HiGraph.py:154-161) -> _sent_cnn_feature (:127-133, which calls
module.Encoder.sentEncoder.forward, Encoder.py:56-76), get_snode_feat (:247-255), _sent_lstm_feature (:135-142), and
n_feature_proj (:96), on graphs built by the reference's own ExampleSet.CreateGraph (dataloader.py:222-268) and batched
by the shim's dgl.batch.  The LSTM runs with its inter-layer dropout switched off (`.eval()`), everything else is
dropout-free.  Parameters are drawn by oracle.fixtures.seeded_encoder_params, so the fixture stores only the seed;
gradients of the large tensors are stored sub-sampled (every STRIDE-th element).
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

STRIDE = 37


def run(name, n_sent_per_graph, vocab, emb, L, doc_max, n_feature, hidden, lstm_hidden, seed, zero_pad_row):
    tokens, ptr = fx.encoder_tokens(n_sent_per_graph, L, vocab, seed)
    graphs = []
    for gi in range(len(n_sent_per_graph)):
        sents = tokens[ptr[gi]:ptr[gi + 1]]
        e = syn.DocExample(sents=sents, w2s=[{} for _ in range(len(sents))], labels=np.zeros(len(sents), np.int64))
        graphs.append(mg.ref_graph_hsg(e, set()))
    BG = mg.shim.batch(graphs)                 # already in batch order (#sentences descending)
    hps = types.SimpleNamespace(n_iter=1, word_emb_dim=emb, sent_max_len=L, doc_max_timesteps=doc_max,
                                n_feature_size=n_feature, hidden_size=hidden, lstm_hidden_state=lstm_hidden,
                                lstm_layers=2, bidirectional=True, n_head=2, atten_dropout_prob=0.0,
                                ffn_inner_hidden_size=16, ffn_dropout_prob=0.0, feat_embed_size=6, cuda=False)
    embed = torch.nn.Embedding(vocab, emb, padding_idx=0)
    model = HiGraph.HSumGraph(hps, embed).eval()
    shapes = fx.encoder_param_shapes(vocab, emb, L, doc_max, n_feature, hidden, lstm_hidden)
    params = fx.seeded_encoder_params(shapes, seed, zero_pad_row)
    sd = model.state_dict()
    for k, v in params.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
        if k.endswith("position_embedding.weight") or k == "sent_pos_embed.weight":
            assert torch.equal(sd[k], v), "sinusoid table restatement differs from the reference's: " + k
    with torch.no_grad():
        model._embed.weight.copy_(params["ngram_enc.embed.weight"])
        for k, p in model.named_parameters():
            if k in params and k != "ngram_enc.embed.weight":
                p.copy_(params[k])
    model._embed.weight.requires_grad_(False)          # train.py:340-342 default (frozen embedding)
    sent_feature = model.n_feature_proj(model.set_snfeature(BG))            # HiGraph.py:96
    ngram = BG.ndata["sent_embedding"][BG.filter_nodes(lambda n: n.data["dtype"] == 1)]
    cot = torch.randn(sent_feature.shape, generator=torch.Generator().manual_seed(seed + 77))
    (sent_feature * cot).sum().backward()
    out = {"tokens": tokens, "graph_sent_ptr": ptr, "seed": np.int64(seed), "zero_pad_row": np.int64(zero_pad_row),
           "dims": np.asarray([vocab, emb, L, doc_max, n_feature, hidden, lstm_hidden], np.int64),
           "stride": np.int64(STRIDE), "cot": cot.numpy(), "sent_feature": sent_feature.detach().numpy(),
           "ngram": ngram.detach().numpy()}
    for k, p in model.named_parameters():
        if k in params and p.requires_grad:
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            out["gp:" + k] = g.numpy() if g.numel() <= 8000 else g.flatten()[::STRIDE].numpy()
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "S", tokens.shape[0], "bytes", os.path.getsize(os.path.join(HERE, name)))


def main():
    # the n-gram feature is always 50 * 6 = 300 wide and is added to a word_emb_dim-wide position embedding
    # (HiGraph.py:131-132), so word_emb_dim = 300 is the only width the reference runs with.
    # small everything else, non-zero PAD embedding row (a loaded pretrained table need not keep it zero), ties in
    # #sentences, a one-sentence graph
    run("encoder_small.npz", [6, 6, 3, 1], vocab=120, emb=300, L=24, doc_max=10, n_feature=12, hidden=8, lstm_hidden=8,
        seed=5, zero_pad_row=False)
    # default dims (train.py:279-309): emb 300, sent_max_len 100, doc_max_timesteps 50, n_feature 128, hidden 64, lstm 128
    run("encoder_default.npz", [7, 5, 2], vocab=400, emb=300, L=100, doc_max=50, n_feature=128, hidden=64,
        lstm_hidden=128, seed=6, zero_pad_row=True)


if __name__ == "__main__":
    main()
