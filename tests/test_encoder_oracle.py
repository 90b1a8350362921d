"""CPU tests: pin oracle/encoder_ref.py (restatement of module/Encoder.py + HiGraph.py:112-161) against the golden
vectors generated from the UNMODIFIED reference (tests/golden/make_golden_encoder.py) and, when /root/reference is
present (build container), against the live reference classes."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import encoder_ref as er
from oracle import fixtures as fx

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5   # BASELINE.json: fp32 normalised max error


def nerr(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


load_encoder_fixture = lambda name: fx.load_encoder_fixture(GOLD, name)  # noqa: E731
golden_grad = fx.golden_grad


@pytest.mark.parametrize("name", ["encoder_small.npz", "encoder_default.npz"])
def test_encoder_restatement_matches_reference_golden(name):
    z, params = load_encoder_fixture(name)
    frozen = ("ngram_enc.embed.weight", "sent_pos_embed.weight", "ngram_enc.position_embedding.weight")
    p = {k: (v.clone().requires_grad_(True) if k not in frozen else v) for k, v in params.items()}
    sf, ngram = er.sent_feature(z["tokens"], z["graph_sent_ptr"], p)
    assert nerr(ngram, z["ngram"]) <= 2e-6
    assert nerr(sf, z["sent_feature"]) <= 2e-6
    (sf * torch.from_numpy(z["cot"])).sum().backward()
    for k, v in p.items():
        if k in frozen:
            continue
        got, ref = golden_grad(z, k, v.grad)
        assert nerr(got, ref) <= TOL, k


def test_lstm_restatement_matches_torch_lstm():
    """the cell-by-cell LSTM against torch.nn.LSTM on a packed sequence (what HiGraph.py:135-142 calls)."""
    torch.manual_seed(3)
    lstm = torch.nn.LSTM(12, 5, num_layers=2, batch_first=True, bidirectional=True)
    lens = [4, 4, 2, 1]
    ptr = np.concatenate([[0], np.cumsum(lens)])
    feats = torch.randn(int(ptr[-1]), 12)
    seqs = [feats[ptr[i]:ptr[i + 1]] for i in range(len(lens))]
    packed = torch.nn.utils.rnn.pack_padded_sequence(torch.nn.utils.rnn.pad_sequence(seqs, batch_first=True), lens,
                                                     batch_first=True)
    out, _ = lstm(packed)
    unp, ulen = torch.nn.utils.rnn.pad_packed_sequence(out, batch_first=True)
    ref = torch.cat([unp[i][:ulen[i]] for i in range(len(lens))], 0)
    got = er.lstm_packed(feats, ptr, dict(lstm.named_parameters()), 2, True)
    assert nerr(got, ref) <= 1e-6


def test_token_positions_edge_cases():
    L = 10
    tok = np.zeros((3, L), np.int64)
    tok[1, :] = 5                       # full sentence
    tok[2, :3] = 7
    pos = er.token_positions(tok, L)
    assert pos[0].tolist() == [0] * L
    assert pos[1].tolist() == list(range(1, L + 1))
    assert pos[2].tolist() == [1, 2, 3] + [0] * (L - 3)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference only in the build container")
def test_ngram_encode_matches_live_reference_sentencoder():
    """module.Encoder.sentEncoder (Encoder.py:18-76) run live on random tokens."""
    import types
    sys.path.insert(0, "/root/reference")
    from module.Encoder import sentEncoder
    torch.manual_seed(11)
    L, V = 30, 90
    hps = types.SimpleNamespace(sent_max_len=L, word_emb_dim=300, cuda=False)
    embed = torch.nn.Embedding(V, 300, padding_idx=0)
    enc = sentEncoder(hps, embed)
    with torch.no_grad():
        for c in enc.convs:
            c.bias.normal_(0, 0.3)
    tokens, _ = fx.encoder_tokens([9], L, V, seed=2)
    ref = enc(torch.from_numpy(tokens).long())
    got = er.ngram_encode(tokens, embed.weight, enc.position_embedding.weight, [c.weight for c in enc.convs],
                          [c.bias for c in enc.convs])
    assert nerr(got, ref) <= 2e-6
    assert torch.equal(er.sinusoid_table(L + 1, 300, padding_idx=0), enc.position_embedding.weight)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference only in the build container")
@pytest.mark.parametrize("hdsg", [False, True])
def test_dropin_models_have_the_reference_state_dict(hdsg):
    """hetersumgraph_b200.HSumGraph / HSumDocGraph: same keys, key order and shapes as HiGraph.HSumGraph / HSumDocGraph
    (so a reference checkpoint loads with strict=True), 1 762 166 / 1 766 390 trainable parameters at the defaults."""
    import types
    sys.path.insert(0, "/root/reference")
    from oracle import dgl04_shim as shim
    shim.install()
    import HiGraph
    import hetersumgraph_b200 as hb
    hps = types.SimpleNamespace(n_iter=1, word_emb_dim=300, sent_max_len=100, doc_max_timesteps=50, n_feature_size=128,
                                hidden_size=64, lstm_hidden_state=128, lstm_layers=2, bidirectional=True, n_head=8,
                                atten_dropout_prob=0.1, ffn_inner_hidden_size=512, ffn_dropout_prob=0.1,
                                feat_embed_size=50, cuda=False)
    e1 = torch.nn.Embedding(1000, 300, padding_idx=0)
    e1.weight.requires_grad_(False)
    ref = (HiGraph.HSumDocGraph if hdsg else HiGraph.HSumGraph)(hps, e1)
    e2 = torch.nn.Embedding(1000, 300, padding_idx=0)
    e2.weight.requires_grad_(False)
    mine = (hb.HSumDocGraph if hdsg else hb.HSumGraph)(hps, e2)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    missing, unexpected = mine.load_state_dict(a, strict=True)
    assert not missing and not unexpected
    back = mine.state_dict()
    assert all(torch.equal(a[k], back[k]) for k in a)          # packed <-> per-head round trip is lossless
    n_train = sum(p.numel() for p in mine.parameters() if p.requires_grad)
    assert n_train == sum(p.numel() for p in ref.parameters() if p.requires_grad) == (1766390 if hdsg else 1762166)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference only in the build container")
def test_ngram_blocking_and_label_metric_match_live_reference():
    """extraction.ngram_blocking / eval_label against SLTester.ngram_blocking (Tester.py:155-184) and tools.utils.eval_label
    (the reference modules import `rouge`, absent here: a stub module stands in, nothing of it is called)."""
    import types
    sys.path.insert(0, "/root/reference")
    from oracle import dgl04_shim as shim
    shim.install()
    for name in ("rouge", "pyrouge"):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            stub.Rouge = stub.Rouge155 = object
            sys.modules[name] = stub
    from Tester import SLTester
    from tools.utils import eval_label as ref_eval
    from hetersumgraph_b200.extraction import eval_label, ngram_blocking
    rng = np.random.default_rng(3)
    words = ["w%d" % i for i in range(12)]
    tester = SLTester(model=None, m=3)
    for trial in range(30):
        n = int(rng.integers(1, 9))
        sents = [" ".join(rng.choice(words, size=int(rng.integers(0, 9)))) for _ in range(n)]
        p = torch.from_numpy(rng.permutation(n).astype(np.float32))          # distinct scores: the order is unique
        n_win, k = int(rng.integers(2, 5)), int(rng.integers(1, n + 1))
        ref = tester.ngram_blocking(sents, p, n_win, k).tolist()
        order = p.sort(descending=True)[1].tolist()
        assert ngram_blocking(sents, order, n_win, k) == ref
    a = ref_eval(torch.tensor(7), torch.tensor(10), torch.tensor(14), 50, torch.tensor(40))
    b = eval_label(7, 10, 14, 50, 40)
    assert all(abs(float(x) - y) < 1e-6 for x, y in zip(a, b))
