"""GPU parity tests (-m gpu) of the sentence encoder: csrc/hsg_encoder.cu + encoder.py through
the C ABI against (i) the golden vectors generated from the UNMODIFIED reference (tests/golden/encoder_*.npz) and
(ii) the oracle restatement (oracle/encoder_ref.py) on larger seeded inputs.

Tolerance (BASELINE.json north_star): max|a-b| / max|b| <= 1e-5 for outputs and every parameter gradient.
"""
import os

import numpy as np
import pytest
import torch

from hetersumgraph_b200 import synthetic as syn
from hetersumgraph_b200.encoder import EncoderPlan, NgramEncodeFn, SentenceEncoder
from oracle import encoder_ref as er
from oracle import fixtures as fx

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5
FROZEN = ("ngram_enc.embed.weight", "sent_pos_embed.weight", "ngram_enc.position_embedding.weight")


def nerr(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def make_encoder(params, dims):
    vocab, emb, L, doc_max, n_feature, hidden, lstm_hidden = dims
    embed = torch.nn.Embedding(vocab, emb, padding_idx=0)
    embed.weight.requires_grad_(False)
    enc = SentenceEncoder(embed, emb, L, doc_max, n_feature, hidden, lstm_hidden, lstm_dropout=0.0)
    missing, unexpected = enc.load_state_dict(params, strict=True)
    assert not missing and not unexpected
    return enc.cuda()


@pytest.mark.parametrize("name", ["encoder_small.npz", "encoder_default.npz"])
def test_sentence_encoder_matches_reference_golden(name):
    """forward, and every parameter gradient, against HSumGraph.set_snfeature + n_feature_proj of the reference."""
    z, params = fx.load_encoder_fixture(GOLD, name)
    dims = [int(v) for v in z["dims"]]
    enc = make_encoder(params, dims)
    plan = EncoderPlan(z["tokens"], z["graph_sent_ptr"], "cuda")
    ngram = enc.ngram(plan)
    assert nerr(ngram, z["ngram"]) <= TOL
    sf = enc(plan)
    assert nerr(sf, z["sent_feature"]) <= TOL
    (sf * torch.from_numpy(z["cot"]).cuda()).sum().backward()
    for k, p in enc.named_parameters():
        if k in FROZEN:
            assert p.grad is None
            continue
        got, ref = fx.golden_grad(z, k, p.grad.cpu())
        assert nerr(got, ref) <= TOL, k


@pytest.mark.parametrize("kind,n_graphs", [("cnndm", 6), ("nyt50", 3)])
def test_ngram_cnn_matches_oracle_on_synthetic_batches(kind, n_graphs):
    """n-gram CNN alone (the custom kernels), CNN/DM- and NYT50-shaped token matrices: forward and all convolution
    gradients against the oracle; also bitwise run-to-run determinism."""
    exs = syn.make_examples(n_graphs, kind, seed=4)
    tb = syn.pack_token_batch(exs)
    V = 50000
    g = torch.Generator().manual_seed(9)
    embed_w = torch.randn(V, 300, generator=g)
    embed_w[0] = 0
    pos = er.sinusoid_table(tb.tokens.shape[1] + 1, 300, padding_idx=0)
    ws = [(torch.randn(50, 1, h, 300, generator=g) / (h * 300) ** 0.5) for h in er.KERNEL_HEIGHTS]
    bs = [0.2 * torch.randn(50, generator=g) for _ in er.KERNEL_HEIGHTS]
    cot = torch.randn(tb.tokens.shape[0], 300, generator=g)
    # oracle (CPU)
    wso = [w.clone().requires_grad_(True) for w in ws]
    bso = [b.clone().requires_grad_(True) for b in bs]
    ref = er.ngram_encode(tb.tokens, embed_w, pos, wso, bso)
    (ref * cot).sum().backward()
    # device
    plan = EncoderPlan.from_token_batch(tb, "cuda")
    assert plan.n_rows < tb.tokens.size            # compact rows: fewer than S * L
    outs = []
    for _ in range(2):
        conv = []
        for w, b in zip(ws, bs):
            conv += [w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)]
        out = NgramEncodeFn.apply(plan, None, embed_w.cuda(), pos.cuda(), *conv)
        (out * cot.cuda()).sum().backward()
        outs.append((out.detach().clone(), [c.grad.clone() for c in conv]))
    out, grads = outs[0]
    assert nerr(out, ref) <= TOL
    for i in range(6):
        assert nerr(grads[2 * i], wso[i].grad) <= TOL, "dW h=%d" % (i + 2)
        assert nerr(grads[2 * i + 1], bso[i].grad) <= TOL, "db h=%d" % (i + 2)
    assert torch.equal(out, outs[1][0]) and all(torch.equal(a, b) for a, b in zip(grads, outs[1][1]))


def test_ngram_cnn_pad_row_and_interior_zero_ids():
    """a non-zero PAD embedding row and zero ids INSIDE a sentence (len counts non-zero ids, Encoder.py:58; the rows
    behind the last real id are what is deduplicated) still match the oracle."""
    L, V = 40, 64
    rng = np.random.default_rng(0)
    tokens = np.zeros((9, L), np.int32)
    for s in range(9):
        n = int(rng.integers(0, L + 1))
        tokens[s, :n] = rng.integers(0, V, size=n)          # zeros allowed inside
    tokens[3] = rng.integers(1, V, size=L)                   # full sentence
    tokens[5] = 0                                            # empty sentence
    ptr = np.asarray([0, 4, 7, 9], np.int32)
    g = torch.Generator().manual_seed(1)
    embed_w = torch.randn(V, 300, generator=g)               # PAD row NOT zero
    pos = er.sinusoid_table(L + 1, 300, padding_idx=0)
    ws = [(torch.randn(50, 1, h, 300, generator=g) / (h * 300) ** 0.5) for h in er.KERNEL_HEIGHTS]
    bs = [0.2 * torch.randn(50, generator=g) for _ in er.KERNEL_HEIGHTS]
    ref = er.ngram_encode(tokens, embed_w, pos, ws, bs)
    plan = EncoderPlan(tokens, ptr, "cuda")
    conv = []
    for w, b in zip(ws, bs):
        conv += [w.cuda(), b.cuda()]
    out = NgramEncodeFn.apply(plan, None, embed_w.cuda(), pos.cuda(), *conv)
    assert nerr(out, ref) <= TOL


@pytest.mark.parametrize("H,n_in,layers,bidir,lens", [(128, 300, 2, True, [9, 6, 6, 1]), (8, 300, 2, True, [5, 2]),
                                                       (32, 64, 1, False, [4, 4, 3]), (100, 40, 2, True, [50, 17]),
                                                       (128, 300, 2, True, [3, 0])])
def test_lstm_kernels_match_oracle(H, n_in, layers, bidir, lens):
    """csrc/hsg_lstm.cu (LstmFn) against the cell-by-cell restatement of nn.LSTM on packed per-graph sequences:
    output, input gradient and every weight / bias gradient; bitwise determinism."""
    from hetersumgraph_b200.encoder import LstmFn
    torch.manual_seed(5)
    ref = torch.nn.LSTM(n_in, H, num_layers=layers, batch_first=True, bidirectional=bidir)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    S = int(ptr[-1])
    x = torch.randn(S, n_in)
    cot = torch.randn(S, H * (2 if bidir else 1))
    xo = x.clone().requires_grad_(True)
    po = {k: v.detach().clone().requires_grad_(True) for k, v in ref.named_parameters()}
    out_ref = er.lstm_packed(xo, ptr, po, layers, bidir)
    (out_ref * cot).sum().backward()
    gptr = torch.from_numpy(ptr).cuda()
    cfg = (len(lens), H, layers, 2 if bidir else 1, 0.0, 0)
    runs = []
    for _ in range(2):
        xd = x.cuda().requires_grad_(True)
        pd = [v.detach().clone().cuda().requires_grad_(True) for v in ref._flat_weights]
        out = LstmFn.apply(xd, gptr, cfg, *pd)
        (out * cot.cuda()).sum().backward()
        runs.append((out.detach().clone(), xd.grad.clone(), [q.grad.clone() for q in pd]))
    out, dx, grads = runs[0]
    assert nerr(out, out_ref) <= TOL
    assert nerr(dx, xo.grad) <= TOL
    for name, g in zip(ref._flat_weights_names, grads):
        assert nerr(g, po[name].grad) <= TOL, name
    assert torch.equal(out, runs[1][0]) and torch.equal(dx, runs[1][1])
    assert all(torch.equal(a, b) for a, b in zip(grads, runs[1][2]))
    # weight-gradient products on the side stream (default) vs everything on one stream: bitwise identical
    LstmFn.overlap_weight_grads = False
    try:
        xd = x.cuda().requires_grad_(True)
        pd = [v.detach().clone().cuda().requires_grad_(True) for v in ref._flat_weights]
        (LstmFn.apply(xd, gptr, cfg, *pd) * cot.cuda()).sum().backward()
        torch.cuda.synchronize()
        assert torch.equal(xd.grad, dx) and all(torch.equal(q.grad, g) for q, g in zip(pd, grads))
    finally:
        LstmFn.overlap_weight_grads = True


def test_lstm_kernels_agree_with_torch_lstm_on_a_packed_sequence():
    """torch.nn.LSTM on a PackedSequence (cuDNN; the call the reference makes, HiGraph.py:136-141) gives the same LSTM
    features as the package's recurrence kernels.  The library call lives HERE: the package has no cuDNN path."""
    z, params = fx.load_encoder_fixture(GOLD, "encoder_small.npz")
    enc = make_encoder(params, [int(v) for v in z["dims"]])
    plan = EncoderPlan(z["tokens"], z["graph_sent_ptr"], "cuda")
    with torch.no_grad():
        ngram = enc.ngram(plan)
        a = enc.lstm_feature(plan, ngram)
        packed = torch.nn.utils.rnn.PackedSequence(ngram.index_select(0, plan.perm), plan.batch_sizes)
        # cuDNN's RNN would otherwise run its products in TF32 (error class 1e-3, outside the fp32 bound of 1e-5)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            out, _ = enc.lstm(packed)
        b = out.data.index_select(0, plan.inv_perm)
    assert nerr(a, b) <= TOL


def test_fused_gradient_accumulation_is_bitwise_identical():
    """fuse_grad_accumulation: the kernels ADD every parameter gradient of the encoder straight into existing .grad
    buffers (33 autograd accumulations fewer per step) - same bits as the autograd route, and a second backward
    accumulates (2x)."""
    z, params = fx.load_encoder_fixture(GOLD, "encoder_default.npz")
    dims = [int(v) for v in z["dims"]]
    plan = EncoderPlan(z["tokens"], z["graph_sent_ptr"], "cuda")
    cot = torch.from_numpy(z["cot"]).cuda()
    enc = make_encoder(params, dims)
    (enc(plan) * cot).sum().backward()
    ref = {k: p.grad.clone() for k, p in enc.named_parameters() if p.requires_grad}
    enc2 = make_encoder(params, dims)
    for p in enc2.parameters():
        if p.requires_grad:
            p.grad = torch.zeros_like(p)
    enc2.fuse_grad_accumulation = True
    (enc2(plan) * cot).sum().backward()
    torch.cuda.synchronize()
    for k, p in enc2.named_parameters():
        if p.requires_grad:
            assert torch.equal(p.grad, ref[k]), k
    (enc2(plan) * cot).sum().backward()
    torch.cuda.synchronize()
    for k, p in enc2.named_parameters():
        if p.requires_grad:
            assert nerr(p.grad, 2 * ref[k]) <= 1e-6, k
