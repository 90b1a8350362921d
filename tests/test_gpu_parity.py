"""GPU parity tests (-m gpu): the sm_100a path, called through the C ABI, against the oracle
and the golden vectors generated from the unmodified reference.

Tolerances (BASELINE.json north_star): integer outputs (node/edge ids, CSR/CSC, top-m) bit-exact;
fp32 node outputs and every gradient: max|a-b| / max|b| <= 1e-5.
"""
import os

import numpy as np
import pytest
import torch

import hetersumgraph_b200 as hb
from hetersumgraph_b200 import synthetic as syn
from hetersumgraph_b200.functional import FFNFn, gemm_nn, gemm_nt, gemm_tn
from oracle import closed_form as cf
from oracle import fixtures as fx
from oracle import graph_builder_ref as gb
from oracle import wswgat_ref as wr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5


def nerr(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    e = float((a - b).abs().max() / (b.abs().max() + 1e-30))
    if os.environ.get("HSG_DEBUG_NERR") and e > 1e-2 and a.dim() == 2:      # where are the wrong elements?
        bad = (a - b).abs() > 1e-2 * float(b.abs().max())
        rows, cols = bad.any(1).nonzero().flatten(), bad.any(0).nonzero().flatten()
        print("NERR", tuple(a.shape), e, "nbad", int(bad.sum()), "rows", rows[:8].tolist(), rows[-4:].tolist(), len(rows),
              "cols", cols[:8].tolist(), cols[-4:].tolist(), len(cols), "nan", int(torch.isnan(a).sum()))
    return e


def oracle_batch(exs, hdsg):
    filt = set(syn.filter_ids().tolist())
    if hdsg:
        graphs = [gb.create_graph_hdsg(e.doc_len, e.sents.tolist(), e.doc_tokens, e.w2s, e.w2d, filt) for e in exs]
    else:
        graphs = [gb.create_graph_hsg(e.sents.tolist(), e.w2s, filt) for e in exs]
    bg, order = gb.collate(graphs)
    return bg, order


def assert_batch_equals_oracle(batch, bg):
    csc = gb.derive_csc(bg)
    cpu = lambda t: t.cpu().numpy().astype(np.int64)  # noqa: E731
    assert batch.n_word == len(csc["wnode_id"]) and batch.n_super == len(csc["snode_id"])
    assert batch.n_total_nodes == bg.n_nodes and batch.n_total_edges == bg.n_edges
    assert np.array_equal(cpu(batch.word_nid), csc["wnode_id"])
    assert np.array_equal(cpu(batch.super_nid), csc["snode_id"])
    assert np.array_equal(cpu(batch.word_wid), bg.wid[csc["wnode_id"]])
    assert np.array_equal(cpu(batch.super_type), bg.ndtype[csc["snode_id"]])
    assert np.array_equal(cpu(batch.super_extra), csc["extra_cnt"])
    for name in ("super", "word"):
        assert np.array_equal(cpu(getattr(batch, name + "_indptr")), csc[name + "_indptr"]), name
        assert np.array_equal(cpu(getattr(batch, name + "_src")), csc[name + "_src"]), name
        assert np.array_equal(cpu(getattr(batch, name + "_bin")), csc[name + "_bin"]), name
        assert np.array_equal(cpu(getattr(batch, name + "_eid")), csc[name + "_eid"]), name
    assert np.array_equal(np.diff(cpu(batch.node_ptr)), np.asarray(bg.batch_num_nodes))
    assert np.array_equal(np.diff(cpu(batch.edge_ptr)), np.asarray(bg.batch_num_edges))
    g_of_node = np.repeat(np.arange(len(bg.batch_num_nodes)), bg.batch_num_nodes)
    assert np.array_equal(cpu(batch.super_graph), g_of_node[csc["snode_id"]])


# ----------------------------------------------------------------------------- builder (K0)
@pytest.mark.parametrize("prefix,hdsg", [("hsg", False), ("hdsg", True)])
def test_builder_bit_exact_vs_reference_golden(prefix, hdsg):
    z = dict(np.load(os.path.join(GOLD, "builder.npz")))
    exs = fx.examples_from_arrays(z, prefix)
    tb = syn.pack_token_batch(exs, hdsg=hdsg)
    assert tb.order == z[prefix + "_order"].tolist()
    batch = hb.HeteroBatch.from_token_batch(tb)
    assert_batch_equals_oracle(batch, fx.graph_from_arrays(z, prefix + "_g_"))


@pytest.mark.parametrize("shape,hdsg,n,seed", [("cnndm", False, 32, 0), ("nyt50", False, 32, 1),
                                               ("multinews", True, 32, 2), ("tiny", False, 7, 3),
                                               ("tiny", True, 5, 4)])
def test_builder_bit_exact_vs_oracle_configs(shape, hdsg, n, seed):
    exs = syn.make_examples(n, shape, seed=seed, hdsg=hdsg)
    tb = syn.pack_token_batch(exs, hdsg=hdsg)
    bg, order = oracle_batch(exs, hdsg)
    assert tb.order == order
    batch = hb.HeteroBatch.from_token_batch(tb)
    assert_batch_equals_oracle(batch, bg)


def test_builder_edge_cases():
    L = 100
    # a graph whose sentences hold only filtered / PAD tokens, a single-sentence graph, an empty batch
    e0 = syn.DocExample(sents=np.zeros((2, L), np.int32), w2s=[{}, {}], labels=np.zeros(2, np.int64))
    e0.sents[0, :3] = [5, 6, 7]
    e1 = syn.DocExample(sents=np.zeros((1, L), np.int32), w2s=[{1000: 0.5, 1001: 0.25}], labels=np.ones(1, np.int64))
    e1.sents[0, :4] = [1000, 1001, 1000, 1]
    tb = syn.pack_token_batch([e0, e1])
    bg, _ = oracle_batch([e0, e1], False)
    assert_batch_equals_oracle(hb.HeteroBatch.from_token_batch(tb), bg)
    # maximum sizes: 50 sentences x 100 distinct tokens
    big = np.arange(5000, dtype=np.int32).reshape(50, 100) + 1000
    w2s = [{int(w): 0.3 for w in row if w % 3} for row in big]
    e2 = syn.DocExample(sents=big, w2s=w2s, labels=np.zeros(50, np.int64))
    tb = syn.pack_token_batch([e2])
    bg, _ = oracle_batch([e2], False)
    assert_batch_equals_oracle(hb.HeteroBatch.from_token_batch(tb), bg)


# ----------------------------------------------------------------------------- dense pieces
GEMM_TOL = {"fp32": 2e-6, "tf32x3": 3e-6, "tf32": 2e-2, "bf16": 2e-2}


@pytest.fixture(params=["fp32", "tf32x3", "tf32", "bf16"])
def gemm_mode(request):
    prev = hb.get_gemm_mode()
    hb.set_gemm_mode(request.param)
    yield request.param
    hb.set_gemm_mode(prev)


@pytest.mark.parametrize("M,N,K", [(1000, 72, 300), (257, 312, 64), (129, 512, 64), (1, 64, 512), (333, 50, 30),
                                   (4099, 300, 512), (2000, 512, 300), (77, 16, 48)])
def test_gemm_variants(M, N, K, gemm_mode):
    tol = GEMM_TOL[gemm_mode]
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda")
    ref = (A.double() @ B.double().t())
    assert nerr(gemm_nt(A, B), ref) <= tol
    assert nerr(gemm_nt(A, B, bias=bias, epi=3), torch.relu(ref + bias.double())) <= tol
    assert nerr(gemm_nt(A, B, bias=bias, R=R, epi=5), ref + bias.double() + R.double()) <= tol
    Bn = torch.randn(K, N, device="cuda")
    refn = A.double() @ Bn.double()
    assert nerr(gemm_nn(A, Bn), refn) <= tol
    assert nerr(gemm_nn(A, Bn, R=R, epi=8), torch.where(R > 0, refn, torch.zeros_like(refn))) <= tol
    assert nerr(gemm_nn(A, Bn, R=R, epi=4), refn + R.double()) <= tol
    A2 = torch.randn(M, N, device="cuda")
    Ct, cs = gemm_tn(A2, A, want_colsum=True)
    assert nerr(Ct, A2.double().t() @ A.double()) <= tol
    assert nerr(cs, A2.double().sum(0)) <= tol
    Ct2, none = gemm_tn(A2, A, want_colsum=False)
    assert none is None and nerr(Ct2, Ct) <= tol          # different tiling -> different summation order


@pytest.mark.parametrize("N,D,Dh", [(777, 64, 512), (300, 300, 512), (5, 16, 32), (1, 48, 32)])
def test_ffn_matches_torch(N, D, Dh):
    torch.manual_seed(1)
    ps = [torch.randn(Dh, D) * 0.1, torch.randn(Dh) * 0.1, torch.randn(D, Dh) * 0.1, torch.randn(D) * 0.1,
          torch.rand(D) + 0.5, torch.randn(D) * 0.1]
    x = torch.randn(N, D)
    c = torch.randn(N, D)

    def run(dev, fn):
        xs = x.clone().to(dev).detach().requires_grad_(True)
        pp = [p.clone().to(dev).detach().requires_grad_(True) for p in ps]
        out = fn(xs, *pp)
        (out * c.to(dev)).sum().backward()
        return [out, xs.grad] + [p.grad for p in pp]

    ref = run("cpu", lambda xs, w1, b1, w2, b2, g, b: cf.ffn_cf(xs, w1, b1, w2, b2, g, b))
    got = run("cuda", lambda xs, *pp: FFNFn.apply(xs, *pp))
    for a, b in zip(got, ref):
        assert nerr(a, b) <= TOL


@pytest.mark.parametrize("n", [1009, 37, 590, 593, 2500])
def test_one_launch_ffn_rows_both_cta_shapes_match_float64(n):
    """hsg_ffn_rows_fwd / _bwd (the sentence-side FFN of the update loop, GATLayer.py:35-44) at 8 rows per CTA (n <= 592)
    and 16 rows per CTA, straight through the C ABI, against the float64 formulas: hdn, r, y, stats / dr, dhp, dx and the
    (dgamma, dbeta) partials reduced by the library."""
    import ctypes as C
    from hetersumgraph_b200 import _lib
    lib = _lib.load()
    F_, Dh = 64, 512
    assert lib.hsg_ffn_rows_ok(n, F_, Dh) == 1
    torch.manual_seed(n)
    dev = "cuda"
    x = torch.randn(n, F_, device=dev)
    w1, b1 = torch.randn(Dh, F_, device=dev) * 0.1, torch.randn(Dh, device=dev) * 0.1
    w2, b2 = torch.randn(F_, Dh, device=dev) * 0.1, torch.randn(F_, device=dev) * 0.1
    gamma, beta = torch.rand(F_, device=dev) + 0.5, torch.randn(F_, device=dev) * 0.1
    dy = torch.randn(n, F_, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    ws = torch.empty(lib.hsg_layernorm_bwd_workspace_bytes(n, F_) + 1024, dtype=torch.uint8, device=dev)
    hdn, r, y, stats = (torch.full((n, Dh), float("nan"), device=dev), torch.full((n, F_), float("nan"), device=dev),
                        torch.full((n, F_), float("nan"), device=dev), torch.full((n, 2), float("nan"), device=dev))
    _lib.check(lib.hsg_ffn_rows_fwd(n, F_, Dh, p(x), p(w1), p(b1), p(w2), p(b2), p(gamma), p(beta), p(hdn), p(r), p(y),
                                    p(stats), st))
    dr, dhp, dx = (torch.full((n, F_), float("nan"), device=dev), torch.full((n, Dh), float("nan"), device=dev),
                   torch.full((n, F_), float("nan"), device=dev))
    dg, db = torch.zeros(F_, device=dev), torch.zeros(F_, device=dev)
    _lib.check(lib.hsg_ffn_rows_bwd(n, F_, Dh, p(dy), p(r), p(stats), p(gamma), p(hdn), p(w1), p(w2), p(dr), p(dhp), p(dx),
                                    p(dg), p(db), 0, p(ws), ws.numel(), st))
    torch.cuda.synchronize()
    X = x.double().requires_grad_(True)
    G = gamma.double().requires_grad_(True)
    Bt = beta.double().requires_grad_(True)
    H = torch.relu(X @ w1.double().t() + b1.double())
    Rr = H @ w2.double().t() + b2.double() + X
    Y = torch.nn.functional.layer_norm(Rr, (F_,), G, Bt, 1e-5)     # HSG_LN_EPS = nn.LayerNorm default
    H.retain_grad()
    Rr.retain_grad()
    (Y * dy.double()).sum().backward()
    mask = (hdn > 0).double()
    for got, want in ((hdn, H), (r, Rr), (y, Y), (stats[:, 0], Rr.mean(1)), (dr, Rr.grad), (dhp, H.grad * mask),
                      (dx, X.grad), (dg, G.grad), (db, Bt.grad)):
        assert nerr(got, want.detach()) <= TOL, nerr(got, want.detach())


# ----------------------------------------------------------------------------- whole path vs golden
def _load_fixture(name):
    z = dict(np.load(os.path.join(GOLD, name)))
    params = {k[2:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("p:")}
    grads = {k[3:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("gp:")}
    return z, params, grads


def _model_from_fixture(z, params):
    emb, hid, nh, ffn_h, fe, n_iter, hdsg = [int(v) for v in z["dims"]]
    m = hb.WSWGATUpdateLoop(emb, hid, nh, 0.0, ffn_h, 0.0, fe, n_iter).cuda()
    m.load_state_dict(params, strict=True)
    return m, bool(hdsg)


@pytest.mark.parametrize("name", ["wswgat_hsg_default.npz", "wswgat_hdsg_small.npz", "wswgat_hsg_small.npz"])
def test_update_loop_matches_reference_golden(name):
    z, params, gold_grads = _load_fixture(name)
    m, hdsg = _model_from_fixture(z, params)
    exs = fx.examples_from_arrays(z, "ex")
    tb = syn.pack_token_batch(exs, hdsg=hdsg)
    assert tb.order == z["order"].tolist()
    batch = hb.HeteroBatch.from_token_batch(tb)
    assert_batch_equals_oracle(batch, fx.graph_from_arrays(z, "g_"))
    w = torch.from_numpy(z["in_w"]).cuda().requires_grad_(True)
    s = torch.from_numpy(z["in_s"]).cuda().requires_grad_(True)
    ws, ss = m(batch, w, s)
    assert nerr(ws, z["out_w"]) <= TOL, nerr(ws, z["out_w"])
    assert nerr(ss, z["out_s"]) <= TOL, nerr(ss, z["out_s"])
    loss = (ws * torch.from_numpy(z["cw"]).cuda()).sum() + (ss * torch.from_numpy(z["cs"]).cuda()).sum()
    loss.backward()
    assert nerr(w.grad, z["grad_in_w"]) <= TOL
    assert nerr(s.grad, z["grad_in_s"]) <= TOL
    sd = m.state_dict(keep_vars=True)
    packed_grads = {}
    for pre in ("word2sent.", "sent2word."):
        lay = getattr(m, pre[:-1]).layer
        d = lay.out_dim
        for k in range(lay.num_heads):
            packed_grads[pre + "layer.heads.%d.fc.weight" % k] = lay.fc_weight.grad[k * d:(k + 1) * d]
            packed_grads[pre + "layer.heads.%d.feat_fc.weight" % k] = lay.feat_fc_weight.grad[k * d:(k + 1) * d]
            if lay.feat_fc_bias is not None:
                packed_grads[pre + "layer.heads.%d.feat_fc.bias" % k] = lay.feat_fc_bias.grad[k * d:(k + 1) * d]
            packed_grads[pre + "layer.heads.%d.attn_fc.weight" % k] = lay.attn_fc_weight.grad[k:k + 1]
    for k, gref in gold_grads.items():
        got = packed_grads[k] if k in packed_grads else sd[k].grad
        assert got is not None, k
        if float(gref.abs().max()) == 0.0:
            assert float(got.abs().max()) == 0.0, k
        else:
            assert nerr(got, gref) <= TOL, (k, nerr(got, gref))
        if k.endswith("attn_fc.weight"):            # dead a_dst third: exactly zero
            d = gref.shape[1] // 3
            assert float(got[:, d:2 * d].abs().max()) == 0.0


def test_bitwise_determinism():
    z, params, _ = _load_fixture("wswgat_hsg_default.npz")
    m, hdsg = _model_from_fixture(z, params)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(fx.examples_from_arrays(z, "ex"), hdsg=hdsg))
    outs = []
    for _ in range(2):
        m.zero_grad()
        w = torch.from_numpy(z["in_w"]).cuda().requires_grad_(True)
        s = torch.from_numpy(z["in_s"]).cuda().requires_grad_(True)
        ws, ss = m(batch, w, s)
        ((ws * ws).sum() + (ss * ss * ss).sum()).backward()
        outs.append([ws, ss, w.grad, s.grad] + [p.grad.clone() for p in m.parameters()])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


# ----------------------------------------------------------------------------- configs vs closed-form oracle
GRAD_TOL = {"fp32": TOL, "tf32x3": TOL, "tf32": 5e-2, "bf16": 5e-2}
KINK_TOL = 1e-2     # sanity bound on gradients downstream of a flipped leaky_relu slope (measured 1e-4 .. 3e-3)
KINK_EPS = 5e-6     # |edge logit| below this: fp32 (3xTF32 products, error ~1e-6 of max|p|) may sit on the other side of 0
FWD_TOL = {"fp32": TOL, "tf32x3": TOL, "tf32": 2e-2, "bf16": 2e-2}


@pytest.mark.parametrize("shape,hdsg,n_iter,n,seed", [("cnndm", False, 1, 8, 0), ("nyt50", False, 3, 4, 1),
                                                      ("multinews", True, 1, 4, 2)])
def test_configs_match_closed_form_oracle(shape, hdsg, n_iter, n, seed, gemm_mode):
    _check_against_closed_form(shape, hdsg, n_iter, n, seed, gemm_mode, torch.float32)


def test_large_batch_gradients_match_float64_closed_form():
    """A data-parallel shard's worth of graphs (160 CNN/DM-shaped graphs, ~60 k word rows): forward and every gradient
    against the closed form evaluated in float64.  Catches errors that grow with the batch - e.g. a single tensor-core
    accumulation chain over a weight-gradient product's rows drifts to 2e-5 at 8 k rows per split (K-chunked since).
    The batch (seed) is one in which no edge logit lies within rounding distance of leaky_relu's kink at 0: about one
    logit in 10^7 does (2 M logits here), and then fp32 and fp64 take different slopes (1 vs 0.01) for that edge, which
    shows at ~1e-4 in every gradient downstream of it - batches 5, 6 and 8 have such an edge (word row 4 222, head 6 of
    the first W2S application: logit -1.2e-7; that one row of ~57 000 is off by 6e-4, all others within 1e-5,
    scratch/large_batch_debug.py); that is a property of the function, not of the kernels."""
    seed = int(os.environ.get("HSG_LARGE_BATCH_SEED", "7"))
    _check_against_closed_form("cnndm", False, 1, 160, seed, "tf32x3", torch.float64)


@pytest.mark.parametrize("seed", [5, 6, 8, 9])
def test_large_batch_gradients_other_seeds_with_kink_allowance(seed):
    """The same 160-graph comparison on the batches the strict test above does not use.  Seeds 5, 6 and 8 contain an
    edge logit within rounding distance of leaky_relu's kink (see above).  The fp64 oracle reports its smallest
    |logit| (closed_form.KINK_LOG): a batch with none below KINK_EPS is compared at the full 1e-5 everywhere; otherwise
    the forward must still hold 1e-5, at most `kink_rows` input-gradient rows may exceed 1e-5 (each <= 1e-3) and the
    parameter gradients, all of which that edge feeds through dp / dq, are bounded at 1e-3."""
    _check_against_closed_form("cnndm", False, 1, 160, seed, "tf32x3", torch.float64, kink_rows=2)


@pytest.mark.parametrize("shape,hdsg,n_iter,seed", [("cnndm", False, 1, 0), ("nyt50", False, 3, 1),
                                                    ("multinews", True, 1, 2)])
def test_baseline_32_graph_configs_against_bucketed_port_without_device_masks(shape, hdsg, n_iter, seed):
    """BASELINE.json's three 32-graph configurations, whole update loop, against `oracle/wswgat_ref.update_loop` - the
    per-head, degree-bucketed port that is pinned bit-identically to the unmodified reference - evaluated on its OWN
    ReLU branches (no device masks).  The forward is continuous in the pre-activations, so it must hold 1e-5 as is.
    A gradient differs wherever the two arithmetics put an FFN unit with |pre-activation| ~ 1e-6 on different sides of
    0; those units are counted separately (closed form with and without the device's active set) and the mask-free
    gradient comparison is strict when there is none, bounded by 1e-3 otherwise; the strict all-gradient comparison on
    the device's active set at this size follows."""
    import hetersumgraph_b200.functional as fn
    exs = syn.make_examples(32, shape, seed=seed, hdsg=hdsg)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs, hdsg=hdsg))
    bg, _ = oracle_batch(exs, hdsg)
    torch.manual_seed(4321)
    m = hb.WSWGATUpdateLoop(n_iter=n_iter, atten_dropout_prob=0.0, ffn_dropout_prob=0.0)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    m = m.cuda()
    w, s = torch.randn(batch.n_word, 300), torch.randn(batch.n_super, 64)
    cw, cs = torch.randn(batch.n_word, 300), torch.randn(batch.n_super, 64)
    wg, sg = w.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
    fn.RELU_MASK_CAPTURE = []
    try:
        gw, gs = m(batch, wg, sg)
        masks = fn.RELU_MASK_CAPTURE
    finally:
        fn.RELU_MASK_CAPTURE = None
    ((gw * cw.cuda()).sum() + (gs * cs.cuda()).sum()).backward()
    wc, sc = w.clone().requires_grad_(True), s.clone().requires_grad_(True)
    ow, os_ = wr.update_loop(bg, wc, sc, params, n_iter)
    assert nerr(gw, ow) <= TOL and nerr(gs, os_) <= TOL
    ((ow * cw).sum() + (os_ * cs).sum()).backward()
    # units on which the device and the oracle choose different ReLU branches
    flips = []
    with torch.no_grad():
        cf.update_loop_cf(gb.derive_csc(bg), w, s, {k: v.detach() for k, v in params.items()}, n_iter, masks=masks,
                          flips=flips)
    n_flip, n_unit, pre = sum(f[0] for f in flips), sum(f[1] for f in flips), max(f[2] for f in flips)
    assert n_flip <= 2e-5 * n_unit + 1 and pre <= 1e-5, (n_flip, n_unit, pre)
    # a flipped unit gates its whole gradient path on or off (d relu is 0 or 1), and with n_iter > 0 the difference
    # spreads to the neighbours of that node: with flips the mask-free comparison can only be a sanity bound
    gtol = TOL if n_flip == 0 else 5e-2
    assert nerr(wg.grad, wc.grad) <= gtol and nerr(sg.grad, sc.grad) <= gtol, (n_flip, nerr(wg.grad, wc.grad))
    assert nerr(m._TFembed.weight.grad, params["_TFembed.weight"].grad) <= gtol
    for pre_ in ("word2sent", "sent2word"):
        mod = getattr(m, pre_)
        H = mod.layer.num_heads
        fc = torch.cat([params["%s.layer.heads.%d.fc.weight" % (pre_, k)].grad for k in range(H)], 0)
        assert nerr(mod.layer.fc_weight.grad, fc) <= gtol, (pre_, n_flip)
        assert nerr(mod.ffn.w_1.weight.grad, params[pre_ + ".ffn.w_1.weight"].grad) <= gtol, (pre_, n_flip)
        assert nerr(mod.ffn.w_2.weight.grad, params[pre_ + ".ffn.w_2.weight"].grad) <= gtol, (pre_, n_flip)
    # strict: forward and EVERY gradient on the device's active set, at the 32-graph size
    _check_against_closed_form(shape, hdsg, n_iter, 32, seed, "tf32x3", torch.float32)


def _check_against_closed_form(shape, hdsg, n_iter, n, seed, gemm_mode, oracle_dtype, kink_rows=0):
    """Whole update loop, forward and every gradient, in all three arithmetic modes.

    ReLU is the one discontinuous function on the path: an FFN unit whose pre-activation lies within rounding
    distance of 0 may take either branch depending on the arithmetic (true of the reference on two BLAS
    libraries as well).  The oracle is therefore evaluated on the device path's own ReLU active set (captured
    through functional.RELU_MASK_CAPTURE); the test also bounds how many units differ from the oracle's own
    branch choice and how close to 0 those units are."""
    import hetersumgraph_b200.functional as fn
    exs = syn.make_examples(n, shape, seed=seed, hdsg=hdsg)
    tb = syn.pack_token_batch(exs, hdsg=hdsg)
    batch = hb.HeteroBatch.from_token_batch(tb)
    bg, _ = oracle_batch(exs, hdsg)
    csc = gb.derive_csc(bg)
    torch.manual_seed(1234)
    m = hb.WSWGATUpdateLoop(n_iter=n_iter, atten_dropout_prob=0.0, ffn_dropout_prob=0.0)
    params = {k: v.detach().clone().to(oracle_dtype).requires_grad_(True) for k, v in m.state_dict().items()}
    m = m.cuda()
    w = torch.randn(batch.n_word, 300)
    s = torch.randn(batch.n_super, 64)
    cw, cs = torch.randn(batch.n_word, 300), torch.randn(batch.n_super, 64)      # random cotangents
    wg, sg = w.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
    fn.RELU_MASK_CAPTURE = []
    try:
        gw, gs = m(batch, wg, sg)
        masks = fn.RELU_MASK_CAPTURE
    finally:
        fn.RELU_MASK_CAPTURE = None
    ((gw * cw.cuda()).sum() + (gs * cs.cuda()).sum()).backward()
    wc, sc = w.clone().to(oracle_dtype).requires_grad_(True), s.clone().to(oracle_dtype).requires_grad_(True)
    flips = []
    cf.KINK_LOG = []
    try:
        ow, os_ = cf.update_loop_cf(csc, wc, sc, params, n_iter, masks=masks, flips=flips)
        min_logit = min(cf.KINK_LOG) if cf.KINK_LOG else 1.0
    finally:
        cf.KINK_LOG = None
    ((ow * cw.to(oracle_dtype)).sum() + (os_ * cs.to(oracle_dtype)).sum()).backward()
    if kink_rows and min_logit > KINK_EPS:
        kink_rows = 0          # no edge logit near leaky_relu's kink in this batch: everything at full tolerance
    n_flip, n_unit = sum(f[0] for f in flips), sum(f[1] for f in flips)
    pre_at_flip = max(f[2] for f in flips)
    limit = {"fp32": (2e-5, 1e-5), "tf32x3": (2e-5, 1e-5), "tf32": (5e-3, 2e-2), "bf16": (2e-2, 1e-1)}[gemm_mode]
    assert n_flip <= limit[0] * n_unit + 1 and pre_at_flip <= limit[1], (n_flip, n_unit, pre_at_flip)
    ftol, gtol = FWD_TOL[gemm_mode], GRAD_TOL[gemm_mode]
    assert nerr(gw, ow) <= ftol and nerr(gs, os_) <= ftol
    checks = [("d word", wg.grad, wc.grad), ("d sent", sg.grad, sc.grad),
              ("TFembed", m._TFembed.weight.grad, params["_TFembed.weight"].grad)]
    for pre in ("word2sent", "sent2word"):
        mod = getattr(m, pre)
        for k in ("w_1.weight", "w_1.bias", "w_2.weight", "w_2.bias", "layer_norm.weight", "layer_norm.bias"):
            obj = mod.ffn
            for part in k.split("."):
                obj = getattr(obj, part)
            checks.append((pre + ".ffn." + k, obj.grad, params["%s.ffn.%s" % (pre, k)].grad))
        H = mod.layer.num_heads
        cat = lambda name: torch.cat([params["%s.layer.heads.%d.%s" % (pre, k, name)].grad for k in range(H)], 0)  # noqa: E731
        checks.append((pre + ".fc", mod.layer.fc_weight.grad, cat("fc.weight")))
        checks.append((pre + ".feat_fc", mod.layer.feat_fc_weight.grad, cat("feat_fc.weight")))
        checks.append((pre + ".attn_fc", mod.layer.attn_fc_weight.grad, cat("attn_fc.weight")))
        if mod.layer.feat_fc_bias is not None:
            checks.append((pre + ".feat_fc_bias", mod.layer.feat_fc_bias.grad, cat("feat_fc.bias")))
    report = {"min_logit": min_logit}
    for name, got, ref in checks:
        if kink_rows and name in ("d word", "d sent"):
            # leaky_relu has a kink at 0: an edge logit within rounding distance of 0 takes the slope 1 in one
            # arithmetic and 0.01 in the other, which changes the input gradient of that edge's source row only (about
            # one in 10^7 logits: expected ~0.4 rows at this size; seen: one row, 6e-4).  Every other row must agree.
            d = (got.detach().cpu().double() - ref.detach().double()).abs().max(dim=1).values / float(ref.abs().max())
            report[name] = (int((d > gtol).sum()), float(d.max()))
        else:
            report[name] = nerr(got, ref)
    if kink_rows:
        # the flipped slope of that edge changes dp of its source row and dq of its TF-IDF box by a factor 100, and
        # through them every parameter gradient (seen: 1e-4 of the tensor's maximum on fc / feat_fc / attn_fc /
        # TFembed, less on the FFN weights): bounded, not compared at full tolerance
        worst = sorted(((v[1] if isinstance(v, tuple) else v), k) for k, v in report.items() if k != "min_logit")[-4:]
        for name, v in report.items():
            if name == "min_logit":
                continue
            if isinstance(v, tuple):
                assert v[0] <= kink_rows and v[1] <= KINK_TOL, (name, v, min_logit, worst)
            else:
                assert v <= KINK_TOL, (name, v, min_logit, worst)
        return
    for name, v in report.items():
        if name != "min_logit":
            assert v <= gtol, (name, v, gemm_mode, report)


def test_bucketed_reference_port_on_device_batch():
    """The per-head, degree-bucketed restatement (what DGL executes) agrees too - small batch."""
    exs = syn.make_examples(3, "tiny", seed=8)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs))
    bg, _ = oracle_batch(exs, False)
    torch.manual_seed(3)
    m = hb.WSWGATUpdateLoop(n_iter=1, atten_dropout_prob=0.0, ffn_dropout_prob=0.0)
    params = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    w, s = torch.randn(batch.n_word, 300), torch.randn(batch.n_super, 64)
    ow, os_ = wr.update_loop(bg, w, s, params, 1)
    with torch.no_grad():
        gw, gs = m(batch, w.cuda(), s.cuda())
    assert nerr(gw, ow) <= TOL and nerr(gs, os_) <= TOL


# ----------------------------------------------------------------------------- properties at large size
def test_stress_graph_properties():
    """Size-independent properties on a large bipartite graph (no oracle at this size):
    all-zero attention => sh = mean of source z scaled by deg/(deg+extra); rows sum check; determinism."""
    n_word, n_super, n_edges = 65536, 8192, 262144
    word, sup, bins, extra = syn.stress_edges(n_word, n_super, n_edges, seed=4, extra=64)
    (sip, ssrc, sbin, seid), (wip, wsrc, wbin, weid) = hb.csc_pair_from_edges(word, sup, bins, n_word, n_super)
    batch = hb.HeteroBatch.from_csc_arrays(sip, ssrc, sbin, extra, wip, wsrc, wbin)
    torch.manual_seed(0)
    lay = hb.MultiHeadLayer(300, 8, 8, 0.0, 50, layer=hb.WSGATLayer).cuda()
    with torch.no_grad():
        lay.attn_fc_weight.zero_()
    T = torch.randn(10, 50, device="cuda")
    batch.set_tfidf_embedding(T)
    h = torch.randn(n_word, 300, device="cuda")
    with torch.no_grad():
        sh = lay(batch, h)
        sh2 = lay(batch, h)
    assert torch.equal(sh, sh2)
    z = h @ lay.fc_weight.t()
    deg = torch.from_numpy(np.diff(sip)).cuda().float()
    dst = torch.repeat_interleave(torch.arange(n_super, device="cuda"), torch.from_numpy(np.diff(sip)).cuda())
    agg = torch.zeros(n_super, 64, device="cuda").index_add_(0, dst, z[torch.from_numpy(ssrc).cuda()])
    want = agg / (deg + 64.0).unsqueeze(1)
    assert nerr(sh, want) <= TOL


# ----------------------------------------------------------------------------- whole-loop C entry points
def _loop_inputs(n_iter, hdsg=False, n=6, seed=11):
    exs = syn.make_examples(n, "tiny", seed=seed, hdsg=hdsg)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs, hdsg=hdsg))
    torch.manual_seed(5)
    m = hb.WSWGATUpdateLoop(n_iter=n_iter, atten_dropout_prob=0.0, ffn_dropout_prob=0.0).cuda()
    w = torch.randn(batch.n_word, 300, device="cuda")
    s = torch.randn(batch.n_super, 64, device="cuda")
    cw, cs = torch.randn_like(w), torch.randn_like(s)
    return batch, m, w, s, cw, cs


def _run_loop(m, fwd, batch, w, s, cw, cs, use_w=True, use_s=True):
    m.zero_grad(set_to_none=True)
    wg, sg = w.clone().requires_grad_(True), s.clone().requires_grad_(True)
    ow, os_ = fwd(batch, wg, sg)
    loss = 0.0
    if use_w:
        loss = loss + (ow * cw).sum()
    if use_s:
        loss = loss + (os_ * cs).sum()
    loss.backward()
    return [ow.detach(), os_.detach(), wg.grad, sg.grad] + [p.grad.clone() if p.grad is not None else torch.zeros_like(p)
                                                            for p in m.parameters()]


@pytest.mark.parametrize("n_iter,hdsg", [(0, False), (1, False), (2, True), (3, False)])
def test_whole_loop_call_equals_per_application_path(n_iter, hdsg):
    """hsg_update_loop_fwd/bwd (one C call each way, in-kernel gradient accumulation) against the per-application
    autograd path (hsg_wswgat_fwd/bwd + autograd's own accumulation): same kernels, so only the order of the
    accumulating adds may differ."""
    batch, m, w, s, cw, cs = _loop_inputs(n_iter, hdsg)
    a = _run_loop(m, m.forward, batch, w, s, cw, cs)
    b = _run_loop(m, m.forward_per_application, batch, w, s, cw, cs)
    for x, y in zip(a, b):
        if float(y.abs().max()) == 0.0:
            assert float(x.abs().max()) == 0.0
        else:
            assert nerr(x, y) <= 2e-6, nerr(x, y)


@pytest.mark.parametrize("use_w,use_s", [(True, False), (False, True)])
def test_whole_loop_call_single_cotangent(use_w, use_s):
    batch, m, w, s, cw, cs = _loop_inputs(1)
    a = _run_loop(m, m.forward, batch, w, s, cw, cs, use_w, use_s)
    b = _run_loop(m, m.forward_per_application, batch, w, s, cw, cs, use_w, use_s)
    for x, y in zip(a, b):
        if float(y.abs().max()) == 0.0:
            assert float(x.abs().max()) == 0.0
        else:
            assert nerr(x, y) <= 2e-6, nerr(x, y)


def test_fused_grad_accumulation_adds_into_existing_grads():
    """fuse_grad_accumulation: the kernels ADD into .grad (flat arena views) - equals autograd's result, and a
    second backward doubles it."""
    from hetersumgraph_b200.dist import FlatGradArena
    batch, m, w, s, cw, cs = _loop_inputs(1)
    ref = _run_loop(m, m.forward, batch, w, s, cw, cs)[4:]
    arena = FlatGradArena(m.parameters())
    m.fuse_grad_accumulation = True
    for rep in (1, 2):
        wg, sg = w.clone().requires_grad_(True), s.clone().requires_grad_(True)
        ow, os_ = m(batch, wg, sg)
        ((ow * cw).sum() + (os_ * cs).sum()).backward()
        for p, r in zip(m.parameters(), ref):
            assert p.grad.data_ptr() >= arena.flat.data_ptr()
            if float(r.abs().max()) == 0.0:
                assert float(p.grad.abs().max()) == 0.0
            else:
                assert nerr(p.grad, rep * r) <= 2e-6


# ----------------------------------------------------------------------------- readout / loss / top-m / Adam
@pytest.mark.parametrize("hdsg", [False, True])
def test_fused_loss_matches_reference_loss(hdsg):
    """hsg_head_fwd/bwd against wh + CrossEntropyLoss + per-graph sum + mean (train.py:114-119) in stock PyTorch."""
    from hetersumgraph_b200.path_model import HSGPath, fused_loss, graph_loss
    exs = syn.make_examples(6, "tiny", seed=21, hdsg=hdsg)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs, hdsg=hdsg))
    torch.manual_seed(2)
    model = HSGPath(n_iter=1, hdsg=hdsg).cuda()
    n_sent = batch.labels.shape[0]
    sf = torch.randn(n_sent, 64, device="cuda")
    res = []
    for fused in (False, True):
        model.zero_grad(set_to_none=True)
        s = sf.clone().requires_grad_(True)
        if fused:
            loss, logits = fused_loss(model, batch, s, n_graphs_global=7)
        else:
            logits = model(batch, s)
            loss = graph_loss(batch, logits, batch.labels, 7)
        (loss * 1.7).backward()
        res.append([loss.detach(), logits.detach(), s.grad] + [p.grad.clone() for p in model.parameters() if p.grad is not None])
    assert len(res[0]) == len(res[1])
    for a, b in zip(res[1], res[0]):
        assert nerr(a, b) <= TOL, nerr(a, b)


@pytest.mark.parametrize("hdsg", [False, True])
@pytest.mark.parametrize("n_graphs,shape", [(6, "tiny"), (32, "cnndm")])
def test_one_launch_head_is_bit_identical_to_forward_then_backward(hdsg, n_graphs, shape):
    """hsg_head_fwd_bwd (what the fused training step launches) against hsg_head_fwd + hsg_head_bwd(gout = NULL): loss,
    logits, d_state and the ACCUMULATED d wh, bit for bit; twice in a row (the ticket counter returns to zero)."""
    from hetersumgraph_b200.functional import SentenceLossFn, head_fwd_bwd
    from hetersumgraph_b200.path_model import _Ctx
    if hdsg and shape == "cnndm":
        shape = "multinews"
    exs = syn.make_examples(n_graphs, shape, seed=3, hdsg=hdsg)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs, hdsg=hdsg))
    torch.manual_seed(5)
    n_super = batch.n_super
    state = torch.randn(n_super, 64, device="cuda")
    w = torch.randn(2, 128 if hdsg else 64, device="cuda") * 0.3
    b = torch.randn(2, device="cuda")
    g0 = (torch.randn_like(w), torch.randn_like(b))
    for rep in range(2):
        ta = (g0[0].clone(), g0[1].clone())
        c = _Ctx([False] * 7)
        loss_a, logits_a = SentenceLossFn.forward(c, batch, 7, ta, state, w, b, batch.labels)
        d_a = SentenceLossFn.backward(c, None, None)[3]
        tb = (g0[0].clone(), g0[1].clone())
        loss_b, logits_b, d_b = head_fwd_bwd(batch, 7, tb, state, w, b, batch.labels)
        torch.cuda.synchronize()
        assert torch.equal(loss_a, loss_b) and torch.equal(logits_a, logits_b) and torch.equal(d_a, d_b)
        assert torch.equal(ta[0], tb[0]) and torch.equal(ta[1], tb[1])
        assert float((ta[0] - g0[0]).abs().max()) > 0.0


def test_topm_bit_exact_vs_torch_topk():
    from hetersumgraph_b200.functional import topm
    rng = np.random.default_rng(0)
    counts = [1, 7, 50, 3, 100, 2]
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    logits = torch.randn(int(ptr[-1]), 2)
    for m in (1, 3, 5):
        got = topm(logits.cuda(), torch.from_numpy(ptr).cuda(), m).cpu()
        for g, c in enumerate(counts):
            want = torch.topk(logits[ptr[g]:ptr[g + 1], 1], min(m, c))[1].tolist()      # Tester.py:128
            assert got[g, :len(want)].tolist() == want
            assert (got[g, len(want):] == -1).all()


@pytest.mark.parametrize("clip", [0.0, 0.5])
def test_fused_adam_matches_torch(clip):
    from hetersumgraph_b200.functional import FusedAdam
    torch.manual_seed(0)
    n = 100003
    p0 = torch.randn(n)
    ref = torch.nn.Parameter(p0.clone().cuda())
    opt = torch.optim.Adam([ref], lr=5e-4)
    mine = p0.clone().cuda()
    g = torch.zeros(n, device="cuda")
    fa = FusedAdam(mine, g, lr=5e-4, max_grad_norm=clip)
    for step in range(5):
        grad = torch.randn(n, device="cuda") * (0.1 + step)
        ref.grad = grad.clone()
        if clip > 0:
            torch.nn.utils.clip_grad_norm_([ref], clip)
        opt.step()
        g.copy_(grad)
        fa.step()
        assert nerr(mine, ref) <= 1e-6


# ----------------------------------------------------------------------------- training-mode dropout
def _drop_masks(seed, n_apps, start, p_attn, p_ffn, n_word, n_super):
    """explicit multipliers of every application, read back from the device generator (hsg_dropout_mask)."""
    from hetersumgraph_b200.functional import dropout_keep_mask
    out = []
    for i in range(n_apps):
        kind = (i + start) % 2
        H, d, in_dim, n_src, n_dst = (8, 8, 300, n_word, n_super) if kind == 0 else (6, 50, 64, n_super, n_word)
        drop = {}
        if p_attn > 0:
            keep = dropout_keep_mask(H * n_src * in_dim, p_attn, seed, 2 * i).cpu().float()
            drop["attn"] = keep.view(H, n_src, in_dim) / (1.0 - p_attn)
        if p_ffn > 0:
            keep = dropout_keep_mask(n_dst * H * d, p_ffn, seed, 2 * i + 1).cpu().float()
            drop["ffn"] = keep.view(n_dst, H * d) / (1.0 - p_ffn)
        out.append(drop)
    return out


@pytest.mark.parametrize("p_attn,p_ffn,n_iter", [(0.1, 0.1, 1), (0.3, 0.0, 1), (0.0, 0.2, 2)])
def test_training_dropout_matches_oracle_with_same_masks(p_attn, p_ffn, n_iter):
    """Training-mode dropout (per-head input dropout GATStackLayer.py:56, FFN dropout GATLayer.py:41-42): the
    device path draws its masks from a counter-based generator; the oracle is evaluated with the very same masks
    (read back through hsg_dropout_mask), forward and all gradients."""
    import hetersumgraph_b200.functional as fn
    import hetersumgraph_b200.modules as md
    exs = syn.make_examples(5, "tiny", seed=31)
    tb = syn.pack_token_batch(exs)
    batch = hb.HeteroBatch.from_token_batch(tb)
    bg, _ = oracle_batch(exs, False)
    csc = gb.derive_csc(bg)
    torch.manual_seed(77)
    m = hb.WSWGATUpdateLoop(n_iter=n_iter, atten_dropout_prob=p_attn, ffn_dropout_prob=p_ffn)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    m = m.cuda().train()
    w, s = torch.randn(batch.n_word, 300), torch.randn(batch.n_super, 64)
    cw, cs = torch.randn_like(w), torch.randn_like(s)
    wg, sg = w.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
    fn.RELU_MASK_CAPTURE, md.DROPOUT_SEED_LOG = [], []
    try:
        gw, gs = m(batch, wg, sg)
        relu_masks, (seed, n_apps, start) = fn.RELU_MASK_CAPTURE, md.DROPOUT_SEED_LOG[0]
    finally:
        fn.RELU_MASK_CAPTURE, md.DROPOUT_SEED_LOG = None, None
    ((gw * cw.cuda()).sum() + (gs * cs.cuda()).sum()).backward()
    drops = _drop_masks(seed, n_apps, start, p_attn, p_ffn, batch.n_word, batch.n_super)
    if p_attn > 0:      # the generator really drops about p of the entries, independently per head
        frac = float((drops[0]["attn"] == 0).float().mean())
        assert abs(frac - p_attn) < 0.02
        assert not torch.equal(drops[0]["attn"][0], drops[0]["attn"][1])
    wc, sc = w.clone().requires_grad_(True), s.clone().requires_grad_(True)
    ow, os_ = cf.update_loop_cf(csc, wc, sc, params, n_iter, masks=relu_masks, drops=drops)
    ((ow * cw).sum() + (os_ * cs).sum()).backward()
    assert nerr(gw, ow) <= TOL and nerr(gs, os_) <= TOL
    assert nerr(wg.grad, wc.grad) <= TOL and nerr(sg.grad, sc.grad) <= TOL
    assert nerr(m._TFembed.weight.grad, params["_TFembed.weight"].grad) <= TOL
    for pre in ("word2sent", "sent2word"):
        mod = getattr(m, pre)
        H = mod.layer.num_heads
        cat = lambda name: torch.cat([params["%s.layer.heads.%d.%s" % (pre, k, name)].grad for k in range(H)], 0)  # noqa: E731
        assert nerr(mod.layer.fc_weight.grad, cat("fc.weight")) <= TOL, pre
        assert nerr(mod.layer.attn_fc_weight.grad, cat("attn_fc.weight")) <= TOL, pre
        assert nerr(mod.ffn.w_1.weight.grad, params[pre + ".ffn.w_1.weight"].grad) <= TOL, pre
        assert nerr(mod.ffn.w_2.weight.grad, params[pre + ".ffn.w_2.weight"].grad) <= TOL, pre
        assert nerr(mod.ffn.w_2.bias.grad, params[pre + ".ffn.w_2.bias"].grad) <= TOL, pre
    # a second forward draws a different mask; eval mode is deterministic and mask-free
    gw2, _ = m(batch, wg, sg)
    assert not torch.equal(gw2, gw)
    m.eval()
    with torch.no_grad():
        e1, e2 = m(batch, wg, sg), m(batch, wg, sg)
    assert torch.equal(e1[0], e2[0]) and torch.equal(e1[1], e2[1])


def test_standalone_wswgat_module_with_dropout():
    """One WSWGAT module called directly (GAT.py:45-59) in training mode, both layer types."""
    import hetersumgraph_b200.functional as fn
    import hetersumgraph_b200.modules as md
    exs = syn.make_examples(4, "tiny", seed=41)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs))
    bg, _ = oracle_batch(exs, False)
    csc = gb.derive_csc(bg)
    T = torch.randn(10, 50)
    batch.set_tfidf_embedding(T.cuda())
    w, s = torch.randn(batch.n_word, 300), torch.randn(batch.n_super, 64)
    for kind, in_dim, out_dim, H in (("W2S", 300, 64, 8), ("S2W", 64, 300, 6)):
        torch.manual_seed(3)
        mod = hb.WSWGAT(in_dim, out_dim, H, 0.2, 512, 0.1, 50, kind)
        params = {"x." + k: v.detach().clone() for k, v in mod.state_dict().items()}
        mod = mod.cuda().train()
        fn.RELU_MASK_CAPTURE, md.DROPOUT_SEED_LOG = [], []
        try:
            got = mod(batch, w.cuda(), s.cuda())
            relu_masks, (seed, n_apps, start) = fn.RELU_MASK_CAPTURE, md.DROPOUT_SEED_LOG[0]
        finally:
            fn.RELU_MASK_CAPTURE, md.DROPOUT_SEED_LOG = None, None
        assert (n_apps, start) == (1, 0 if kind == "W2S" else 1)
        drops = _drop_masks(seed, 1, start, 0.2, 0.1, batch.n_word, batch.n_super)
        want = cf.wswgat_cf(csc, w, s, params, "x.", kind, T, relu_masks[0], None, drops[0])
        assert nerr(got, want) <= TOL, kind


def test_standalone_ffn_training_dropout():
    """PositionwiseFeedForward called on its own in training mode with p > 0 (GATLayer.py:35-44):
    LayerNorm(x + dropout(w_2 relu(w_1 x))) with the library's mask, forward and every gradient against torch."""
    import hetersumgraph_b200.functional as fn
    import hetersumgraph_b200.modules as md
    torch.manual_seed(2)
    ffn = hb.PositionwiseFeedForward(64, 512, 0.3).cuda().train()
    x = torch.randn(700, 64, device="cuda", requires_grad=True)
    cot = torch.randn(700, 64, device="cuda")
    torch.manual_seed(77)
    seed = md._next_dropout_seed()
    torch.manual_seed(77)                                   # the module draws the same seed
    fn.RELU_MASK_CAPTURE = None
    y = ffn(x)
    (y * cot).sum().backward()
    got = [y.detach(), x.grad.clone()] + [p.grad.clone() for p in ffn.parameters()]
    mult = fn.dropout_keep_mask(700 * 64, 0.3, seed, 1).float().view(700, 64) / 0.7
    assert 0.6 < float((mult > 0).float().mean()) < 0.8
    w1, b1, w2, b2, gamma, beta = [t.detach().double().requires_grad_(True) for t in ffn.packed()]
    xd = x.detach().double().requires_grad_(True)
    inner = torch.relu(xd @ w1.t() + b1) @ w2.t() + b2
    ref = torch.nn.functional.layer_norm(xd + inner * mult.double(), (64,), gamma, beta, 1e-5)
    (ref * cot.double()).sum().backward()
    want = [ref.detach(), xd.grad, w1.grad.view_as(ffn.w_1.weight), b1.grad, w2.grad.view_as(ffn.w_2.weight), b2.grad,
            gamma.grad, beta.grad]
    for a, b in zip(got, want):
        assert nerr(a, b) <= TOL, nerr(a, b)
    ffn.eval()                                              # evaluation mode: no mask
    assert nerr(ffn(x.detach()), torch.nn.functional.layer_norm(
        xd + (torch.relu(xd @ w1.t() + b1) @ w2.t() + b2), (64,), gamma, beta, 1e-5).detach()) <= TOL


@pytest.mark.parametrize("kind", ["W2S", "S2W"])
def test_standalone_multihead_layer_training_dropout(kind):
    """MultiHeadLayer called on its own in training mode with p > 0: every head aggregates ITS OWN dropout(h)
    (GATStackLayer.py:56).  Heads are independent, so head k of the result must equal head k of the p = 0 layer fed
    with the k-th masked input - forward and every gradient."""
    import hetersumgraph_b200.functional as fn
    import hetersumgraph_b200.modules as md
    exs = syn.make_examples(5, "tiny", seed=43)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs))
    T = torch.randn(10, 50, device="cuda", requires_grad=True)
    batch.set_tfidf_embedding(T)
    in_dim, d, H, n_src, n_dst = (300, 8, 8, batch.n_word, batch.n_super) if kind == "W2S" else \
        (64, 50, 6, batch.n_super, batch.n_word)
    torch.manual_seed(4)
    lay = hb.MultiHeadLayer(in_dim, d, H, 0.25, 50, layer=hb.WSGATLayer if kind == "W2S" else hb.SWGATLayer).cuda()
    h = torch.randn(n_src, in_dim, device="cuda", requires_grad=True)
    cot = torch.randn(n_dst, H * d, device="cuda")
    params = list(lay.parameters()) + [T]

    def grads():
        out = [h.grad.clone()] + [p.grad.clone() for p in params]
        h.grad = None
        for p in params:
            p.grad = None
        return out

    lay.train()
    torch.manual_seed(91)
    seed = md._next_dropout_seed()
    torch.manual_seed(91)
    got = lay(batch, h)
    (got * cot).sum().backward()
    got_g = grads()
    mult = fn.dropout_keep_mask(H * n_src * in_dim, 0.25, seed, 0).float().view(H, n_src, in_dim) / 0.75
    lay.eval()
    want = torch.cat([lay(batch, h * mult[k])[:, k * d:(k + 1) * d] for k in range(H)], dim=1)
    (want * cot).sum().backward()
    want_g = grads()
    assert nerr(got, want) <= TOL
    for a, b in zip(got_g, want_g):
        assert nerr(a, b) <= TOL, nerr(a, b)


@pytest.mark.parametrize("kind,H,d", [("W2S", 8, 8), ("S2W", 6, 50)])
def test_edge_bwd_row_mappings_agree(kind, H, d):
    """hsg_edge_bwd's row mappings - warp per row with the groups sharing the row's edge list (baseline), each group
    walking its own row (low-degree rows), a whole CTA per row (few high-degree rows) - give the same dzp / dq up to
    summation order."""
    import ctypes as C
    from hetersumgraph_b200 import _lib
    from hetersumgraph_b200.functional import _Workspace
    lib = _lib.load()
    exs = syn.make_examples(16, "cnndm", seed=5)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs))
    csc, csc_t = batch.csc(kind)
    fp, ldz = _lib.edge_layout(H, d)
    torch.manual_seed(0)
    dev = "cuda"
    zp = torch.randn(csc.n_src, ldz, device=dev)
    q = torch.randn(10, H, device=dev)
    origin = torch.randn(csc.n_dst, H * d, device=dev)
    sh, x = torch.empty_like(origin), torch.empty_like(origin)
    stat = torch.empty(csc.n_dst, 3 * H, device=dev)
    g = torch.empty(csc.n_dst, fp, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hsg_edge_fwd(C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(), origin.data_ptr(), sh.data_ptr(),
                                x.data_ptr(), stat.data_ptr(), st))
    _lib.check(lib.hsg_edge_bwd_prep(csc.n_dst, H, d, origin.data_ptr(), None, sh.data_ptr(), g.data_ptr(),
                                     stat.data_ptr(), st))
    ws = _Workspace.get(lib.hsg_edge_bwd_workspace_bytes(H), torch.device(dev), "edge")
    outs = []
    try:
        for rowpar, blockrow, asy in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)):
            lib.hsg_set_edge_rowpar(rowpar)
            lib.hsg_set_edge_blockrow(blockrow)
            lib.hsg_set_edge_bwd_async(asy)
            dzp = torch.full((csc.n_src, ldz), float("nan"), device=dev)
            dq = torch.empty(10, H, device=dev)
            _lib.check(lib.hsg_edge_bwd(C.byref(csc_t), H, d, zp.data_ptr(), ldz, q.data_ptr(), g.data_ptr(),
                                        stat.data_ptr(), dzp.data_ptr(), dq.data_ptr(), ws.data_ptr(), ws.numel(), st))
            outs.append((dzp, dq))
    finally:
        lib.hsg_set_edge_rowpar(-1)
        lib.hsg_set_edge_blockrow(-1)
        lib.hsg_set_edge_bwd_async(-1)
    for dzp, dq in outs[1:]:
        assert torch.isfinite(dzp).all()
        assert nerr(dzp, outs[0][0]) <= 2e-6 and nerr(dq, outs[0][1]) <= 2e-6


@pytest.mark.parametrize("hdsg", [False, True])
def test_fused_train_step_equals_autograd_path(hdsg):
    """path_model.FusedTrainStep (forward + backward driven without the autograd engine) against
    fused_loss(...).backward(): same loss, same d_sent_feature, same gradients in the flat arena (HSG and HDSG)."""
    from hetersumgraph_b200.dist import FlatGradArena
    from hetersumgraph_b200.path_model import FusedTrainStep, HSGPath, fused_loss
    if hdsg:
        exs = syn.make_examples(6, "multinews", seed=51, hdsg=True)
    else:
        exs = syn.make_examples(6, "tiny", seed=51)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs, hdsg=hdsg))
    torch.manual_seed(9)
    model = HSGPath(n_iter=1, hdsg=hdsg).cuda()
    arena = FlatGradArena(model.parameters())
    sf = torch.randn(batch.labels.shape[0], 64, device="cuda")
    model.loop.fuse_grad_accumulation = True
    s1 = sf.clone().requires_grad_(True)
    loss1, logits1 = fused_loss(model, batch, s1, 6, fuse_grad_accumulation=True)
    loss1.backward()
    g1 = arena.flat.clone()
    arena.flat.zero_()
    loss2, logits2, d_sf = FusedTrainStep(model, 6)(batch, sf)
    assert torch.equal(loss1.detach(), loss2) and torch.equal(logits1, logits2)
    assert torch.equal(s1.grad, d_sf)
    assert torch.equal(g1, arena.flat)
    assert float(g1.abs().max()) > 0


@pytest.mark.parametrize("name,hdsg", [("s2s_hsg.npz", False), ("s2s_hdsg.npz", True)])
def test_s2s_layer_matches_reference_golden(name, hdsg):
    """WSWGAT(..., "S2S") (GAT.py:38-39,50-52) on the device against the reference's own classes (golden)."""
    z = dict(np.load(os.path.join(GOLD, name)))
    hid, nh, ffn_h, _ = [int(v) for v in z["dims"]]
    params = {k[2:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("p:")}
    m = hb.WSWGAT(hid, hid, nh, 0.0, ffn_h, 0.0, 50, "S2S").cuda()
    m.load_state_dict(params, strict=True)
    exs = fx.examples_from_arrays(z, "ex")
    tb = syn.pack_token_batch(exs, hdsg=hdsg)
    assert tb.order == z["order"].tolist()
    batch = hb.HeteroBatch.from_token_batch(tb)
    s = torch.from_numpy(z["in_s"]).cuda().requires_grad_(True)
    out = m(batch, s, s)
    assert nerr(out, z["out_s"]) <= TOL, nerr(out, z["out_s"])
    (out * torch.from_numpy(z["cs"]).cuda()).sum().backward()
    assert nerr(s.grad, z["grad_in_s"]) <= TOL
    d = m.layer.out_dim
    for k in range(nh):
        assert nerr(m.layer.fc_weight.grad[k * d:(k + 1) * d], z["gp:layer.heads.%d.fc.weight" % k]) <= TOL
        ga = m.layer.attn_fc_weight.grad[k:k + 1]
        assert nerr(ga, z["gp:layer.heads.%d.attn_fc.weight" % k]) <= TOL
        assert float(ga[:, :d].abs().max()) == 0.0                     # a_src multiplies the zero-filled word z
    for key in ("ffn.w_1.weight", "ffn.w_1.bias", "ffn.w_2.weight", "ffn.w_2.bias", "ffn.layer_norm.weight",
                "ffn.layer_norm.bias"):
        obj = m
        for part in key.split("."):
            obj = getattr(obj, part)
        assert nerr(obj.grad, z["gp:" + key]) <= TOL, key


def test_hdsg_doc_init_matches_reference_restatement():
    """hsg_doc_mean / hsg_super_assemble / hsg_doc_init_bwd against set_dnfeature + dn_feature_proj
    (HiGraph.py:196-203, 231-244) restated in oracle/closed_form.py:doc_init_ref."""
    from hetersumgraph_b200.functional import DocInitFn
    exs = syn.make_examples(5, "multinews", seed=71, hdsg=True)
    tb = syn.pack_token_batch(exs, hdsg=True)
    batch = hb.HeteroBatch.from_token_batch(tb)
    bg, _ = oracle_batch(exs, True)
    csc = gb.derive_csc(bg)
    # known answers for the row maps from the literal graph
    row = np.full(bg.n_nodes, -1, np.int64)
    row[csc["snode_id"]] = np.arange(len(csc["snode_id"]))
    assert np.array_equal(batch.sentence_rows().cpu().numpy(), row[csc["sent_id"]])
    assert np.array_equal(batch.doc_rows().cpu().numpy(), row[csc["doc_id"]])
    torch.manual_seed(0)
    n_sent = batch.labels.shape[0]
    sf = torch.randn(n_sent, 64)
    W = torch.randn(64, 64) * 0.2
    c = torch.randn(batch.n_super, 64)
    sfg, Wg = sf.cuda().requires_grad_(True), W.cuda().requires_grad_(True)
    got = DocInitFn.apply(batch, sfg, Wg)
    (got * c.cuda()).sum().backward()
    sfc, Wc = sf.clone().requires_grad_(True), W.clone().requires_grad_(True)
    want = cf.doc_init_ref(sfc, batch.sentence_rows().cpu(), batch.doc_rows().cpu(), batch.sent_doc_row.cpu(), Wc,
                           batch.n_super)
    (want * c).sum().backward()
    assert nerr(got, want) <= TOL and nerr(sfg.grad, sfc.grad) <= TOL and nerr(Wg.grad, Wc.grad) <= TOL


def test_update_loop_degenerate_batches():
    """Ragged / empty inputs through the whole-loop entry points: a batch whose graphs have NO word nodes at all
    (every token filtered: n_word = 0, no edges) and a batch mixing such a graph with ordinary ones."""
    L = 100
    empty = syn.DocExample(sents=np.zeros((3, L), np.int32), w2s=[{}, {}, {}], labels=np.zeros(3, np.int64))
    empty.sents[:, :2] = [[5, 6], [7, 8], [9, 10]]                      # stop-word surrogates: filtered ids
    normal = syn.make_examples(2, "tiny", seed=81)
    for exs in ([empty], [empty, normal[0], normal[1]]):
        tb = syn.pack_token_batch(exs)
        batch = hb.HeteroBatch.from_token_batch(tb)
        bg, _ = oracle_batch(exs, False)
        assert_batch_equals_oracle(batch, bg)
        csc = gb.derive_csc(bg)
        torch.manual_seed(4)
        m = hb.WSWGATUpdateLoop(n_iter=1, atten_dropout_prob=0.0, ffn_dropout_prob=0.0)
        params = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
        m = m.cuda()
        w, s = torch.randn(batch.n_word, 300), torch.randn(batch.n_super, 64)
        cs = torch.randn(batch.n_super, 64)
        wg, sg = w.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
        gw, gs = m(batch, wg, sg)
        (gs * cs.cuda()).sum().backward()
        wc, sc = w.clone().requires_grad_(True), s.clone().requires_grad_(True)
        ow, os_ = cf.update_loop_cf(csc, wc, sc, params, 1)
        (os_ * cs).sum().backward()
        assert gw.shape == ow.shape and nerr(gs, os_) <= TOL
        if batch.n_word > 0:
            assert nerr(gw, ow) <= TOL
        assert nerr(sg.grad, sc.grad) <= TOL
        assert torch.isfinite(m.word2sent.layer.fc_weight.grad).all()
        assert nerr(m.word2sent.ffn.w_1.weight.grad, params["word2sent.ffn.w_1.weight"].grad) <= TOL


ALL_EDGE_CONFIGS = [(8, 8), (6, 50), (8, 16), (6, 16), (8, 32), (6, 32), (4, 4), (6, 8), (4, 16), (1, 64), (16, 4),
                    (2, 32), (4, 32), (12, 25)]


@pytest.fixture
def row_mapping(request):
    """Forces the edge kernels' row mapping: "auto"; "shared" = the lane groups of a warp share one row; "rowpar" = a
    row per lane group (forward and backward); "async" = backward with cp.async gathers through a shared-memory ring."""
    from hetersumgraph_b200 import _lib
    lib = _lib.load()
    fwd_rp, bwd_rp, blk, asy = {"auto": (-1, -1, -1, -1), "shared": (0, 0, 0, 0), "rowpar": (1, 1, 0, 0),
                                "async": (0, 0, 0, 1)}[request.param]
    lib.hsg_set_edge_fwd_rowpar(fwd_rp)
    lib.hsg_set_edge_rowpar(bwd_rp)
    lib.hsg_set_edge_blockrow(blk)
    lib.hsg_set_edge_bwd_async(asy)
    yield request.param
    lib.hsg_set_edge_fwd_rowpar(-1)
    lib.hsg_set_edge_rowpar(-1)
    lib.hsg_set_edge_blockrow(-1)
    lib.hsg_set_edge_bwd_async(-1)


@pytest.mark.parametrize("H,d", ALL_EDGE_CONFIGS)
@pytest.mark.parametrize("kind", ["W2S", "S2W"])
@pytest.mark.parametrize("row_mapping", ["auto", "shared", "rowpar", "async"], indirect=True)
def test_every_instantiated_head_shape_matches_closed_form(H, d, kind, row_mapping):
    """Every (heads, head_dim) instantiation of the edge kernels (HSG_EDGE_CONFIGS), both layer types, every row
    mapping of the forward and backward kernels: forward and all gradients of MultiHeadLayer against the closed form
    on a small HSG batch."""
    exs = syn.make_examples(4, "tiny", seed=91)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs))
    bg, _ = oracle_batch(exs, False)
    csc = gb.derive_csc(bg)
    in_dim, fe = 20, 12
    torch.manual_seed(H * 100 + d)
    lay = hb.MultiHeadLayer(in_dim, d, H, 0.0, fe, layer=hb.WSGATLayer if kind == "W2S" else hb.SWGATLayer)
    W, Wf, a = lay.fc_weight.detach().clone(), lay.feat_fc_weight.detach().clone(), lay.attn_fc_weight.detach().clone()
    bf = lay.feat_fc_bias.detach().clone() if lay.feat_fc_bias is not None else torch.zeros(H * d)
    lay = lay.cuda()
    T = torch.randn(10, fe)
    batch.set_tfidf_embedding(T.cuda().requires_grad_(True))
    n_src, n_dst = (batch.n_word, batch.n_super) if kind == "W2S" else (batch.n_super, batch.n_word)
    h = torch.randn(n_src, in_dim)
    c = torch.randn(n_dst, H * d)
    hg = h.cuda().requires_grad_(True)
    out = lay(batch, hg)
    (out * c.cuda()).sum().backward()
    hc = h.clone().requires_grad_(True)
    Wc, Wfc, bfc, ac, Tc = (t.clone().requires_grad_(True) for t in (W, Wf, bf, a, T))
    if kind == "W2S":
        ref = cf.multi_head_cf(hc, n_dst, csc["super_indptr"], csc["super_src"], csc["super_bin"], csc["extra_cnt"],
                               Wc, Wfc, bfc, ac, Tc)
    else:
        ref = cf.multi_head_cf(hc, n_dst, csc["word_indptr"], csc["word_src"], csc["word_bin"], csc["extra_cnt_word"],
                               Wc, Wfc, bfc, ac, Tc)
    (ref * c).sum().backward()
    assert nerr(out, ref) <= TOL, (H, d, kind, nerr(out, ref))
    assert nerr(hg.grad, hc.grad) <= TOL
    assert nerr(lay.fc_weight.grad, Wc.grad) <= TOL
    assert nerr(lay.feat_fc_weight.grad, Wfc.grad) <= TOL
    assert nerr(lay.attn_fc_weight.grad, ac.grad) <= TOL
    assert nerr(batch.tfidfembed_weight.grad, Tc.grad) <= TOL
    if lay.feat_fc_bias is not None:
        assert nerr(lay.feat_fc_bias.grad, bfc.grad) <= TOL


@pytest.mark.parametrize("M,N1,N2", [(200000, 512, 300), (150001, 64, 512), (40000, 300, 72)])
def test_weight_gradient_product_stays_in_tolerance_for_long_reductions(M, N1, N2):
    """hsg_gemm_tn over a data-parallel shard's worth of rows (config 5): the tensor core's fp32 accumulation truncates,
    so a single accumulation chain over 10^4..10^5 rows drifts to 1e-4 (measured); the kernel walks every split in
    K-chunks of 1 024 rows and adds the chunks in fp32 - the product and the column sums stay <= 1e-5 of an fp64
    evaluation, like the FFMA mode."""
    g = torch.Generator(device="cuda").manual_seed(M)
    A = torch.randn(M, N1, device="cuda", generator=g) + 0.25
    B = torch.randn(M, N2, device="cuda", generator=g) - 0.1
    ref = A.double().t() @ B.double()
    ref_cs = A.double().sum(0)
    Cm, cs = gemm_tn(A, B, want_colsum=True)
    assert nerr(Cm, ref) <= TOL, nerr(Cm, ref)
    assert nerr(cs, ref_cs) <= TOL, nerr(cs, ref_cs)
    C2, cs2 = gemm_tn(A, B, want_colsum=True)
    assert torch.equal(Cm, C2) and torch.equal(cs, cs2)          # deterministic


@pytest.mark.parametrize("M,N,K", [(11817, 512, 300), (11817, 300, 512), (5000, 72, 300), (700, 300, 72),
                                   (513, 112, 300), (4099, 304, 64), (12001, 64, 512)])
@pytest.mark.parametrize("mode", ["tf32x3", "tf32", "bf16"])
def test_cta_pair_gemm_is_bit_identical_to_single_cta(M, N, K, mode):
    """hsg_gemm_tc2.cu (tcgen05 cta_group::2: 256-row tiles over a CTA pair, B tile shared between the two SMs) against
    the single-CTA kernel: same hi/lo split, same k order, same accumulators -> identical bits; and both against
    float64.  Shapes: the FFN / projection products of the 32-graph step, ragged edges, odd tile counts."""
    from hetersumgraph_b200 import _lib
    lib = _lib.load()
    prev = hb.get_gemm_mode()
    hb.set_gemm_mode(mode)
    try:
        torch.manual_seed(3)
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(N, K, device="cuda")
        Bn = torch.randn(K, N, device="cuda")
        bias = torch.randn(N, device="cuda")
        R = torch.randn(M, N, device="cuda")
        outs = []
        for pair in (0, 1):
            _lib.check(lib.hsg_set_gemm_pair(pair))
            outs.append((gemm_nt(A, B), gemm_nt(A, B, bias=bias, epi=3), gemm_nt(A, B, bias=bias, R=R, epi=5),
                         gemm_nn(A, Bn), gemm_nn(A, Bn, R=R, epi=8), gemm_nn(A, Bn, R=R, epi=4)))
        for x, y in zip(*outs):
            assert torch.equal(x, y)
        tol = GEMM_TOL[mode]
        assert nerr(outs[1][0], A.double() @ B.double().t()) <= tol
        assert nerr(outs[1][3], A.double() @ Bn.double()) <= tol
    finally:
        _lib.check(lib.hsg_set_gemm_pair(1))
        hb.set_gemm_mode(prev)


@pytest.mark.parametrize("M,N1,N2", [(4099, 300, 512), (11817, 512, 300), (2000, 72, 300)])
@pytest.mark.parametrize("colsum", [False, True])
def test_weight_gradient_workspace_is_large_enough(M, N1, N2, colsum):
    """hsg_gemm_tn_workspace_bytes must cover BOTH plans of the tensor-core product: without the column-sum column there
    are fewer column tiles and therefore more splits (4 099 x 300 x 512: 12 against 9).  Round 1 sized for the
    column-sum plan only and the other plan's partials ran past the buffer.  Here the workspace is exactly the advertised
    size with a sentinel-filled guard behind it."""
    import ctypes as C
    from hetersumgraph_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(5)
    A = torch.randn(M, N1, device="cuda")
    B = torch.randn(M, N2, device="cuda")
    nbytes = int(lib.hsg_gemm_tn_workspace_bytes(M, N1, N2))
    guard = 8 << 20
    buf = torch.full((nbytes + guard,), 0x5A, dtype=torch.uint8, device="cuda")
    Cm = torch.empty(N1, N2, device="cuda")
    cs = torch.empty(N1, device="cuda") if colsum else None
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.hsg_gemm_tn(M, N1, N2, A.data_ptr(), N1, B.data_ptr(), N2, Cm.data_ptr(), N2,
                               cs.data_ptr() if colsum else None, buf.data_ptr(), nbytes, C.c_void_p(st)))
    torch.cuda.synchronize()
    assert bool((buf[nbytes:] == 0x5A).all()), "the product wrote past its workspace"
    assert nerr(Cm, A.double().t() @ B.double()) <= 3e-6
    if colsum:
        assert nerr(cs, A.double().sum(0)) <= 3e-6
