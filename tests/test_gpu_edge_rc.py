"""GPU tests of the recomputing backward prep (csrc/hsg_edge_rc.cu) through the C ABI: it must reproduce
hsg_edge_bwd_prep (which the oracle / golden tests pin) WITHOUT the forward's `sh`, and the update loop must give the
same results whichever prep it takes.

Tolerance: 2e-6 normalised (sh is recomputed with alpha through ex2.approx and a product instead of a division)."""
import ctypes as C

import pytest
import torch

import hetersumgraph_b200 as hb
from hetersumgraph_b200 import _lib
from hetersumgraph_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def nerr(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize("shape,hdsg,n,H,d", [("cnndm", False, 40, 6, 50), ("multinews", True, 24, 6, 50),
                                              ("tiny", False, 7, 6, 16), ("nyt50", False, 300, 6, 50),
                                              ("cnndm", False, 9, 12, 25)])
def test_recomputing_prep_equals_prep_from_saved_sh(shape, hdsg, n, H, d):
    lib = _lib.load()
    exs = syn.make_examples(n, shape, seed=3, hdsg=hdsg)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs, hdsg=hdsg))
    csc, _ = batch.csc("S2W")
    fp, ldz = _lib.edge_layout(H, d)
    assert lib.hsg_edge_bwd_prep_rc_ok(H, d, ldz) == 1
    torch.manual_seed(0)
    dev = "cuda"
    zp = torch.randn(csc.n_src, ldz, device=dev)
    q = torch.randn(10, H, device=dev)
    origin = torch.randn(csc.n_dst, H * d, device=dev)
    dx = torch.randn(csc.n_dst, H * d, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    sh, x0 = torch.empty_like(origin), torch.empty_like(origin)
    stat0 = torch.zeros(csc.n_dst, 3 * H, device=dev)
    g0 = torch.full((csc.n_dst, fp), float("nan"), device=dev)
    _lib.check(lib.hsg_edge_fwd(C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(), origin.data_ptr(), sh.data_ptr(),
                                x0.data_ptr(), stat0.data_ptr(), st))
    _lib.check(lib.hsg_edge_bwd_prep(csc.n_dst, H, d, dx.data_ptr(), None, sh.data_ptr(), g0.data_ptr(),
                                     stat0.data_ptr(), st))
    x1 = torch.empty_like(origin)
    stat1 = torch.zeros(csc.n_dst, 3 * H, device=dev)
    g1 = torch.full((csc.n_dst, fp), float("nan"), device=dev)
    _lib.check(lib.hsg_edge_fwd(C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(), origin.data_ptr(), None,
                                x1.data_ptr(), stat1.data_ptr(), st))
    _lib.check(lib.hsg_edge_bwd_prep_rc(C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(), dx.data_ptr(),
                                        g1.data_ptr(), stat1.data_ptr(), st))
    assert torch.equal(x0, x1)                                  # the forward without `sh` is the same forward
    assert torch.equal(stat0[:, :2 * H], stat1[:, :2 * H])
    assert torch.isfinite(g1).all()
    assert nerr(g1, g0) <= 2e-6, nerr(g1, g0)
    assert nerr(stat1[:, 2 * H:], stat0[:, 2 * H:]) <= 2e-6
    g2 = torch.empty_like(g1)
    _lib.check(lib.hsg_edge_bwd_prep_rc(C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(), dx.data_ptr(),
                                        g2.data_ptr(), stat1.data_ptr(), st))
    assert torch.equal(g1, g2)


def test_recomputing_prep_applicability():
    lib = _lib.load()
    _, ldz = _lib.edge_layout(6, 50)
    assert lib.hsg_edge_bwd_prep_rc_ok(6, 50, ldz) == 1
    _, ldz88 = _lib.edge_layout(8, 8)
    assert lib.hsg_edge_bwd_prep_rc_ok(8, 8, ldz88) == 0     # several lane groups per warp: general prep
    assert lib.hsg_edge_bwd_prep_rc(None, 6, 50, None, ldz, None, None, None, None, None) != 0


@pytest.mark.parametrize("n_iter,hdsg", [(1, False), (2, True)])
def test_update_loop_same_results_with_either_prep(n_iter, hdsg):
    lib = _lib.load()
    exs = syn.make_examples(20, "cnndm" if not hdsg else "multinews", seed=21, hdsg=hdsg)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs, hdsg=hdsg))
    torch.manual_seed(5)
    m = hb.WSWGATUpdateLoop(n_iter=n_iter, atten_dropout_prob=0.0, ffn_dropout_prob=0.0).cuda()
    w = torch.randn(batch.n_word, 300, device="cuda")
    s = torch.randn(batch.n_super, 64, device="cuda")
    cw, cs = torch.randn_like(w), torch.randn_like(s)

    def run():
        m.zero_grad(set_to_none=True)
        wg, sg = w.clone().requires_grad_(True), s.clone().requires_grad_(True)
        ow, os_ = m(batch, wg, sg)
        ((ow * cw).sum() + (os_ * cs).sum()).backward()
        return [ow.detach(), os_.detach(), wg.grad, sg.grad] + [p.grad.clone() for p in m.parameters()
                                                                 if p.grad is not None]

    try:
        lib.hsg_set_edge_recompute(0)
        a = run()
        lib.hsg_set_edge_recompute(1)
        b = run()
    finally:
        lib.hsg_set_edge_recompute(-1)
    assert len(a) == len(b)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for x, y in zip(b, a):
        assert nerr(x, y) <= 3e-6, nerr(x, y)


@pytest.mark.parametrize("shape,hdsg,n,kind", [("cnndm", False, 40, "S2W"), ("multinews", True, 24, "S2W"),
                                               ("nyt50", False, 64, "S2W"), ("cnndm", False, 12, "W2S")])
def test_lowdeg_forward_equals_general_forward(shape, hdsg, n, kind):
    """edge_fwd_lowdeg_kernel (rows prefetched one destination ahead, straight-line softmax up to two in-edges, online
    update beyond) against edge_fwd_kernel on the (6,50) layout: word rows (S2W) and, to exercise long edge lists and
    the implicit extra in-edges, the supernode rows of the other direction."""
    lib = _lib.load()
    exs = syn.make_examples(n, shape, seed=13, hdsg=hdsg)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs, hdsg=hdsg))
    csc, _ = batch.csc(kind)
    H, d = 6, 50
    _, ldz = _lib.edge_layout(H, d)
    torch.manual_seed(1)
    dev = "cuda"
    zp = torch.randn(csc.n_src, ldz, device=dev)
    q = torch.randn(10, H, device=dev)
    origin = torch.randn(csc.n_dst, H * d, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    try:
        for mode in (0, 1):
            lib.hsg_set_edge_fwd_lowdeg(mode)
            sh = torch.full((csc.n_dst, H * d), float("nan"), device=dev)
            x = torch.full((csc.n_dst, H * d), float("nan"), device=dev)
            stat = torch.zeros(csc.n_dst, 3 * H, device=dev)
            _lib.check(lib.hsg_edge_fwd(C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(), origin.data_ptr(),
                                        sh.data_ptr(), x.data_ptr(), stat.data_ptr(), st))
            outs.append((sh, x, stat[:, :2 * H].clone()))
        x_only = torch.full((csc.n_dst, H * d), float("nan"), device=dev)       # sh = NULL, x only
        stat = torch.zeros(csc.n_dst, 3 * H, device=dev)
        _lib.check(lib.hsg_edge_fwd(C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(), origin.data_ptr(), None,
                                    x_only.data_ptr(), stat.data_ptr(), st))
    finally:
        lib.hsg_set_edge_fwd_lowdeg(-1)
    for a, b in zip(outs[1], outs[0]):
        assert torch.isfinite(a).all()
        assert nerr(a, b) <= 2e-6, nerr(a, b)
    assert torch.equal(x_only, outs[1][1])
