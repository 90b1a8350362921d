"""GPU tests of the segment-resident edge backward (csrc/hsg_edge_seg.cu) through the C ABI: it must reproduce
hsg_edge_bwd_prep + hsg_edge_bwd (which the oracle / golden tests pin) without the forward's `sh`, bit-for-bit
reproducibly, and the update loop must give the same results whichever path it takes.

Tolerance: 2e-6 normalised (fp32 reassociation of the per-source sums: the two kernels add the same terms in a
different fixed order)."""
import ctypes as C

import numpy as np
import pytest
import torch

import hetersumgraph_b200 as hb
from hetersumgraph_b200 import _lib
from hetersumgraph_b200 import synthetic as syn
from hetersumgraph_b200.functional import _Workspace

pytestmark = pytest.mark.gpu


def nerr(a, b):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _edge_pair(batch, kind, H, d, seed=0):
    lib = _lib.load()
    csc, csc_t = batch.csc(kind)
    fp, ldz = _lib.edge_layout(H, d)
    torch.manual_seed(seed)
    dev = "cuda"
    zp = torch.randn(csc.n_src, ldz, device=dev)
    q = torch.randn(10, H, device=dev)
    origin = torch.randn(csc.n_dst, H * d, device=dev)
    dx = torch.randn(csc.n_dst, H * d, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ws = _Workspace.get(lib.hsg_edge_bwd_workspace_bytes(H), torch.device(dev), "edge")

    def general():
        sh, x = torch.empty_like(origin), torch.empty_like(origin)
        stat = torch.empty(csc.n_dst, 3 * H, device=dev)
        g = torch.empty(csc.n_dst, fp, device=dev)
        _lib.check(lib.hsg_edge_fwd(C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(), origin.data_ptr(),
                                    sh.data_ptr(), x.data_ptr(), stat.data_ptr(), st))
        _lib.check(lib.hsg_edge_bwd_prep(csc.n_dst, H, d, dx.data_ptr(), None, sh.data_ptr(), g.data_ptr(),
                                         stat.data_ptr(), st))
        dzp = torch.full((csc.n_src, ldz), float("nan"), device=dev)
        dq = torch.empty(10, H, device=dev)
        _lib.check(lib.hsg_edge_bwd(C.byref(csc_t), H, d, zp.data_ptr(), ldz, q.data_ptr(), g.data_ptr(),
                                    stat.data_ptr(), dzp.data_ptr(), dq.data_ptr(), ws.data_ptr(), ws.numel(), st))
        return x, stat, dzp, dq

    def segment(csc_use=csc):
        x = torch.empty_like(origin)
        stat = torch.zeros(csc.n_dst, 3 * H, device=dev)
        _lib.check(lib.hsg_edge_fwd(C.byref(csc_use), H, d, zp.data_ptr(), ldz, q.data_ptr(), origin.data_ptr(), None,
                                    x.data_ptr(), stat.data_ptr(), st))
        dzp = torch.full((csc.n_src, ldz), float("nan"), device=dev)
        dq = torch.empty(10, H, device=dev)
        _lib.check(lib.hsg_edge_bwd_seg(C.byref(csc_use), H, d, zp.data_ptr(), ldz, q.data_ptr(), dx.data_ptr(),
                                        stat.data_ptr(), dzp.data_ptr(), dq.data_ptr(), ws.data_ptr(), ws.numel(), st))
        return x, stat, dzp, dq

    return general, segment, csc


@pytest.mark.parametrize("shape,hdsg,n,H,d", [("cnndm", False, 40, 6, 50), ("multinews", True, 24, 6, 50),
                                              ("tiny", False, 7, 6, 16), ("nyt50", False, 300, 6, 50)])
def test_segment_backward_equals_general_backward(shape, hdsg, n, H, d):
    lib = _lib.load()
    exs = syn.make_examples(n, shape, seed=3, hdsg=hdsg)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs, hdsg=hdsg))
    assert batch.max_super_per_graph > 0
    general, segment, csc = _edge_pair(batch, "S2W", H, d)
    _, ldz = _lib.edge_layout(H, d)
    assert lib.hsg_edge_bwd_seg_ok(C.byref(csc), H, d, ldz) == 1
    x0, stat0, dzp0, dq0 = general()
    x1, stat1, dzp1, dq1 = segment()
    assert torch.equal(x0, x1)                                  # the forward without `sh` is the same forward
    assert torch.equal(stat0[:, :2 * H], stat1[:, :2 * H])
    assert torch.isfinite(dzp1).all()
    assert nerr(dzp1, dzp0) <= 2e-6, nerr(dzp1, dzp0)
    assert nerr(dq1, dq0) <= 2e-6, nerr(dq1, dq0)
    _, _, dzp2, dq2 = segment()                                 # fixed summation order: bitwise reproducible
    assert torch.equal(dzp1, dzp2) and torch.equal(dq1, dq2)


def test_segment_backward_rejects_what_it_cannot_hold():
    """No segment information, a layout with several lane groups per warp, or more source rows per graph than fit in
    shared memory: not applicable (the loop falls back to the general kernels).  A WRONG bound is visible as NaN
    gradients of exactly the graphs that break it - never an out-of-bounds access."""
    lib = _lib.load()
    exs = syn.make_examples(12, "cnndm", seed=9)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs))
    csc, _ = batch.csc("S2W")
    _, ldz = _lib.edge_layout(6, 50)
    assert lib.hsg_edge_bwd_seg_ok(C.byref(csc), 6, 50, ldz) == 1
    _, ldz88 = _lib.edge_layout(8, 8)
    assert lib.hsg_edge_bwd_seg_ok(C.byref(batch.csc("W2S")[0]), 8, 8, ldz88) == 0
    plain = _lib.CscC(csc.n_dst, csc.n_src, csc.n_edges, 0, csc.indptr, csc.nbr, csc.bin, csc.extra)
    assert lib.hsg_edge_bwd_seg_ok(C.byref(plain), 6, 50, ldz) == 0
    huge = _lib.CscC(csc.n_dst, csc.n_src, csc.n_edges, 0, csc.indptr, csc.nbr, csc.bin, csc.extra,
                     csc.seg_dst_ptr, csc.seg_src_ptr, csc.n_seg, 4096, 0, 0)
    assert lib.hsg_edge_bwd_seg_ok(C.byref(huge), 6, 50, ldz) == 0
    # a bound below the real maximum: graphs above it come back NaN, the others are right
    per = np.diff(batch.super_ptr.cpu().numpy())
    bound = int(np.sort(per)[len(per) // 2])
    assert bound < per.max()
    small = _lib.CscC(csc.n_dst, csc.n_src, csc.n_edges, 0, csc.indptr, csc.nbr, csc.bin, csc.extra,
                      csc.seg_dst_ptr, csc.seg_src_ptr, csc.n_seg, bound, 0, 0)
    general, segment, _ = _edge_pair(batch, "S2W", 6, 50)
    _, _, dzp0, _ = general()
    _, _, dzp1, _ = segment(small)
    sp = batch.super_ptr.cpu().numpy()
    for g in range(batch.n_graphs):
        rows = slice(int(sp[g]), int(sp[g + 1]))
        if per[g] > bound:
            assert torch.isnan(dzp1[rows]).all()
        else:
            assert nerr(dzp1[rows], dzp0[rows]) <= 2e-6


@pytest.mark.parametrize("n_iter,hdsg", [(1, False), (2, True)])
def test_update_loop_same_results_on_either_backward_path(n_iter, hdsg):
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
        lib.hsg_set_edge_seg(0)
        a = run()
        lib.hsg_set_edge_seg(1)
        launches0 = lib.hsg_launch_count()
        b = run()
        launches_seg = lib.hsg_launch_count() - launches0
        lib.hsg_set_edge_seg(0)
        launches0 = lib.hsg_launch_count()
        run()
        launches_gen = lib.hsg_launch_count() - launches0
    finally:
        lib.hsg_set_edge_seg(-1)
    assert launches_seg == launches_gen - n_iter            # one launch less per S2W application (no bwd-prep)
    assert len(a) == len(b)
    for x, y in zip(b, a):
        assert nerr(x, y) <= 3e-6, nerr(x, y)
