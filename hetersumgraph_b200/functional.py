"""Autograd functions of the WSWGAT path; every stage is a call into libhsg_b200.so.

Forward / backward data flow of one WSWGAT application (module/GAT.py:45-59):

  prep      W_aug, q           <- hsg_attn_prep_fwd       (attn_fc folded into fc / the TF-IDF table)
  proj      zp = h_src W_aug^T <- hsg_gemm_nt             (all heads' fc + p = a_src.z in one product)
  edge      sh, x, stat        <- hsg_edge_fwd            (logits, edge_softmax, aggregation, ELU, +origin)
  ffn       hdn, r, out        <- hsg_gemm_nt x2, hsg_layernorm_fwd
  backward  mirrors it with hsg_layernorm_bwd, hsg_gemm_nn / hsg_gemm_tn, hsg_edge_bwd_prep, hsg_edge_bwd,
            hsg_attn_prep_bwd; attention is recomputed from (p, q, m, den), no per-edge tensor is stored.

There is no eager/PyTorch fallback: without the CUDA library or a B200 these raise.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import EPI_ADD, EPI_BIAS, EPI_RELU, EPI_RELU_MASK


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _st():
    """cudaStream_t of torch's current stream on the current device (the raw getter skips ~15 us of Python)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _f32c(t):
    if t.dtype != torch.float32:
        raise TypeError("hsg_b200 expects float32 tensors, got %s" % t.dtype)
    if not t.is_cuda:
        raise RuntimeError("hsg_b200: tensor is not on a CUDA device (no CPU fallback)")
    return t.contiguous()


def round_up(x, m):
    return (x + m - 1) // m * m


class _Workspace:
    """Per-device scratch reused across calls (stream-ordered use only)."""
    _bufs = {}

    @classmethod
    def get(cls, nbytes, device, tag):
        key = (device, tag)
        buf = cls._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
            cls._bufs[key] = buf
        return buf


def gemm_nt(A, B, bias=None, R=None, epi=0, N=None, out=None):
    """C = A @ B[:N].T (+bias, relu, +R / relu-mask R)."""
    lib = _lib.load()
    M, K = A.shape
    N = B.shape[0] if N is None else N
    Cm = out if out is not None else torch.empty(M, N, dtype=torch.float32, device=A.device)
    _lib.check(lib.hsg_gemm_nt(M, N, K, _p(A), A.stride(0), _p(B), B.stride(0), _p(Cm), Cm.stride(0), _p(bias), _p(R),
                               R.stride(0) if R is not None else 0, epi, _st()))
    return Cm


def gemm_nn(A, B, R=None, epi=0):
    """C = A @ B (+R / relu-mask R)."""
    lib = _lib.load()
    M, K = A.shape
    N = B.shape[1]
    Cm = torch.empty(M, N, dtype=torch.float32, device=A.device)
    _lib.check(lib.hsg_gemm_nn(M, N, K, _p(A), A.stride(0), _p(B), B.stride(0), _p(Cm), Cm.stride(0), _p(R),
                               R.stride(0) if R is not None else 0, epi, _st()))
    return Cm


def gemm_tn(A, B, want_colsum=False, ws_tag="tn", out=None, cs_out=None):
    """C = A.T @ B  (and column sums of A), deterministic.  `ws_tag` names the scratch buffer: calls that may run
    concurrently on different streams must use different tags.  out / cs_out: existing buffers the product (and the
    column sums) are ADDED to (gradient accumulation straight into .grad, hsg_gemm_tn_acc)."""
    lib = _lib.load()
    M, N1 = A.shape
    N2 = B.shape[1]
    nbytes = lib.hsg_gemm_tn_workspace_bytes(M, N1, N2)
    ws = _Workspace.get(nbytes, A.device, ws_tag)
    if out is not None:
        _lib.check(lib.hsg_gemm_tn_acc(M, N1, N2, _p(A), A.stride(0), _p(B), B.stride(0), _p(out), out.stride(0),
                                       _p(cs_out), 1, _p(ws), ws.numel(), _st()))
        return out, cs_out
    Cm = torch.empty(N1, N2, dtype=torch.float32, device=A.device)
    cs = torch.empty(N1, dtype=torch.float32, device=A.device) if want_colsum else None
    _lib.check(lib.hsg_gemm_tn(M, N1, N2, _p(A), A.stride(0), _p(B), B.stride(0), _p(Cm), Cm.stride(0), _p(cs), _p(ws),
                               ws.numel(), _st()))
    return Cm, cs


# --------------------------------------------------------------------------------------------
# multi-head attention layer: proj + edge kernel
# --------------------------------------------------------------------------------------------
def _mh_forward(csc, H, d, h_src, origin, W, Wf, bf, a, T):
    lib = _lib.load()
    F = H * d
    fp, ldz = _lib.edge_layout(H, d)
    in_dim = h_src.shape[1]
    dev = h_src.device
    n_dst = csc.n_dst
    W_aug = torch.empty(ldz, in_dim, dtype=torch.float32, device=dev)
    q = torch.empty(_N_BINS, H, dtype=torch.float32, device=dev)
    _lib.check(lib.hsg_attn_prep_fwd(H, d, in_dim, Wf.shape[1], ldz, _p(W), _p(Wf), _p(bf), _p(a), _p(T), _p(W_aug),
                                     _p(q), _st()))
    zp = gemm_nt(h_src, W_aug)
    sh = torch.empty(n_dst, F, dtype=torch.float32, device=dev)
    x = torch.empty(n_dst, F, dtype=torch.float32, device=dev) if origin is not None else None
    stat = torch.empty(n_dst, 3 * H, dtype=torch.float32, device=dev)
    _lib.check(lib.hsg_edge_fwd(C.byref(csc), H, d, _p(zp), ldz, _p(q), _p(origin), _p(sh), _p(x), _p(stat), _st()))
    return W_aug, q, zp, sh, x, stat


def _mh_backward(csc_t, H, d, h_src, W, Wf, bf, a, T, W_aug, q, zp, sh, stat, dx=None, dsh=None):
    """Returns (dh_src, dW, dWf, dbf, da, dT)."""
    lib = _lib.load()
    F = H * d
    ldz = zp.shape[1]
    dev = zp.device
    n_dst, n_src = sh.shape[0], zp.shape[0]
    fp, _ = _lib.edge_layout(H, d)
    g = torch.empty(n_dst, fp, dtype=torch.float32, device=dev)
    _lib.check(lib.hsg_edge_bwd_prep(n_dst, H, d, _p(dx), _p(dsh), _p(sh), _p(g), _p(stat), _st()))
    dzp = torch.empty(n_src, ldz, dtype=torch.float32, device=dev)
    dq = torch.empty(_N_BINS, H, dtype=torch.float32, device=dev)
    ws = _Workspace.get(lib.hsg_edge_bwd_workspace_bytes(H), dev, "edge")
    _lib.check(lib.hsg_edge_bwd(C.byref(csc_t), H, d, _p(zp), ldz, _p(q), _p(g), _p(stat), _p(dzp), _p(dq), _p(ws),
                                ws.numel(), _st()))
    dh_src = gemm_nn(dzp, W_aug)
    dW_aug, _ = gemm_tn(dzp, h_src)
    dW = torch.empty_like(W)
    dWf = torch.empty_like(Wf)
    dbf = torch.empty_like(bf) if bf is not None else None
    da = torch.empty_like(a)
    dT = torch.empty_like(T)
    _lib.check(lib.hsg_attn_prep_bwd(H, d, W.shape[1], Wf.shape[1], ldz, _p(W), _p(Wf), _p(bf), _p(a), _p(T),
                                     _p(dW_aug), _p(dq), _p(dW), _p(dWf), _p(dbf), _p(da), _p(dT), _st()))
    return dh_src, dW, dWf, dbf, da, dT


_N_BINS = 10


class MultiHeadFn(torch.autograd.Function):
    """MultiHeadLayer.forward(g, h) -> cat_k head_k(g, h)   (GATStackLayer.py:55-59)."""

    @staticmethod
    def forward(ctx, batch, kind, H, d, h_src, W, Wf, bf, a, T):
        _lib.require_device()
        csc, csc_t = batch.csc(kind)
        h_src, W, Wf, a, T = (_f32c(t) for t in (h_src, W, Wf, a, T))
        bf = _f32c(bf) if bf is not None else None
        if h_src.shape[0] != csc.n_src:
            raise ValueError("%s: input has %d rows, graph has %d source nodes" % (kind, h_src.shape[0], csc.n_src))
        W_aug, q, zp, sh, _, stat = _mh_forward(csc, H, d, h_src, None, W, Wf, bf, a, T)
        ctx.batch, ctx.kind, ctx.H, ctx.d, ctx.has_bf = batch, kind, H, d, bf is not None
        ctx.save_for_backward(h_src, W, Wf, bf if bf is not None else W.new_empty(0), a, T, W_aug, q, zp, sh, stat)
        return sh

    @staticmethod
    def backward(ctx, dsh):
        h_src, W, Wf, bf, a, T, W_aug, q, zp, sh, stat = ctx.saved_tensors
        bf = bf if ctx.has_bf else None
        _, csc_t = ctx.batch.csc(ctx.kind)
        dh, dW, dWf, dbf, da, dT = _mh_backward(csc_t, ctx.H, ctx.d, h_src, W, Wf, bf, a, T, W_aug, q, zp, sh,
                                                stat.clone(), dsh=_f32c(dsh))
        return None, None, None, None, dh, dW, dWf, dbf, da, dT


_ROW_HEAD = {}


def _waug_row_head(H, d, device):
    """head of every row of W_aug (rows = lane-interleaved columns of z, then the H rows of p, then zero padding)"""
    key = (H, d, str(device))
    if key not in _ROW_HEAD:
        lib = _lib.load()
        fp, ldz = _lib.edge_layout(H, d)
        rh = [0] * ldz
        for c in range(H * d):
            rh[lib.hsg_edge_perm(H, d, c)] = c // d
        for k in range(H):
            rh[fp + k] = k
        _ROW_HEAD[key] = torch.tensor(rh, dtype=torch.long, device=device)
    return _ROW_HEAD[key]


class MultiHeadDropFn(torch.autograd.Function):
    """MultiHeadLayer.forward(g, h) in TRAINING mode with input dropout p > 0 on its own: every head projects its OWN
    dropout(h) (GATStackLayer.py:56).  Same construction as the update loop's dropout path (hsg_dropout.cu): the H
    masked copies of the input side by side, [n_src, H in], times the head-blocked weight [ldz, H in]; masks from the
    library generator (stream 0 = attention input of application 0, element ((head n_src) + row) in + col)."""

    @staticmethod
    def forward(ctx, batch, kind, H, d, h_src, W, Wf, bf, a, T, p, seed):
        _lib.require_device()
        lib = _lib.load()
        csc, _ = batch.csc(kind)
        h_src, W, Wf, a, T = (_f32c(t) for t in (h_src, W, Wf, a, T))
        bf = _f32c(bf) if bf is not None else None
        n_src, in_dim = h_src.shape
        if n_src != csc.n_src:
            raise ValueError("%s: input has %d rows, graph has %d source nodes" % (kind, n_src, csc.n_src))
        dev = h_src.device
        fp, ldz = _lib.edge_layout(H, d)
        W_aug = torch.empty(ldz, in_dim, dtype=torch.float32, device=dev)
        q = torch.empty(_N_BINS, H, dtype=torch.float32, device=dev)
        _lib.check(lib.hsg_attn_prep_fwd(H, d, in_dim, Wf.shape[1], ldz, _p(W), _p(Wf), _p(bf), _p(a), _p(T), _p(W_aug),
                                         _p(q), _st()))
        mult = _keep_mult(H * n_src * in_dim, p, seed, 0, dev).view(H, n_src, in_dim)
        a_exp = (h_src.unsqueeze(0) * mult).permute(1, 0, 2).reshape(n_src, H * in_dim).contiguous()
        row_head = _waug_row_head(H, d, dev)
        w_blk = torch.zeros(ldz, H, in_dim, dtype=torch.float32, device=dev)
        w_blk[torch.arange(ldz, device=dev), row_head] = W_aug
        w_blk = w_blk.view(ldz, H * in_dim)
        zp = gemm_nt(a_exp, w_blk)
        sh = torch.empty(csc.n_dst, H * d, dtype=torch.float32, device=dev)
        stat = torch.empty(csc.n_dst, 3 * H, dtype=torch.float32, device=dev)
        _lib.check(lib.hsg_edge_fwd(C.byref(csc), H, d, _p(zp), ldz, _p(q), None, _p(sh), None, _p(stat), _st()))
        ctx.batch, ctx.kind, ctx.H, ctx.d, ctx.has_bf = batch, kind, H, d, bf is not None
        ctx.save_for_backward(W, Wf, bf if bf is not None else W.new_empty(0), a, T, q, zp, sh, stat, mult, a_exp, w_blk)
        return sh

    @staticmethod
    def backward(ctx, dsh):
        lib = _lib.load()
        W, Wf, bf, a, T, q, zp, sh, stat, mult, a_exp, w_blk = ctx.saved_tensors
        bf = bf if ctx.has_bf else None
        H, d = ctx.H, ctx.d
        _, csc_t = ctx.batch.csc(ctx.kind)
        dev = zp.device
        n_dst, n_src, ldz = sh.shape[0], zp.shape[0], zp.shape[1]
        in_dim = W.shape[1]
        fp, _ = _lib.edge_layout(H, d)
        g = torch.empty(n_dst, fp, dtype=torch.float32, device=dev)
        stat = stat.clone()
        _lib.check(lib.hsg_edge_bwd_prep(n_dst, H, d, None, _p(_f32c(dsh)), _p(sh), _p(g), _p(stat), _st()))
        dzp = torch.empty(n_src, ldz, dtype=torch.float32, device=dev)
        dq = torch.empty(_N_BINS, H, dtype=torch.float32, device=dev)
        ws = _Workspace.get(lib.hsg_edge_bwd_workspace_bytes(H), dev, "edge")
        _lib.check(lib.hsg_edge_bwd(C.byref(csc_t), H, d, _p(zp), ldz, _p(q), _p(g), _p(stat), _p(dzp), _p(dq), _p(ws),
                                    ws.numel(), _st()))
        d_exp = gemm_nn(dzp, w_blk)                                                   # [n_src, H in]
        dh = (d_exp.view(n_src, H, in_dim).permute(1, 0, 2) * mult).sum(0)
        dw_blk, _ = gemm_tn(dzp, a_exp)                                               # [ldz, H in]
        row_head = _waug_row_head(H, d, dev)
        dW_aug = dw_blk.view(ldz, H, in_dim)[torch.arange(ldz, device=dev), row_head].contiguous()
        dW, dWf = torch.empty_like(W), torch.empty_like(Wf)
        dbf = torch.empty_like(bf) if bf is not None else None
        da, dT = torch.empty_like(a), torch.empty_like(T)
        _lib.check(lib.hsg_attn_prep_bwd(H, d, in_dim, Wf.shape[1], ldz, _p(W), _p(Wf), _p(bf), _p(a), _p(T),
                                         _p(dW_aug), _p(dq), _p(dW), _p(dWf), _p(dbf), _p(da), _p(dT), _st()))
        return None, None, None, None, dh, dW, dWf, dbf, da, dT, None, None


class S2SFn(torch.autograd.Function):
    """MultiHeadSGATLayer.forward(g, h) (GATStackLayer.py:36-44) -> cat of the heads' sh, or, with `origin`,
    elu(.) + origin (GAT.py:56-57 for layerType "S2S")."""

    @staticmethod
    def forward(ctx, batch, H, d, h, origin, W, a):
        _lib.require_device()
        lib = _lib.load()
        h, W, a = _f32c(h), _f32c(W), _f32c(a)
        origin = _f32c(origin) if origin is not None else None
        n, F = batch.n_super, H * d
        if h.shape[0] != n:
            raise ValueError("S2S: input has %d rows, graph has %d supernodes" % (h.shape[0], n))
        xgrp, xmember, mult = batch.s2s_groups()
        gc = _lib.S2SGraphC(batch.n_graphs, n, H, d, mult, 0, _p(batch.super_ptr), _p(batch.super_indptr),
                            _p(batch.super_extra), _p(xgrp), _p(xmember))
        z = gemm_nt(h, W)
        S = torch.empty(n, F, dtype=torch.float32, device=h.device)
        sh = torch.empty_like(S)
        x = torch.empty_like(S) if origin is not None else None
        _lib.check(lib.hsg_s2s_fwd(C.byref(gc), _p(z), _p(a), _p(origin), _p(S), _p(sh), _p(x), _st()))
        ctx.gc, ctx.keep, ctx.has_origin, ctx.dims = gc, (batch, xgrp, xmember), origin is not None, (H, d)
        ctx.save_for_backward(h, W, a, z, S)
        return x if origin is not None else sh

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        h, W, a, z, S = ctx.saved_tensors
        H, d = ctx.dims
        dout = _f32c(dout)
        dS, dz, da = torch.empty_like(S), torch.empty_like(z), torch.empty_like(a)
        ws = _Workspace.get(lib.hsg_s2s_bwd_workspace_bytes(ctx.gc.n_graphs, H, d), z.device, "s2s")
        _lib.check(lib.hsg_s2s_bwd(C.byref(ctx.gc), _p(z), _p(a), _p(S), _p(dout) if ctx.has_origin else None,
                                   None if ctx.has_origin else _p(dout), _p(dS), _p(dz), _p(da), 0, ws.data_ptr(),
                                   ws.numel(), _st()))
        dh = gemm_nn(dz, W)
        dW, _ = gemm_tn(dz, h)
        return None, None, None, dh, (dout if ctx.has_origin else None), dW, da


# --------------------------------------------------------------------------------------------
# position-wise FFN
# --------------------------------------------------------------------------------------------
# test hook: when set to a list, every FFN forward appends its ReLU active set (hdn > 0) - used by the parity
# tests to evaluate the oracle on the same active set (ReLU is the one discontinuous function on the path)
RELU_MASK_CAPTURE = None


def _ffn_forward(x, w1, b1, w2, b2, gamma, beta):
    lib = _lib.load()
    N, D = x.shape
    hdn = gemm_nt(x, w1, bias=b1, epi=EPI_BIAS | EPI_RELU)
    if RELU_MASK_CAPTURE is not None:
        RELU_MASK_CAPTURE.append((hdn > 0).cpu())
    r = gemm_nt(hdn, w2, bias=b2, R=x, epi=EPI_BIAS | EPI_ADD)
    out = torch.empty_like(r)
    stats = torch.empty(N, 2, dtype=torch.float32, device=x.device)
    _lib.check(lib.hsg_layernorm_fwd(N, D, _p(r), _p(gamma), _p(beta), _p(out), _p(stats), _st()))
    return hdn, r, stats, out


def _ffn_backward(dout, x, w1, w2, gamma, hdn, r, stats):
    """Returns (dx, dw1, db1, dw2, db2, dgamma, dbeta)."""
    lib = _lib.load()
    N, D = x.shape
    dev = x.device
    dr = torch.empty_like(r)
    dgamma = torch.empty_like(gamma)
    dbeta = torch.empty_like(gamma)
    ws = _Workspace.get(lib.hsg_layernorm_bwd_workspace_bytes(N, D), dev, "ln")
    _lib.check(lib.hsg_layernorm_bwd(N, D, _p(dout), _p(r), _p(stats), _p(gamma), _p(dr), _p(dgamma), _p(dbeta),
                                     _p(ws), ws.numel(), _st()))
    dhp = gemm_nn(dr, w2, R=hdn, epi=EPI_RELU_MASK)
    dw2, db2 = gemm_tn(dr, hdn, want_colsum=True)
    dw1, db1 = gemm_tn(dhp, x, want_colsum=True)
    dx = gemm_nn(dhp, w1, R=dr, epi=EPI_ADD)
    return dx, dw1, db1, dw2, db2, dgamma, dbeta


class FFNFn(torch.autograd.Function):
    """PositionwiseFeedForward.forward (GATLayer.py:35-44) on x [N, d_in] (dropout p = 0)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta):
        _lib.require_device()
        x, w1, b1, w2, b2, gamma, beta = (_f32c(t) for t in (x, w1, b1, w2, b2, gamma, beta))
        hdn, r, stats, out = _ffn_forward(x, w1, b1, w2, b2, gamma, beta)
        ctx.save_for_backward(x, w1, w2, gamma, hdn, r, stats)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w1, w2, gamma, hdn, r, stats = ctx.saved_tensors
        return _ffn_backward(_f32c(dout), x, w1, w2, gamma, hdn, r, stats)


def _keep_mult(n, p, seed, stream_id, device):
    """multipliers keep / (1 - p) of the library's counter-based mask (hsg_dropout.cu), element indices 0..n-1"""
    return dropout_keep_mask(n, p, seed, stream_id, device).to(torch.float32).mul_(1.0 / (1.0 - p))


class FFNDropFn(torch.autograd.Function):
    """PositionwiseFeedForward.forward in TRAINING mode with dropout p > 0 on its own (GATLayer.py:35-44:
    LayerNorm(x + dropout(w_2 relu(w_1 x)))).  Inside WSWGAT / the update loop the mask is applied by the kernels
    (hsg_update_loop_fwd); this stand-alone form uses the same mask generator (stream 1 = FFN of application 0,
    element row * F + col) and the same GEMM / LayerNorm entry points, with two elementwise torch ops for the mask."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta, p, seed):
        _lib.require_device()
        lib = _lib.load()
        x, w1, b1, w2, b2, gamma, beta = (_f32c(t) for t in (x, w1, b1, w2, b2, gamma, beta))
        N, D = x.shape
        hdn = gemm_nt(x, w1, bias=b1, epi=EPI_BIAS | EPI_RELU)
        y = gemm_nt(hdn, w2, bias=b2, epi=EPI_BIAS)
        mult = _keep_mult(N * D, p, seed, 1, x.device).view(N, D)
        r = torch.addcmul(x, y, mult)
        out = torch.empty_like(r)
        stats = torch.empty(N, 2, dtype=torch.float32, device=x.device)
        _lib.check(lib.hsg_layernorm_fwd(N, D, _p(r), _p(gamma), _p(beta), _p(out), _p(stats), _st()))
        ctx.save_for_backward(x, w1, w2, gamma, hdn, r, stats, mult)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x, w1, w2, gamma, hdn, r, stats, mult = ctx.saved_tensors
        N, D = x.shape
        dout = _f32c(dout)
        dr = torch.empty_like(r)
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
        ws = _Workspace.get(lib.hsg_layernorm_bwd_workspace_bytes(N, D), x.device, "ln")
        _lib.check(lib.hsg_layernorm_bwd(N, D, _p(dout), _p(r), _p(stats), _p(gamma), _p(dr), _p(dgamma), _p(dbeta),
                                         _p(ws), ws.numel(), _st()))
        drm = dr * mult                                    # the w_2 path sees the mask, the residual path does not
        dhp = gemm_nn(drm, w2, R=hdn, epi=EPI_RELU_MASK)
        dw2, db2 = gemm_tn(drm, hdn, want_colsum=True)
        dw1, db1 = gemm_tn(dhp, x, want_colsum=True)
        dx = gemm_nn(dhp, w1, R=dr, epi=EPI_ADD)
        return dx, dw1, db1, dw2, db2, dgamma, dbeta, None, None


# --------------------------------------------------------------------------------------------
# whole WSWGAT application: attention prep (once per layer) + one coarse-grained C call per direction
# --------------------------------------------------------------------------------------------
class AttnPrepFn(torch.autograd.Function):
    """(W, Wf, bf, a, T) -> (W_aug, q): attn_fc folded into the projection weight and the TF-IDF table.
    Depends on parameters only, so the update loop runs it once per layer and step; autograd sums the
    dW_aug / dq of every application before the single backward call."""

    @staticmethod
    def forward(ctx, H, d, W, Wf, bf, a, T):
        _lib.require_device()
        lib = _lib.load()
        W, Wf, a, T = (_f32c(t) for t in (W, Wf, a, T))
        bf = _f32c(bf) if bf is not None else None
        _, ldz = _lib.edge_layout(H, d)
        W_aug = torch.empty(ldz, W.shape[1], dtype=torch.float32, device=W.device)
        q = torch.empty(_N_BINS, H, dtype=torch.float32, device=W.device)
        _lib.check(lib.hsg_attn_prep_fwd(H, d, W.shape[1], Wf.shape[1], ldz, _p(W), _p(Wf), _p(bf), _p(a), _p(T),
                                         _p(W_aug), _p(q), _st()))
        ctx.H, ctx.d, ctx.has_bf = H, d, bf is not None
        ctx.save_for_backward(W, Wf, bf if bf is not None else W.new_empty(0), a, T)
        return W_aug, q

    @staticmethod
    def backward(ctx, dW_aug, dq):
        lib = _lib.load()
        W, Wf, bf, a, T = ctx.saved_tensors
        bf = bf if ctx.has_bf else None
        _, ldz = _lib.edge_layout(ctx.H, ctx.d)
        dW_aug, dq = _f32c(dW_aug), _f32c(dq)
        dW, dWf, da, dT = torch.empty_like(W), torch.empty_like(Wf), torch.empty_like(a), torch.empty_like(T)
        dbf = torch.empty_like(bf) if bf is not None else None
        _lib.check(lib.hsg_attn_prep_bwd(ctx.H, ctx.d, W.shape[1], Wf.shape[1], ldz, _p(W), _p(Wf), _p(bf), _p(a),
                                         _p(T), _p(dW_aug), _p(dq), _p(dW), _p(dWf), _p(dbf), _p(da), _p(dT), _st()))
        return None, None, dW, dWf, dbf, da, dT


_LAYOUT_CACHE = {}


def _layout(key, sizes):
    """cached (names, sizes, offsets, total) of an arena; every size is padded to a multiple of 4 floats."""
    lay = _LAYOUT_CACHE.get(key)
    if lay is None:
        names = [n for n, _ in sizes]
        padded = [(n + 3) & ~3 for _, n in sizes]
        offs, off = [], 0
        for n in padded:
            offs.append(off)
            off += n
        lay = (names, [n for _, n in sizes], padded, offs, max(off, 4))
        _LAYOUT_CACHE[key] = lay
    return lay


def _fill(struct, ints, ptrs):
    """bulk-fill a ctypes argument block: 8 leading int32 + pointer-sized fields (much cheaper than 30+ keyword
    conversions per call)."""
    n = len(ptrs)
    arr = (C.c_uint64 * (4 + n)).from_buffer(struct)
    arr[0] = (ints[0] & 0xFFFFFFFF) | (ints[1] << 32)
    arr[1] = (ints[2] & 0xFFFFFFFF) | (ints[3] << 32)
    arr[2] = (ints[4] & 0xFFFFFFFF) | (ints[5] << 32)
    arr[3] = (ints[6] & 0xFFFFFFFF) | (ints[7] << 32)
    arr[4:4 + n] = ptrs
    return struct


class WSWGATCoreFn(torch.autograd.Function):
    """out = FFN(elu(multi_head(g, neighbor)) + origin)   (GAT.py:56-58) given the prepared (W_aug, q)."""

    @staticmethod
    def forward(ctx, batch, kind, H, d, neighbor, origin, W_aug, q, w1, b1, w2, b2, gamma, beta):
        _lib.require_device()
        lib = _lib.load()
        csc, csc_t = batch.csc(kind)
        neighbor, origin, W_aug, q, w1, b1, w2, b2, gamma, beta = (
            _f32c(t) for t in (neighbor, origin, W_aug, q, w1, b1, w2, b2, gamma, beta))
        n_src, n_dst, F, in_dim, d_hid = csc.n_src, csc.n_dst, H * d, neighbor.shape[1], w1.shape[0]
        if neighbor.shape[0] != n_src or origin.shape[0] != n_dst:
            raise ValueError("%s: got %d neighbor / %d origin rows, graph has %d / %d" %
                             (kind, neighbor.shape[0], origin.shape[0], n_src, n_dst))
        if origin.shape[1] != F:
            raise ValueError("origin width %d != heads*head_dim %d" % (origin.shape[1], F))
        fp, ldz = _lib.edge_layout(H, d)
        _, _, _, offs, total = _layout(("f", H, d, n_src, n_dst, in_dim, d_hid), (
            ("out", n_dst * F), ("zp", n_src * ldz), ("sh", n_dst * F), ("x", n_dst * F), ("stat", n_dst * 3 * H),
            ("hdn", n_dst * d_hid), ("r", n_dst * F), ("ln", n_dst * 2)))
        arena = torch.empty(total, dtype=torch.float32, device=neighbor.device)
        base = arena.data_ptr()
        o_out, o_zp, o_sh, o_x, o_stat, o_hdn, o_r, o_ln = (base + 4 * o for o in offs)
        args = _fill(_lib.WswgatFwdArgsC(), (H, d, in_dim, d_hid, n_src, n_dst, ldz, 0),
                     [C.addressof(csc), neighbor.data_ptr(), origin.data_ptr(), W_aug.data_ptr(), q.data_ptr(),
                      w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                      o_zp, o_sh, o_x, o_stat, o_hdn, o_r, o_ln, o_out])
        _lib.check(lib.hsg_wswgat_fwd(C.byref(args), _st()))
        if RELU_MASK_CAPTURE is not None:
            RELU_MASK_CAPTURE.append((arena[offs[5]:offs[5] + n_dst * d_hid].view(n_dst, d_hid) > 0).cpu())
        ctx.batch, ctx.kind, ctx.H, ctx.d, ctx.dims = batch, kind, H, d, (n_src, n_dst, F, in_dim, d_hid, fp, ldz)
        ctx.ptr = (o_zp, o_sh, o_x, o_stat, o_hdn, o_r, o_ln)
        ctx.save_for_backward(neighbor, W_aug, q, w1, w2, gamma, arena)
        return arena[:n_dst * F].view(n_dst, F)

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        neighbor, W_aug, q, w1, w2, gamma, arena = ctx.saved_tensors
        n_src, n_dst, F, in_dim, d_hid, fp, ldz = ctx.dims
        H, d = ctx.H, ctx.d
        o_zp, o_sh, o_x, o_stat, o_hdn, o_r, o_ln = ctx.ptr
        _, csc_t = ctx.batch.csc(ctx.kind)
        dout = _f32c(dout)
        dev = dout.device
        # returned gradients first (sizes are multiples of 4 floats: one split_with_sizes gives all views), scratch after
        names, sizes, padded, offs, total = _layout(("b", H, d, n_src, n_dst, in_dim, d_hid), (
            ("d_neighbor", n_src * in_dim), ("dx", n_dst * F), ("dW_aug", ldz * in_dim), ("dq", _N_BINS * H),
            ("dw1", d_hid * F), ("db1", d_hid), ("dw2", F * d_hid), ("db2", F), ("dgamma", F), ("dbeta", F),
            ("dr", n_dst * F), ("dhp", n_dst * d_hid), ("g", n_dst * fp), ("dzp", n_src * ldz)))
        garena = torch.empty(total, dtype=torch.float32, device=dev)
        gb_ = garena.data_ptr()
        (p_dn, p_dx, p_dWa, p_dq, p_dw1, p_db1, p_dw2, p_db2, p_dg, p_db, p_dr, p_dhp, p_g, p_dzp) = (
            gb_ + 4 * o for o in offs)
        ws_bytes = lib.hsg_wswgat_bwd_workspace_bytes(H, d, in_dim, d_hid, n_src, n_dst)
        ws = _Workspace.get(ws_bytes, dev, "wswgat")
        args = _fill(_lib.WswgatBwdArgsC(), (H, d, in_dim, d_hid, n_src, n_dst, ldz, 0),
                     [C.addressof(csc_t), dout.data_ptr(), neighbor.data_ptr(), W_aug.data_ptr(), q.data_ptr(),
                      w1.data_ptr(), w2.data_ptr(), gamma.data_ptr(), o_zp, o_sh, o_x, o_hdn, o_r, o_ln, o_stat, p_dr,
                      p_dhp, p_g, p_dzp, p_dx, p_dn, p_dWa, p_dq, p_dw1, p_db1, p_dw2, p_db2, p_dg, p_db,
                      ws.data_ptr(), ws.numel()])
        _lib.check(lib.hsg_wswgat_bwd(C.byref(args), _st()))
        n_ret = 10
        if padded[:n_ret] == sizes[:n_ret]:
            parts = garena[:offs[n_ret]].split_with_sizes(sizes[:n_ret])
        else:
            parts = [garena[offs[i]:offs[i] + sizes[i]] for i in range(n_ret)]
        dn, dx, dWa, dq, dw1, db1, dw2, db2, dg, db = parts
        return (None, None, None, None, dn.view(n_src, in_dim), dx.view(n_dst, F), dWa.view(ldz, in_dim),
                dq.view(_N_BINS, H), dw1.view(d_hid, F), db1, dw2.view(F, d_hid), db2, dg, db)


# --------------------------------------------------------------------------------------------
# whole update loop: ONE C call forward, ONE backward (hsg_update_loop_fwd / _bwd)
# --------------------------------------------------------------------------------------------
_PARAM_ORDER = ("W", "Wf", "bf", "a", "w1", "b1", "w2", "b2", "gamma", "beta")


def _layer_params_c(H, d, in_dim, feat_dim, d_hid, tensors):
    return _lib.LayerParamsC(H, d, in_dim, feat_dim, d_hid, 0, *[_p(t) for t in tensors])


class UpdateLoopFn(torch.autograd.Function):
    """(word_state, super_state) of a chain of WSWGAT applications whose kinds alternate from `start_kind` (0 = W2S,
    1 = S2W) on a HeteroBatch: the update loop of HSumGraph.forward / HSumDocGraph.forward (HiGraph.py:98-106,
    205-214) is (n_apps = 1 + 2 n_iter, start 0); one stand-alone WSWGAT.forward (GAT.py:45-59) is (1, kind).
    Tensor arguments after `cfg`: word_feature, super_feature, T, then the ten packed parameters of word2sent and of
    sent2word (W, Wf, bf, a, w1, b1, w2, b2, gamma, beta; word2sent's bf is None).

    cfg = dict(n_apps, start_kind, w2s=(H, d, d_hid), s2w=(H, d, d_hid), grad_targets, attn_p, ffn_p, seed):
    grad_targets is None (gradients are returned to autograd) or a list of 21 tensors [dT, 10 x W2S, 10 x S2W] that
    the backward ADDS the parameter gradients into (fused accumulation into .grad; None is returned for those
    inputs).  attn_p / ffn_p > 0 enable training-mode dropout with masks drawn from `seed` (hsg_dropout.cu)."""

    @staticmethod
    def forward(ctx, batch, cfg, word_feature, super_feature, T, *params):
        _lib.require_device()
        lib = _lib.load()
        n_apps, start = cfg["n_apps"], cfg["start_kind"]
        (H1, d1, hid1), (H2, d2, hid2) = cfg["w2s"], cfg["s2w"]
        word_feature, super_feature = _f32c(word_feature), _f32c(super_feature)
        n_word, n_super = batch.n_word, batch.n_super
        if word_feature.shape[0] != n_word or super_feature.shape[0] != n_super:
            raise ValueError("update loop: got %d word / %d supernode rows, graph has %d / %d" %
                             (word_feature.shape[0], super_feature.shape[0], n_word, n_super))
        Dw, Ds = word_feature.shape[1], super_feature.shape[1]
        # the argument block of the parameters is cached on the caller (cfg["cache"]) and reused while every
        # parameter tensor keeps its storage: ~40 ctypes field conversions per call otherwise
        cache = cfg.get("cache")
        sig = (T.data_ptr(), Dw, Ds) + tuple(0 if t is None else t.data_ptr() for t in params)
        ent = cache.get("fwd") if cache is not None else None
        if ent is not None and ent[0] == sig:
            args = type(ent[1]).from_buffer_copy(ent[1])
        else:
            T = _f32c(T)
            params = tuple(_f32c(t) if t is not None else None for t in params)
            fe = T.shape[1]
            uses = [start == 0 or n_apps > 1, start == 1 or n_apps > 1]
            if (uses[0] and Ds != H1 * d1) or (uses[1] and Dw != H2 * d2):
                raise ValueError("update loop: feature widths (%d, %d) do not match heads*head_dim" % (Dw, Ds))
            args = _lib.LoopArgsC(n_apps, start, 0, 0, None, None,
                                  _layer_params_c(H1, d1, Dw, fe, hid1, params[:10]),
                                  _layer_params_c(H2, d2, Ds, fe, hid2, params[10:]),
                                  _p(T), None, None, None, 0, 0.0, 0.0, 0)
            if cache is not None and all(t is None or t.is_contiguous() for t in params) and T.is_contiguous():
                cache["fwd"] = (sig, type(args).from_buffer_copy(args))
        csc_s, csc_w = batch.csc("W2S")
        args.n_apps, args.start_kind, args.n_word, args.n_super = n_apps, start, n_word, n_super
        args.csc_super, args.csc_word = C.pointer(csc_s), C.pointer(csc_w)
        args.word_feature, args.super_feature = word_feature.data_ptr(), super_feature.data_ptr()
        args.attn_p, args.ffn_p, args.seed = float(cfg.get("attn_p", 0.0)), float(cfg.get("ffn_p", 0.0)), int(cfg.get("seed", 0))
        sd = cfg.get("seed_dev")                    # device step counter mixed into the dropout key (CUDA-graph replay)
        args.seed_dev = sd.data_ptr() if sd is not None else None
        ev = cfg.get("input_ready")                 # torch.cuda.Event recorded on the stream that produces the inputs
        args.input_ready = ev.cuda_event if ev is not None else None
        pkey = (n_apps, start, n_word, n_super, args.attn_p > 0, args.ffn_p > 0, sig[1:])
        pent = cache.get("plan") if cache is not None else None
        if pent is not None and pent[0] == pkey:
            plan = pent[1]
        else:
            plan = _lib.LoopPlanC()
            _lib.check(lib.hsg_update_loop_plan(C.byref(args), C.byref(plan)))
            if cache is not None:
                cache["plan"] = (pkey, plan)
        state = torch.empty(plan.state_floats, dtype=torch.float32, device=word_feature.device)
        args.state, args.state_floats = state.data_ptr(), plan.state_floats
        _lib.check(lib.hsg_update_loop_fwd(C.byref(args), _st()))
        if RELU_MASK_CAPTURE is not None:
            for i in range(n_apps):
                n_rows, hid = (n_super, hid1) if (i + start) % 2 == 0 else (n_word, hid2)
                off = plan.hdn_off[i % 2] + (i // 2) * plan.pair_stride
                RELU_MASK_CAPTURE.append((state[off:off + n_rows * hid].view(n_rows, hid) > 0).cpu())
        none = (1 << 64) - 1
        so, wo = plan.super_state_off, plan.word_state_off
        super_state = state[so:so + n_super * Ds].view(n_super, Ds) if so != none else super_feature.clone()
        word_state = state[wo:wo + n_word * Dw].view(n_word, Dw) if wo != none else word_feature.clone()
        ctx.batch, ctx.cfg, ctx.plan, ctx.args = batch, cfg, plan, args
        ctx.set_materialize_grads(False)          # an unused result's gradient arrives as None, not as a zero tensor
        ctx.save_for_backward(word_feature, super_feature, T, state, *[t for t in params if t is not None])
        ctx.param_present = [t is not None for t in params]
        return word_state, super_state

    @staticmethod
    def backward(ctx, d_word, d_super):
        lib = _lib.load()
        cfg = ctx.cfg
        grad_targets = cfg.get("grad_targets")
        saved = ctx.saved_tensors
        word_feature, super_feature, T, state = saved[:4]
        n_params = len(ctx.param_present)
        dev = state.device
        plan, args = ctx.plan, ctx.args
        if d_word is None and d_super is None:
            return (None,) * (5 + n_params)
        d_word = _f32c(d_word) if d_word is not None else None
        d_super = _f32c(d_super) if d_super is not None else None
        d_wf = torch.empty_like(word_feature) if ctx.needs_input_grad[2] else None
        d_sf = torch.empty_like(super_feature) if ctx.needs_input_grad[3] else None
        n_apps, start = cfg["n_apps"], cfg["start_kind"]
        if grad_targets is not None:
            targets = grad_targets
            acc = 1
        else:
            it = iter(saved[4:])
            params = [next(it) if present else None for present in ctx.param_present]
            uses = [start == 0 or n_apps > 1, start == 1 or n_apps > 1]
            targets = [torch.empty_like(T)] + [torch.empty_like(p) if p is not None else None for p in params]
            for k in (0, 1):                                 # a layer no application used: zero gradients
                if not uses[k]:
                    for i in range(1 + 10 * k, 11 + 10 * k):
                        if targets[i] is not None:
                            targets[i].zero_()
            acc = 0
        scratch = torch.empty(plan.scratch_floats, dtype=torch.float32, device=dev)
        ws = _Workspace.get(plan.ws_bytes, dev, "loop")
        cache = cfg.get("cache") if grad_targets is not None else None
        tsig = tuple(0 if t is None else t.data_ptr() for t in targets) if cache is not None else None
        ent = cache.get("bwd") if cache is not None else None
        if ent is not None and ent[0] == tsig:
            b = type(ent[1]).from_buffer_copy(ent[1])
        else:
            b = _lib.LoopBwdArgsC(None, None, None, None,
                                  _lib.LayerGradsC(*[_p(t) for t in targets[1:11]]),
                                  _lib.LayerGradsC(*[_p(t) for t in targets[11:21]]),
                                  _p(targets[0]), acc, 0, None, 0, None, 0)
            if cache is not None:
                cache["bwd"] = (tsig, type(b).from_buffer_copy(b))
        b.d_word_state, b.d_super_state, b.d_word_feature, b.d_super_feature = _p(d_word), _p(d_super), _p(d_wf), _p(d_sf)
        b.accumulate = acc
        b.scratch, b.scratch_floats, b.ws, b.ws_bytes = scratch.data_ptr(), plan.scratch_floats, ws.data_ptr(), ws.numel()
        _lib.check(lib.hsg_update_loop_bwd(C.byref(args), C.byref(b), _st()))
        if grad_targets is not None:
            return (None, None, d_wf, d_sf) + (None,) * (1 + n_params)
        return (None, None, d_wf, d_sf, targets[0]) + tuple(targets[1:])


def dropout_keep_mask(n, p, seed, stream_id, device="cuda"):
    """Test hook: the keep flags (uint8) of the first n element indices of the mask (p, seed, stream_id)."""
    out = torch.empty(n, dtype=torch.uint8, device=device)
    _lib.check(_lib.load().hsg_dropout_mask(n, float(p), int(seed), int(stream_id), out.data_ptr(), _st()))
    return out


# --------------------------------------------------------------------------------------------
# readout + loss (hsg_head_fwd / _bwd), top-m extraction, fused Adam
# --------------------------------------------------------------------------------------------
class SentenceLossFn(torch.autograd.Function):
    """(loss, logits) = mean over graphs of the per-graph sum of sentence cross-entropies of wh(state)
    (HiGraph.py:108 / :216-228 + train.py:114-119).  `logits` is returned for extraction (not differentiable here:
    the reference only back-propagates the loss)."""

    @staticmethod
    def forward(ctx, batch, n_graphs_global, grad_targets, state, wh_w, wh_b, labels):
        _lib.require_device()
        lib = _lib.load()
        state, wh_w, wh_b = _f32c(state), _f32c(wh_w), _f32c(wh_b)
        n_super, hidden = state.shape
        two_part = 1 if wh_w.shape[1] == 2 * hidden else 0
        if wh_w.shape[0] != 2 or wh_w.shape[1] != hidden * (1 + two_part):
            raise ValueError("wh weight must be [2, hidden] (HSG) or [2, 2*hidden] (HDSG)")
        n_sent = labels.shape[0]
        labels = labels.contiguous()
        if labels.dtype != torch.int64:
            raise TypeError("labels must be int64")
        sent_row = batch.sent_row
        if sent_row is None and n_sent != n_super:
            raise ValueError("HSG batch: %d labels for %d sentence nodes" % (n_sent, n_super))
        doc_row = batch.sent_doc_row.int() if two_part else None
        gptr = batch.graph_sent_ptr
        args = _lib.HeadArgsC(n_sent, n_super, hidden, two_part, batch.n_graphs, 0, _p(state), _p(sent_row), _p(doc_row),
                              _p(gptr), _p(wh_w), _p(wh_b), _p(labels), 1.0 / float(n_graphs_global), 0.0)
        dev = state.device
        out = torch.empty(4 * n_sent + 4, dtype=torch.float32, device=dev)
        logits, dlogits, loss = out[:2 * n_sent].view(n_sent, 2), out[2 * n_sent:4 * n_sent], out[4 * n_sent:4 * n_sent + 1]
        ws_bytes = lib.hsg_head_workspace_bytes(n_sent, hidden * (1 + two_part))
        ws = _Workspace.get(ws_bytes, dev, "head")
        _lib.check(lib.hsg_head_fwd(C.byref(args), _p(logits), _p(dlogits), _p(loss), ws.data_ptr(), ws.numel(), _st()))
        ctx.args, ctx.keep, ctx.grad_targets = args, (sent_row, doc_row, gptr, labels), grad_targets
        ctx.save_for_backward(state, wh_w, wh_b, dlogits)
        ctx.mark_non_differentiable(logits)
        return loss.view(()), logits

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        lib = _lib.load()
        state, wh_w, wh_b, dlogits = ctx.saved_tensors
        dev = state.device
        d_state = torch.empty_like(state)
        if ctx.grad_targets is not None:
            d_w, d_b = ctx.grad_targets
            acc = 1
        else:
            d_w, d_b = torch.empty_like(wh_w), torch.empty_like(wh_b)
            acc = 0
        gout = _f32c(dloss) if dloss is not None else None
        ws = _Workspace.get(lib.hsg_head_workspace_bytes(ctx.args.n_sent, wh_w.shape[1]), dev, "head")
        _lib.check(lib.hsg_head_bwd(C.byref(ctx.args), _p(dlogits), _p(gout), _p(d_state), _p(d_w), _p(d_b), acc,
                                    ws.data_ptr(), ws.numel(), _st()))
        if ctx.grad_targets is not None:
            return None, None, None, d_state, None, None, None
        return None, None, None, d_state, d_w, d_b, None


def head_fwd_bwd(batch, n_graphs_global, grad_targets, state, wh_w, wh_b, labels):
    """(loss, logits, d_state) of SentenceLossFn forward + backward with d L / d loss = 1 in ONE launch
    (hsg_head_fwd_bwd; train.py:114-121): bit-identical to the two calls, d wh ADDED into grad_targets = (d_w, d_b)."""
    _lib.require_device()
    lib = _lib.load()
    state, wh_w, wh_b = _f32c(state), _f32c(wh_w), _f32c(wh_b)
    n_super, hidden = state.shape
    two_part = 1 if wh_w.shape[1] == 2 * hidden else 0
    if wh_w.shape[0] != 2 or wh_w.shape[1] != hidden * (1 + two_part):
        raise ValueError("wh weight must be [2, hidden] (HSG) or [2, 2*hidden] (HDSG)")
    n_sent = labels.shape[0]
    labels = labels.contiguous()
    if labels.dtype != torch.int64:
        raise TypeError("labels must be int64")
    sent_row = batch.sent_row
    if sent_row is None and n_sent != n_super:
        raise ValueError("HSG batch: %d labels for %d sentence nodes" % (n_sent, n_super))
    doc_row = batch.sent_doc_row.int() if two_part else None
    gptr = batch.graph_sent_ptr
    args = _lib.HeadArgsC(n_sent, n_super, hidden, two_part, batch.n_graphs, 0, _p(state), _p(sent_row), _p(doc_row),
                          _p(gptr), _p(wh_w), _p(wh_b), _p(labels), 1.0 / float(n_graphs_global), 0.0)
    dev = state.device
    out = torch.empty(4 * n_sent + 4, dtype=torch.float32, device=dev)
    logits, dlogits, loss = out[:2 * n_sent].view(n_sent, 2), out[2 * n_sent:4 * n_sent], out[4 * n_sent:4 * n_sent + 1]
    d_state = torch.empty_like(state)
    d_w, d_b = grad_targets
    ws = _Workspace.get(lib.hsg_head_workspace_bytes(n_sent, hidden * (1 + two_part)), dev, "head")
    _lib.check(lib.hsg_head_fwd_bwd(C.byref(args), _p(logits), _p(dlogits), _p(loss), _p(d_state), _p(d_w), _p(d_b), 1,
                                    ws.data_ptr(), ws.numel(), _st()))
    return loss.view(()), logits, d_state


class DocInitFn(torch.autograd.Function):
    """supernode init features of an HDSG batch (HSumDocGraph.forward + set_dnfeature, HiGraph.py:196-203,231-244):
    sentence rows = sent_feature, document rows = dn_feature_proj(mean of the document's sentence features)."""

    @staticmethod
    def forward(ctx, batch, sent_feature, W):
        _lib.require_device()
        lib = _lib.load()
        sent_feature, W = _f32c(sent_feature), _f32c(W)
        n_sent, hid = sent_feature.shape
        if batch.doc_row is None or batch.sent_doc_gidx is None:
            raise ValueError("DocInitFn needs an HDSG batch built by the device builder")
        doc_row32, sent_row32 = batch.doc_row.int(), batch.sent_row
        n_doc = doc_row32.shape[0]
        dm = _lib.DocMapC(n_sent, n_doc, hid, 0, _p(sent_row32), _p(doc_row32), _p(batch.sent_doc_gidx),
                          _p(batch.doc_graph), _p(batch.graph_sent_ptr))
        dev = sent_feature.device
        doc_mean = torch.empty(n_doc, hid, dtype=torch.float32, device=dev)
        _lib.check(lib.hsg_doc_mean(C.byref(dm), _p(sent_feature), _p(doc_mean), _st()))
        doc_feature = gemm_nt(doc_mean, W)
        sup = torch.empty(batch.n_super, hid, dtype=torch.float32, device=dev)
        _lib.check(lib.hsg_super_assemble(C.byref(dm), _p(sent_feature), _p(doc_feature), _p(sup), _st()))
        ctx.dm, ctx.keep = dm, (batch, doc_row32, sent_row32)
        ctx.save_for_backward(doc_mean, W)
        return sup

    @staticmethod
    def backward(ctx, d_sup):
        lib = _lib.load()
        doc_mean, W = ctx.saved_tensors
        d_sup = _f32c(d_sup)
        dm = ctx.dm
        d_doc_feature = torch.empty_like(doc_mean)
        _lib.check(lib.hsg_doc_init_bwd(C.byref(dm), _p(d_sup), None, _p(d_doc_feature), None, _st()))
        d_doc_mean = gemm_nn(d_doc_feature, W)
        dW, _ = gemm_tn(d_doc_feature, doc_mean)
        d_sent = torch.empty(dm.n_sent, dm.hidden, dtype=torch.float32, device=d_sup.device)
        _lib.check(lib.hsg_doc_init_bwd(C.byref(dm), _p(d_sup), _p(d_doc_mean), None, _p(d_sent), _st()))
        return None, d_sent, dW


def topm(logits, graph_sent_ptr, m):
    """[n_graphs, m] int32: per graph the local indices of the m sentences with the largest class-1 logit, in
    descending order, -1 padded (Tester.py:128 torch.topk(p_sent[:, 1], min(m, N)))."""
    _lib.require_device()
    lib = _lib.load()
    logits = _f32c(logits)
    n_graphs = graph_sent_ptr.shape[0] - 1
    out = torch.empty(n_graphs, m, dtype=torch.int32, device=logits.device)
    _lib.check(lib.hsg_topm(_p(logits), _p(graph_sent_ptr), n_graphs, m, _p(out), _st()))
    return out


class FusedAdam:
    """torch.optim.Adam(lr, betas, eps) (train.py:90) over ONE flat fp32 parameter arena and its flat gradient
    (dist.FlatGradArena), one kernel per step; max_grad_norm > 0 folds clip_grad_norm_ (train.py:132-133) in."""

    def __init__(self, flat_param, flat_grad, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=0.0):
        self.p, self.g = flat_param, flat_grad
        self.m, self.v = torch.zeros_like(flat_param), torch.zeros_like(flat_param)
        self.lr, self.betas, self.eps, self.max_grad_norm, self.t = lr, betas, eps, float(max_grad_norm), 0
        self.ws = torch.empty(_lib.load().hsg_adam_workspace_bytes(), dtype=torch.uint8, device=flat_param.device)
        self.step_state = None          # device-resident step counter (step_dev), created on first use

    def step(self):
        if self.step_state is not None:
            raise RuntimeError("FusedAdam: the step count lives on the device (step_dev was used); keep using step_dev")
        self.t += 1
        _lib.check(_lib.load().hsg_adam_step(self.p.numel(), _p(self.p), _p(self.g), _p(self.m), _p(self.v), self.lr,
                                             self.betas[0], self.betas[1], self.eps, self.t, self.max_grad_norm,
                                             self.ws.data_ptr(), self.ws.numel(), _st()))

    def device_step_counter(self):
        """[4] int64 on the device: [0] = completed steps (starts at the host count), [1..3] = kernel-internal tickets."""
        if self.step_state is None:
            self.step_state = torch.tensor([self.t, 0, 0, 0], dtype=torch.int64, device=self.p.device)
        return self.step_state

    def step_dev(self, zero_grad=True):
        """The same update with the step number read from (and advanced in) device memory: identical kernel arguments
        every step, so the launch can be replayed from a CUDA graph.  zero_grad clears the gradient arena after use."""
        st = self.device_step_counter()
        _lib.check(_lib.load().hsg_adam_step_dev(self.p.numel(), _p(self.p), _p(self.g), _p(self.m), _p(self.v), self.lr,
                                                 self.betas[0], self.betas[1], self.eps, st.data_ptr(),
                                                 1 if zero_grad else 0, self.max_grad_norm, self.ws.data_ptr(),
                                                 self.ws.numel(), _st()))


def embed_gather(ids, table, out=None):
    """table[ids] for int32 ids (the frozen word-embedding lookup, HiGraph.py:147-148) in one library kernel."""
    _lib.require_device()
    n, dim = ids.shape[0], table.shape[1]
    if ids.dtype != torch.int32:
        raise TypeError("embed_gather expects int32 ids")
    if out is None:
        out = torch.empty(n, dim, dtype=torch.float32, device=table.device)
    _lib.check(_lib.load().hsg_embed_gather(n, dim, _p(ids), _p(_f32c(table)), _p(out), _st()))
    return out
