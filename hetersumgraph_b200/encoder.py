"""Sentence encoder in front of the WSWGAT path, B200-native.

Mirrors the sentence-node initialisation of the reference:

  sentEncoder (n-gram CNN)        module/Encoder.py:18-76
  _init_sn_param                  HiGraph.py:112-125     (parameter names / shapes: same state_dict keys)
  _sent_cnn_feature               HiGraph.py:127-133
  _sent_lstm_feature              HiGraph.py:135-142  +  get_snode_feat HiGraph.py:247-255
  set_snfeature + n_feature_proj  HiGraph.py:154-161, :96

What runs where
  * n-gram CNN: csrc/hsg_encoder.cu + the tcgen05 GEMM (compact rows, six convolutions as one product, segmented
    max, sparse weight gradient) - `NgramEncodeFn`.  The reference's per-sentence host loop with a device sync per
    sentence (Encoder.py:61-66) becomes an `EncoderPlan` computed once from the host token matrix at batch time.
  * cnn_proj / lstm_proj / n_feature_proj and the position-embedding add: the library's GEMMs, one autograd Function
    (`SentHeadFn`), the concatenation of HiGraph.py:160 is just the column layout of one buffer.
  * BiLSTM: csrc/hsg_lstm.cu - one persistent kernel per layer and direction set (CTA per graph x direction, W_hh
    on chip for the whole sequence) + the library's GEMMs for the input products and all weight gradients -
    `LstmFn`.  Every graph's sentences are contiguous rows, so the pad / pack / unpack / per-graph Python loops of
    HiGraph.py:136-141,247-255 disappear.  `nn.LSTM` is kept as the PARAMETER CONTAINER (same state_dict keys and
    initialisation as the reference); its own forward (cuDNN) is never called by the package - the comparison with
    torch's LSTM on a PackedSequence (the call the reference makes) lives in tests/test_gpu_encoder.py.
No CPU fallback: every custom op raises without the sm_100a library.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .functional import _Workspace, _f32c, _p, _st, gemm_nn, gemm_nt, gemm_tn

N_CHANNELS = 50
KERNEL_HEIGHTS = (2, 3, 4, 5, 6, 7)
NGRAM_DIM = N_CHANNELS * len(KERNEL_HEIGHTS)
H_MAX = 7
TAIL_ROWS = 8
CH_PAD = 52                              # channels per height in the product's column layout (csrc/hsg_encoder.cu)
Y_COLS = CH_PAD * len(KERNEL_HEIGHTS)
# K-chunks of the convolution product: (first kernel row, rows, first height group that reaches them)
CONV_CHUNKS = ((0, 2, 0), (2, 2, 1), (4, 2, 3), (6, 1, 5))


def sinusoid_table(n_position, d_hid, padding_idx=None):
    """Same values as module/PositionEmbedding.py:21-40 (float64 table cast to float32), vectorised."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    tab = pos / np.power(10000, 2 * (j // 2) / d_hid)
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    if padding_idx is not None:
        tab[padding_idx] = 0.0
    return torch.FloatTensor(tab)


class EncoderPlan:
    """Host-side plan of one batch for the encoder kernels (computed from the HOST token matrix, no device sync):
    per-sentence length (number of non-zero ids, Encoder.py:58), compact-row offsets, sentence position inside its
    graph (dataloader.py:241) and the time-major permutation of a PackedSequence over the per-graph sentence lists
    (HiGraph.py:136-137; graphs are in batch order = #sentences descending, dataloader.py:479)."""

    def __init__(self, tokens, graph_sent_ptr, device="cuda", tokens_dev=None):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        ptr = np.asarray(graph_sent_ptr, dtype=np.int64)
        S, L = tokens.shape
        if L < H_MAX:
            raise ValueError("sent_max_len must be >= 7 (the largest convolution kernel, Encoder.py:37)")
        if int(ptr[-1]) != S:
            raise ValueError("graph_sent_ptr does not cover the token matrix")
        B = len(ptr) - 1
        # one int32 host blob [sent_len | row_ptr | sent_pos | graph_sent_ptr], filled by the library's host routine
        # (one pass over the token matrix in C instead of several numpy passes) and uploaded with ONE copy
        blob = np.empty(3 * S + 1 + B + 1, np.int32)
        gptr32 = blob[3 * S + 1:]
        gptr32[:] = ptr
        base = blob.ctypes.data
        _lib.check(_lib.load().hsg_enc_plan_host(S, L, tokens.ctypes.data, B, gptr32.ctypes.data, base, base + 4 * S,
                                                  base + 4 * (2 * S + 1)))
        counts = np.diff(ptr)
        self.n_sent, self.L, self.n_rows = S, L, int(blob[2 * S])
        self._ptr, self._counts, self._device, self._packed = ptr, counts, device, None
        dev = torch.from_numpy(blob).to(device, non_blocking=True)
        self.sent_len, self.row_ptr, self.sent_pos = dev[:S], dev[S:2 * S + 1], dev[2 * S + 1:3 * S + 1]
        self.graph_sent_ptr = dev[3 * S + 1:]
        self.n_graphs = len(ptr) - 1
        self.tokens = tokens_dev if tokens_dev is not None else torch.from_numpy(tokens).to(device, non_blocking=True)

    def _packed_order(self):
        """(batch_sizes on the host, perm, inv_perm): the time-major order of a PackedSequence over the per-graph sentence
        lists - only the comparison with torch's own LSTM in the tests needs it, so it is built on first use."""
        if self._packed is None:
            ptr, counts, S = self._ptr, self._counts, self.n_sent
            if np.any(counts[1:] > counts[:-1]):
                raise ValueError("a PackedSequence needs the graphs in batch order (#sentences descending, "
                                 "dataloader.py:479)")
            t_max = int(counts[0]) if len(counts) else 0
            batch_sizes = (counts[None, :] > np.arange(t_max)[:, None]).sum(axis=1).astype(np.int64)
            perm = np.concatenate([ptr[:b] + t for t, b in enumerate(batch_sizes)]) if t_max else np.zeros(0, np.int64)
            inv = np.empty_like(perm)
            inv[perm] = np.arange(len(perm))
            both = torch.from_numpy(np.concatenate([perm, inv])).to(self._device)
            self._packed = (torch.from_numpy(batch_sizes), both[:S], both[S:])
        return self._packed

    batch_sizes = property(lambda self: self._packed_order()[0])
    perm = property(lambda self: self._packed_order()[1])
    inv_perm = property(lambda self: self._packed_order()[2])

    @staticmethod
    def from_token_batch(tb, device="cuda", tokens_dev=None):
        return EncoderPlan(tb.tokens, tb.graph_sent_ptr, device, tokens_dev)


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class NgramEncodeFn(torch.autograd.Function):
    """ngram [S, 300] = sentEncoder.forward(words) (Encoder.py:56-76).  Inputs: plan, embedding table, position table,
    then (weight, bias) of the six Conv2d modules.  Gradients for the convolution parameters only (frozen embedding and
    position table, train.py:340-342 / Encoder.py:44-45)."""

    @staticmethod
    def forward(ctx, plan, targets, embed_w, pos_table, *conv):
        """targets: None, or the twelve existing .grad buffers (w0, b0, ..., w5, b5) the backward ADDS into."""
        _lib.require_device()
        lib = _lib.load()
        ctx.targets = targets
        ws_, bs_ = [_f32c(t) for t in conv[0::2]], [_f32c(t) for t in conv[1::2]]
        embed_w, pos_table = _f32c(embed_w), _f32c(pos_table)
        D = embed_w.shape[1]
        S, R = plan.n_sent, plan.n_rows
        dev = embed_w.device
        xc = torch.empty(R + TAIL_ROWS, D, dtype=torch.float32, device=dev)
        _lib.check(lib.hsg_enc_gather(S, plan.L, D, R, _p(plan.tokens), _p(plan.sent_len), _p(plan.row_ptr), _p(embed_w),
                                      _p(pos_table), _p(xc), _st()))
        K = H_MAX * D
        wpad = torch.empty(Y_COLS, K, dtype=torch.float32, device=dev)
        _lib.check(lib.hsg_enc_pack_weights(D, _ptr_array(ws_), _p(wpad), _st()))
        y = torch.empty(max(R, 1), Y_COLS, dtype=torch.float32, device=dev)
        # the six convolutions: A = xc with row pitch D (overlapping rows), K-chunks accumulated in place
        for j0, nj, g0 in CONV_CHUNKS:
            c0 = g0 * CH_PAD
            a_ptr = xc.data_ptr() + 4 * j0 * D
            b_ptr = wpad.data_ptr() + 4 * (c0 * K + j0 * D)
            y_ptr = y.data_ptr() + 4 * c0
            _lib.check(lib.hsg_gemm_nt(R, Y_COLS - c0, nj * D, a_ptr, D, b_ptr, K, y_ptr, Y_COLS, None,
                                       y_ptr if j0 else None, Y_COLS, _lib.EPI_ADD if j0 else 0, _st()))
        out = torch.empty(S, NGRAM_DIM, dtype=torch.float32, device=dev)
        arg_t = torch.empty(NGRAM_DIM, max(S, 1), dtype=torch.int32, device=dev)
        _lib.check(lib.hsg_enc_pool_fwd(S, _p(plan.row_ptr), _p(y), Y_COLS, _ptr_array(bs_), _p(out), NGRAM_DIM,
                                        _p(arg_t), _st()))
        ctx.save_for_backward(xc, arg_t)
        ctx.shapes = [tuple(t.shape) for t in conv]
        ctx.S, ctx.D = S, D
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        xc, arg_t = ctx.saved_tensors
        d_out = _f32c(d_out)
        dev = d_out.device
        fused = ctx.targets is not None
        if fused:
            dws, dbs = list(ctx.targets[0::2]), list(ctx.targets[1::2])
        else:
            dws = [torch.empty(s, dtype=torch.float32, device=dev) for s in ctx.shapes[0::2]]
            dbs = [torch.empty(s, dtype=torch.float32, device=dev) for s in ctx.shapes[1::2]]
        nbytes = lib.hsg_enc_conv_wgrad_workspace_bytes(ctx.S, ctx.D)
        ws = _Workspace.get(nbytes, dev, "enc_wgrad")
        _lib.check(lib.hsg_enc_conv_wgrad(ctx.S, ctx.D, _p(xc), _p(d_out), d_out.stride(0), _p(arg_t), _ptr_array(dws),
                                          _ptr_array(dbs), 1 if fused else 0, _p(ws), ws.numel(), _st()))
        grads = [None, None, None, None]
        for w, b in zip(dws, dbs):
            grads += [None, None] if fused else [w, b]
        return tuple(grads)


_SIDE_STREAMS = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class LstmFn(torch.autograd.Function):
    """out [S, ndir * H] = the last layer's output of nn.LSTM run on every graph's sentence rows as one sequence
    (pack_padded_sequence semantics, HiGraph.py:135-141).  Arguments: x [S, in], graph_sent_ptr (device int32 [B+1]),
    (n_graphs, H, n_layers, ndir, p, seed), then nn.LSTM's parameters in its own order (per layer, per direction:
    weight_ih, weight_hh, bias_ih, bias_hh).  p > 0: inter-layer dropout (nn.LSTM(dropout=0.1), HiGraph.py:118) with
    the library's counter-based masks (hsg_dropout_mask, stream 1000 + layer), applied by two elementwise torch ops -
    training-mode only, off in every parity test and in both bench arms."""

    overlap_weight_grads = True       # class-level switch (tests compare both settings)

    @staticmethod
    def forward(ctx, x, gptr, cfg, *params):
        _lib.require_device()
        lib = _lib.load()
        n_graphs, H, n_layers, ndir, p_drop, seed = cfg[:6]
        ctx.targets = cfg[6] if len(cfg) > 6 else None       # existing .grad buffers in nn.LSTM's parameter order
        x = _f32c(x)
        params = [_f32c(p) for p in params]
        S = x.shape[0]
        dev = x.device
        saved, inp, scales = [], x, []
        for layer in range(n_layers):
            pl = [params[(layer * ndir + d) * 4:(layer * ndir + d) * 4 + 4] for d in range(ndir)]
            xproj = torch.empty(S, ndir * 4 * H, dtype=torch.float32, device=dev)
            for d in range(ndir):
                gemm_nt(inp, pl[d][0], out=xproj[:, d * 4 * H:(d + 1) * 4 * H])
            out = torch.empty(S, ndir * H, dtype=torch.float32, device=dev)
            gates = torch.empty(S, ndir, 4 * H, dtype=torch.float32, device=dev)
            cst = torch.empty(S, ndir, H, dtype=torch.float32, device=dev)
            hprev = torch.empty(S, ndir, H, dtype=torch.float32, device=dev)
            _lib.check(lib.hsg_lstm_fwd(n_graphs, H, ndir, _p(gptr), _p(xproj), _ptr_array([q[1] for q in pl]),
                                        _ptr_array([q[2] for q in pl]), _ptr_array([q[3] for q in pl]), _p(out),
                                        _p(gates), _p(cst), _p(hprev), _st()))
            saved += [inp, gates, cst, hprev]
            inp = out
            if p_drop > 0.0 and layer + 1 < n_layers:
                from .functional import dropout_keep_mask
                keep = dropout_keep_mask(out.numel(), p_drop, seed, 1000 + layer, device=dev)
                scales.append(keep.view_as(out).float().mul_(1.0 / (1.0 - p_drop)))
                inp = out * scales[-1]
        ctx.save_for_backward(gptr, *saved, *params)
        ctx.cfg = cfg[:4]
        ctx.scales = scales
        ctx.need_dx = ctx.needs_input_grad[0]
        return inp

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        n_graphs, H, n_layers, ndir = ctx.cfg
        gptr = ctx.saved_tensors[0]
        saved = ctx.saved_tensors[1:1 + 4 * n_layers]
        params = ctx.saved_tensors[1 + 4 * n_layers:]
        d_cur = _f32c(d_out)
        S = d_cur.shape[0]
        grads = [None] * len(params)
        G4 = 4 * H
        # weight-gradient products (dW_ih, dW_hh, db of both directions) run on a side stream: only dx is on the chain
        # to the next (lower) layer's recurrence, which occupies n_graphs * ndir CTAs and leaves most SMs idle.  Same
        # kernels, same results; the main stream joins before backward returns.
        main = torch.cuda.current_stream()
        side = _side_stream(d_cur.device) if LstmFn.overlap_weight_grads else None
        for layer in range(n_layers - 1, -1, -1):
            if ctx.scales and layer + 1 < n_layers:
                d_cur = d_cur * ctx.scales[layer]
            inp, gates, cst, hprev = saved[4 * layer:4 * layer + 4]
            pl = [params[(layer * ndir + d) * 4:(layer * ndir + d) * 4 + 4] for d in range(ndir)]
            da = torch.empty(S, ndir * G4, dtype=torch.float32, device=d_cur.device)
            _lib.check(lib.hsg_lstm_bwd(n_graphs, H, ndir, _p(gptr), _p(d_cur), _p(gates), _p(cst),
                                        _ptr_array([q[1] for q in pl]), _p(da), _st()))

            def weight_grads(tag):
                for d in range(ndir):
                    da_d = da[:, d * G4:(d + 1) * G4]
                    base = (layer * ndir + d) * 4
                    if ctx.targets is not None:      # added straight into the .grad buffers; db goes to both biases
                        t = ctx.targets[base:base + 4]
                        gemm_tn(da_d, inp, ws_tag=tag, out=t[0], cs_out=t[2])
                        gemm_tn(da_d, hprev[:, d, :], ws_tag=tag, out=t[1], cs_out=t[3])
                        continue
                    dW_ih, db = gemm_tn(da_d, inp, want_colsum=True, ws_tag=tag)
                    dW_hh, _ = gemm_tn(da_d, hprev[:, d, :], ws_tag=tag)
                    grads[base], grads[base + 1], grads[base + 2], grads[base + 3] = dW_ih, dW_hh, db, db.clone()
                    if tag != "tn":                  # allocated on the side stream, consumed on the main stream
                        for g in grads[base:base + 4]:
                            g.record_stream(main)

            if side is not None:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    weight_grads("tn_side")
            else:
                weight_grads("tn")
            dx = None
            if layer > 0 or ctx.need_dx:
                for d in range(ndir):
                    da_d = da[:, d * G4:(d + 1) * G4]
                    dx = gemm_nn(da_d, pl[d][0]) if dx is None else gemm_nn(da_d, pl[d][0], R=dx, epi=_lib.EPI_ADD)
            d_cur = dx
        if side is not None:
            main.wait_stream(side)
        return (d_cur if ctx.need_dx else None, None, None) + tuple(grads)


class SentHeadFn(torch.autograd.Function):
    """sent_feature = n_feature_proj(cat[cnn_proj(ngram + sent_pos_embed(position)), lstm_proj(lstm_out)])
    (HiGraph.py:130-132, :141, :160, :96).  The concatenation is the column layout of one [S, 2 nf] buffer."""

    @staticmethod
    def forward(ctx, targets, ngram, lstm_out, sent_pos, pos_table, Wc, bc, Wl, bl, Wn):
        """targets: None, or the existing .grad buffers of (Wc, bc, Wl, bl, Wn) the backward ADDS into."""
        _lib.require_device()
        lib = _lib.load()
        ctx.targets = targets
        ngram, lstm_out = _f32c(ngram), _f32c(lstm_out)
        Wc, bc, Wl, bl, Wn, pos_table = [_f32c(t) for t in (Wc, bc, Wl, bl, Wn, pos_table)]
        S, D = ngram.shape
        nf = Wc.shape[0]
        cnn_in = torch.empty_like(ngram)
        _lib.check(lib.hsg_add_rows(S, D, _p(ngram), D, _p(sent_pos), _p(pos_table), _p(cnn_in), D, _st()))
        node = torch.empty(S, 2 * nf, dtype=torch.float32, device=ngram.device)
        gemm_nt(cnn_in, Wc, bias=bc, epi=_lib.EPI_BIAS, out=node[:, :nf])
        gemm_nt(lstm_out, Wl, bias=bl, epi=_lib.EPI_BIAS, out=node[:, nf:])
        sf = gemm_nt(node, Wn)
        ctx.save_for_backward(cnn_in, lstm_out, node, Wc, Wl, Wn)
        return sf

    @staticmethod
    def backward(ctx, d_sf):
        cnn_in, lstm_out, node, Wc, Wl, Wn = ctx.saved_tensors
        d_sf = _f32c(d_sf)
        nf = Wc.shape[0]
        d_node = gemm_nn(d_sf, Wn)
        d_cnn, d_lf = d_node[:, :nf], d_node[:, nf:]
        # the three weight-gradient products run on the side stream next to the two input-gradient products (all five
        # are small launches that leave most SMs idle); joined before returning, because autograd accumulates the
        # returned gradients on the main stream
        main = torch.cuda.current_stream()
        side = _side_stream(d_sf.device) if LstmFn.overlap_weight_grads else None
        if side is not None:
            side.wait_stream(main)
        with torch.cuda.stream(side if side is not None else main):
            tag = "tn_side" if side is not None else "tn"
            if ctx.targets is not None:
                tWc, tbc, tWl, tbl, tWn = ctx.targets
                gemm_tn(d_sf, node, ws_tag=tag, out=tWn)
                gemm_tn(d_cnn, cnn_in, ws_tag=tag, out=tWc, cs_out=tbc)
                gemm_tn(d_lf, lstm_out, ws_tag=tag, out=tWl, cs_out=tbl)
                dWn = dWc = dbc = dWl = dbl = None
            else:
                dWn, _ = gemm_tn(d_sf, node, ws_tag=tag)
                dWc, dbc = gemm_tn(d_cnn, cnn_in, want_colsum=True, ws_tag=tag)
                dWl, dbl = gemm_tn(d_lf, lstm_out, want_colsum=True, ws_tag=tag)
                if side is not None:
                    for g in (dWn, dWc, dbc, dWl, dbl):
                        g.record_stream(main)
        d_ngram = gemm_nn(d_cnn, Wc)
        d_lstm_out = gemm_nn(d_lf, Wl)
        if side is not None:
            main.wait_stream(side)
        return None, d_ngram, d_lstm_out, None, None, dWc, dbc, dWl, dbl, dWn


class _NgramParams(nn.Module):
    """Parameter container with the reference sentEncoder's names (Encoder.py:41-54): embed, position_embedding,
    convs.{0..5}.{weight,bias}; same initialisation (xavier_normal_, gain sqrt(6))."""

    def __init__(self, embed, sent_max_len, embed_size):
        super().__init__()
        self.embed = embed
        self.position_embedding = nn.Embedding.from_pretrained(sinusoid_table(sent_max_len + 1, embed_size, padding_idx=0),
                                                               freeze=True)
        self.convs = nn.ModuleList([nn.Conv2d(1, N_CHANNELS, kernel_size=(h, embed_size)) for h in KERNEL_HEIGHTS])
        for conv in self.convs:
            nn.init.xavier_normal_(conv.weight.data, gain=np.sqrt(6.0))


class SentenceEncoder(nn.Module):
    """sent_feature [S, hidden] = n_feature_proj(set_snfeature(graph)) of HSumGraph / HSumDocGraph (HiGraph.py:96,
    154-161).  Attribute names equal the reference's, so `state_dict()` carries the reference's keys for this part:
    sent_pos_embed.weight, cnn_proj.*, lstm.*, lstm_proj.*, ngram_enc.embed.weight, ngram_enc.position_embedding.weight,
    ngram_enc.convs.{i}.*, n_feature_proj.weight; parameters are created in the reference's order (same RNG stream)."""

    def __init__(self, embed, word_emb_dim=300, sent_max_len=100, doc_max_timesteps=50, n_feature_size=128,
                 hidden_size=64, lstm_hidden_state=128, lstm_layers=2, bidirectional=True, lstm_dropout=0.1):
        super().__init__()
        self._build_sn_param(embed, word_emb_dim, sent_max_len, doc_max_timesteps, n_feature_size, lstm_hidden_state,
                             lstm_layers, bidirectional, lstm_dropout)
        self._build_n_feature_proj(n_feature_size, hidden_size)

    def _build_sn_param(self, embed, word_emb_dim, sent_max_len, doc_max_timesteps, n_feature_size, lstm_hidden_state,
                        lstm_layers, bidirectional, lstm_dropout):
        """HSumGraph._init_sn_param (HiGraph.py:112-125), same creation order."""
        if word_emb_dim != NGRAM_DIM:
            # HiGraph.py:131-132 adds the 300-wide n-gram feature to a word_emb_dim-wide position embedding
            raise ValueError("word_emb_dim must be 300 (= 50 channels x 6 kernel heights), as in the reference")
        if embed.weight.requires_grad:
            raise NotImplementedError("the n-gram encoder assumes the frozen word embedding of the reference default "
                                      "(train.py:340-342)")
        self.sent_max_len = sent_max_len
        self.sent_pos_embed = nn.Embedding.from_pretrained(sinusoid_table(doc_max_timesteps + 1, word_emb_dim,
                                                                          padding_idx=0), freeze=True)
        self.cnn_proj = nn.Linear(word_emb_dim, n_feature_size)
        self.lstm = nn.LSTM(word_emb_dim, lstm_hidden_state, num_layers=lstm_layers, dropout=lstm_dropout,
                            batch_first=True, bidirectional=bidirectional)
        self.lstm_proj = nn.Linear(lstm_hidden_state * (2 if bidirectional else 1), n_feature_size)
        self.ngram_enc = _NgramParams(embed, sent_max_len, word_emb_dim)

    def _build_n_feature_proj(self, n_feature_size, hidden_size):
        self.n_feature_proj = nn.Linear(n_feature_size * 2, hidden_size, bias=False)      # HiGraph.py:53

    def ngram(self, plan: EncoderPlan):
        if plan.L != self.sent_max_len:
            raise ValueError("token matrix width %d != sent_max_len %d" % (plan.L, self.sent_max_len))
        conv = []
        for c in self.ngram_enc.convs:
            conv += [c.weight, c.bias]
        return NgramEncodeFn.apply(plan, self._targets(conv), self.ngram_enc.embed.weight,
                                   self.ngram_enc.position_embedding.weight, *conv)

    # When True, backward ADDS the parameter gradients straight into the existing `.grad` buffers (e.g. the views of a
    # dist.FlatGradArena) from inside the library's kernels instead of returning them to autograd, which would launch
    # one add per parameter (33 here).  Same arithmetic; requires every parameter to have a `.grad` tensor.  Same
    # switch as WSWGATUpdateLoop.fuse_grad_accumulation (the whole-model classes share it).
    fuse_grad_accumulation = False

    def _targets(self, params):
        if not (self.fuse_grad_accumulation and torch.is_grad_enabled()):
            return None
        if any(p.grad is None for p in params):
            raise RuntimeError("fuse_grad_accumulation needs a .grad buffer on every parameter")
        return [p.grad for p in params]

    def lstm_feature(self, plan: EncoderPlan, ngram):
        """LSTM over every graph's sentence sequence (HiGraph.py:135-141), rows in batch order."""
        lstm = self.lstm
        p_drop = float(lstm.dropout) if (self.training and lstm.num_layers > 1) else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p_drop > 0.0 else 0     # CPU generator: no sync
        cfg = (plan.n_graphs, lstm.hidden_size, lstm.num_layers, 2 if lstm.bidirectional else 1, p_drop, seed,
               self._targets(lstm._flat_weights))
        return LstmFn.apply(ngram, plan.graph_sent_ptr, cfg, *lstm._flat_weights)

    def forward(self, plan: EncoderPlan):
        return self.encode(plan)

    def encode(self, plan: EncoderPlan):
        ngram = self.ngram(plan)
        lstm_out = self.lstm_feature(plan, ngram)
        head = [self.cnn_proj.weight, self.cnn_proj.bias, self.lstm_proj.weight, self.lstm_proj.bias,
                self.n_feature_proj.weight]
        return SentHeadFn.apply(self._targets(head), ngram, lstm_out, plan.sent_pos, self.sent_pos_embed.weight, *head)
