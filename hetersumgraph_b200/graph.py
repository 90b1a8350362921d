"""HeteroBatch: the device-resident replacement of the reference's BatchedDGLGraph.

The reference hands a DGL-0.4 batched graph `g` to WSWGAT.forward (module/GAT.py:45)
and every head re-derives the node / edge id sets with three filter UDFs
(module/GATLayer.py:105-107, 143-145).  HeteroBatch carries those id sets once per
batch, built on the GPU by hsg_build_count / hsg_build_fill from the same inputs
ExampleSet.CreateGraph consumes (module/dataloader.py:222-268, 328-406):

  word rows      = ascending DGL node id over unit == 0   (row order of `w`)
  supernode rows = ascending DGL node id over unit == 1   (row order of `s`: sentences, then docs, per graph)
  csc_super      = in-edges of supernodes from words  (w->s / w->d), ascending DGL edge id
  csc_word       = in-edges of words from supernodes  (s->w / d->w), ascending DGL edge id
  super_extra    = number of other in-edges of a supernode (sent<->sent dtype 1, sent->doc dtype 2):
                   they carry e = 0 and z_src = 0 into the reference's softmax
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from .synthetic import TokenBatch


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


class DeviceTokenBatch:
    """The builder's inputs resident on the device (see synthetic.TokenBatch for the fields)."""

    def __init__(self):
        self.c_struct = None

    @staticmethod
    def host_buffers(tb: TokenBatch, pin=True):
        """Everything that is copied per step packed into ONE contiguous (pinned) host blob - a single H2D copy per
        batch - with the segment layout and the host-known sizes.  Returns (host, bytes copied per step)."""
        arrays = dict(tokens=tb.tokens, sent_bin=tb.sent_bin, graph_sent_ptr=tb.graph_sent_ptr, labels=tb.labels)
        if tb.hdsg:
            per = np.diff(tb.graph_sent_ptr)
            sent_graph = np.repeat(np.arange(tb.n_graphs, dtype=np.int64), per)
            sent_local = (np.arange(int(per.sum()), dtype=np.int32) - np.repeat(tb.graph_sent_ptr[:-1], per)).astype(np.int32)
            per_d = np.diff(tb.graph_doc_ptr)
            doc_graph = np.repeat(np.arange(tb.n_graphs, dtype=np.int64), per_d)
            doc_local = (np.arange(int(per_d.sum()), dtype=np.int32) - np.repeat(tb.graph_doc_ptr[:-1], per_d)).astype(np.int32)
            sent_doc_g = (tb.graph_doc_ptr[:-1][sent_graph] + tb.sent_doc[:len(sent_graph)]).astype(np.int32)
            arrays.update(graph_doc_ptr=tb.graph_doc_ptr, sent_doc=tb.sent_doc, doc_tok_ptr=tb.doc_tok_ptr,
                          doc_tokens=tb.doc_tokens, doc_bin=tb.doc_bin, sent_graph=sent_graph, sent_local=sent_local,
                          doc_graph=doc_graph, doc_local=doc_local, sent_doc_g=sent_doc_g,
                          doc_graph32=doc_graph.astype(np.int32))
        layout, off = {}, 0
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            arrays[name] = a
            layout[name] = (off, a.nbytes, torch.from_numpy(a[:0].copy()).dtype if a.size == 0 else torch.from_numpy(a.reshape(-1)[:1]).dtype,
                            tuple(a.shape))
            off += (a.nbytes + 15) & ~15
        blob = np.zeros(max(off, 16), np.uint8)
        for name, a in arrays.items():
            o, nb, _, _ = layout[name]
            blob[o:o + nb] = a.reshape(-1).view(np.uint8)
        t = torch.from_numpy(blob)
        if pin and torch.cuda.is_available():
            t = t.pin_memory()
        S, L = tb.tokens.shape
        per_graph = np.diff(tb.graph_sent_ptr)
        meta = dict(S=int(S), L=int(L), max_sent=int(per_graph.max()) if tb.n_graphs > 0 else 0,
                    n_doc=int(tb.graph_doc_ptr[-1]) if tb.hdsg else 0, n_doc_tok=int(tb.doc_tok_ptr[-1]) if tb.hdsg else 0,
                    vocab=int(tb.filter_bitmap.shape[0]) * 32, n_graphs=int(tb.n_graphs), hdsg=bool(tb.hdsg))
        nbytes = sum(v[1] for v in layout.values())
        return dict(blob=t, layout=layout, meta=meta, tb=tb), nbytes

    @staticmethod
    def upload(tb: TokenBatch, device="cuda", vocab_size: Optional[int] = None, host=None,
               filter_bitmap_dev: Optional[torch.Tensor] = None, blob_dev: Optional[torch.Tensor] = None,
               copy: bool = True) -> "DeviceTokenBatch":
        """blob_dev: an existing device byte buffer (>= the blob size) that receives the copy instead of a fresh
        allocation - fixed addresses for CUDA-graph replay (step_graph.StaticBatchSlot); copy=False only creates the
        views over it (the caller enqueues the copy itself)."""
        dev = torch.device(device)
        if host is None:
            host, _ = DeviceTokenBatch.host_buffers(tb)
        d = DeviceTokenBatch()
        d.device, d.hdsg, d.n_graphs = dev, bool(tb.hdsg), tb.n_graphs
        d.host_tb = tb                                  # host arrays: the sentence encoder's plan is made from them
        if blob_dev is None:
            blob = host["blob"].to(dev, non_blocking=True)             # the ONE host -> device copy of the batch
        else:
            nb = host["blob"].numel()
            if blob_dev.numel() < nb:
                raise ValueError("blob_dev holds %d bytes, the batch needs %d" % (blob_dev.numel(), nb))
            blob = blob_dev[:nb]
            if copy:
                blob.copy_(host["blob"], non_blocking=True)
        d._blob = blob
        lay, meta = host["layout"], host["meta"]

        def view(name):
            o, nb, dt, shape = lay[name]
            if nb == 0:
                return None
            return blob[o:o + nb].view(dt).view(shape)

        d.tokens, d.sent_bin, d.graph_sent_ptr, d.labels = view("tokens"), view("sent_bin"), view("graph_sent_ptr"), view("labels")
        if filter_bitmap_dev is None:   # constant per dataset: upload once and pass it back in for later batches
            filter_bitmap_dev = torch.from_numpy(tb.filter_bitmap.view(np.int32).copy()).to(dev)
        d.filter_bitmap = filter_bitmap_dev
        if vocab_size is None:
            vocab_size = meta["vocab"]
        d.graph_doc_ptr = d.sent_doc = d.doc_tok_ptr = d.doc_tokens = d.doc_bin = d.sent_graph = d.sent_local = None
        d.doc_graph = d.doc_local = d.sent_doc_g = d.doc_graph32 = None
        d.n_sent = meta["S"]
        if tb.hdsg:
            d.graph_doc_ptr, d.sent_doc, d.doc_tok_ptr = view("graph_doc_ptr"), view("sent_doc"), view("doc_tok_ptr")
            d.doc_tokens, d.doc_bin = view("doc_tokens"), view("doc_bin")
            d.sent_graph, d.sent_local = view("sent_graph"), view("sent_local")
            d.doc_graph, d.doc_local = view("doc_graph"), view("doc_local")
            d.sent_doc_g, d.doc_graph32 = view("sent_doc_g"), view("doc_graph32")
        d.c_struct = _lib.TokenBatchC(tb.n_graphs, meta["S"], meta["L"], int(tb.hdsg), int(vocab_size), meta["n_doc"],
                                      meta["n_doc_tok"], meta["max_sent"],
                                      _ptr(d.tokens), _ptr(d.sent_bin), _ptr(d.graph_sent_ptr), _ptr(d.filter_bitmap),
                                      _ptr(d.graph_doc_ptr), _ptr(d.sent_doc), _ptr(d.doc_tok_ptr), _ptr(d.doc_tokens),
                                      _ptr(d.doc_bin))
        return d


@dataclass
class HeteroBatch:
    n_graphs: int
    n_word: int
    n_super: int
    n_pair: int
    # [B+1] offsets
    word_ptr: torch.Tensor
    super_ptr: torch.Tensor
    node_ptr: torch.Tensor
    edge_ptr: torch.Tensor
    pair_ptr: torch.Tensor
    # node maps
    word_wid: torch.Tensor
    word_nid: torch.Tensor
    super_nid: torch.Tensor
    super_type: torch.Tensor
    super_graph: torch.Tensor
    super_extra: torch.Tensor
    # CSCs
    super_indptr: torch.Tensor
    super_src: torch.Tensor
    super_bin: torch.Tensor
    super_eid: torch.Tensor
    word_indptr: torch.Tensor
    word_src: torch.Tensor
    word_bin: torch.Tensor
    word_eid: torch.Tensor
    word_extra: Optional[torch.Tensor] = None
    tfidfembed_weight: Optional[torch.Tensor] = None     # set by set_tfidf_embedding (HiGraph.py:150-151)
    labels: Optional[torch.Tensor] = None                # [n sentence rows] int64
    sent_doc_row: Optional[torch.Tensor] = None          # HDSG: supernode row of each sentence's document
    sent_row: Optional[torch.Tensor] = None              # HDSG: supernode row of each sentence (None: identity)
    doc_row: Optional[torch.Tensor] = None               # HDSG: supernode row of each document
    sent_doc_gidx: Optional[torch.Tensor] = None         # HDSG: global document index of each sentence (int32)
    doc_graph: Optional[torch.Tensor] = None             # HDSG: graph index of each document (int32)
    graph_sent_ptr: Optional[torch.Tensor] = None        # [B+1] sentence offsets per graph
    n_total_nodes: int = 0
    n_total_edges: int = 0

    def __post_init__(self):
        self._csc_super = _lib.CscC(self.n_super, self.n_word, self.n_pair, 0, _ptr(self.super_indptr),
                                    _ptr(self.super_src), _ptr(self.super_bin), _ptr(self.super_extra))
        self._csc_word = _lib.CscC(self.n_word, self.n_super, self.n_pair, 0, _ptr(self.word_indptr),
                                   _ptr(self.word_src), _ptr(self.word_bin), _ptr(self.word_extra))

    # ---- reference-facing helpers ------------------------------------------------
    def set_tfidf_embedding(self, weight: torch.Tensor):
        """Counterpart of HSumGraph.set_wnfeature's edata['tfidfembed'] write (HiGraph.py:150-151):
        only the 10 x feat_embed table is needed, the per-edge gather is folded into the edge kernel."""
        self.tfidfembed_weight = weight
        return self

    @property
    def device(self):
        return self.super_indptr.device

    def csc(self, kind: str):
        """(forward CSC, transposed CSC) for layer type `kind`."""
        if kind == "W2S":
            return self._csc_super, self._csc_word
        if kind == "S2W":
            return self._csc_word, self._csc_super
        raise NotImplementedError("GAT Layer has not been implemented!")   # module/GAT.py:41

    def s2s_groups(self):
        """(xgrp, xmember, mult) int32 maps of the implicit extra in-edges used by the S2S layer type
        (csrc/hsg_s2s.cu): HSG - every sentence belongs to and reads its graph's group, each ordered pair twice
        (dataloader.py:262-263); HDSG - sentences belong to their document's group, documents read it (:385)."""
        cached = self.__dict__.get("_s2s_groups")
        if cached is not None:
            return cached
        if self.sent_doc_row is None:                       # HSG: group id = first supernode row of the graph
            grp = self.super_ptr[:-1][self.super_graph.long()].int().contiguous()
            out = (grp, grp, 2)
        else:
            n = self.n_super
            xgrp = torch.full((n,), -1, dtype=torch.int32, device=self.device)
            xmember = torch.full((n,), -1, dtype=torch.int32, device=self.device)
            drows = self.doc_rows()
            xgrp[drows] = drows.int()
            xmember[self.sentence_rows()] = self.sent_doc_row.int()
            out = (xgrp, xmember, 1)
        self.__dict__["_s2s_groups"] = out
        return out

    @property
    def encoder_plan(self):
        """EncoderPlan of this batch for encoder.SentenceEncoder (the counterpart of the `words` / `position` node data
        the reference's set_snfeature reads, HiGraph.py:128-131): made on first use from the host token matrix the
        batch was built from, sharing the device copy of the tokens."""
        plan = self.__dict__.get("_encoder_plan")
        if plan is None:
            dtb = self.__dict__.get("_keepalive", (None,))[0]
            tb = getattr(dtb, "host_tb", None)
            if tb is None:
                raise RuntimeError("this HeteroBatch was not built from a TokenBatch: pass an EncoderPlan explicitly")
            from .encoder import EncoderPlan
            plan = EncoderPlan.from_token_batch(tb, self.device, tokens_dev=dtb.tokens)
            self.__dict__["_encoder_plan"] = plan
        return plan

    def sentence_rows(self) -> torch.Tensor:
        """rows of `s` that are sentence nodes (dtype == 1), ascending (HiGraph.py:191).  Taken from the builder's
        maps when present (no host synchronisation), else derived from super_type."""
        if self.sent_row is not None:
            return self.sent_row.long()
        return torch.nonzero(self.super_type == 1).reshape(-1)

    def doc_rows(self) -> torch.Tensor:
        if self.doc_row is not None:
            return self.doc_row
        return torch.nonzero(self.super_type == 2).reshape(-1)

    # ---- construction -----------------------------------------------------------
    @staticmethod
    def from_token_batch(tb: TokenBatch, device="cuda", vocab_size: Optional[int] = None) -> "HeteroBatch":
        """Host arrays -> device (pinned H2D) -> device-side build (K0)."""
        return HeteroBatch.build(DeviceTokenBatch.upload(tb, device, vocab_size))

    @staticmethod
    def build(dtb: "DeviceTokenBatch") -> "HeteroBatch":
        """Device-side build (K0) from device-resident token arrays.  One small D2H read of the
        five totals (+ status) sizes the outputs of the fill phase."""
        return HeteroBatch._build_fill(HeteroBatch._build_count(dtb))

    @staticmethod
    def _build_count(dtb: "DeviceTokenBatch", stream_obj=None):
        """Phase 1 on the current stream: per-graph counts + offsets; the totals start their way to pinned host
        memory.  Returns the state _build_fill needs."""
        _lib.require_device()
        lib = _lib.load()
        dev = dtb.device
        B = dtb.n_graphs
        tbc = dtb.c_struct
        ws_bytes = lib.hsg_build_workspace_bytes(C.byref(tbc))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        # [5, B+1] offsets followed by the status word: one buffer, one zero fill, ONE small D2H (no gather launch)
        meta = torch.zeros(5 * (B + 1) + 1, dtype=torch.int32, device=dev)
        offs = meta[:5 * (B + 1)].view(5, B + 1)
        status = meta[5 * (B + 1):]
        off_c = _lib.GraphOffsetsC(*[offs.data_ptr() + 4 * i * (B + 1) for i in range(5)])
        st = _stream()
        _lib.check(lib.hsg_build_count(C.byref(tbc), off_c, status.data_ptr(), ws.data_ptr(), ws_bytes, st))
        totals_host = torch.empty(5 * (B + 1) + 1, dtype=torch.int32, pin_memory=True)
        totals_host.copy_(meta, non_blocking=True)                       # the one D2H of the build
        totals_dev = meta
        ev = torch.cuda.Event()
        ev.record(stream_obj) if stream_obj is not None else ev.record()
        return dict(dtb=dtb, ws=ws, ws_bytes=ws_bytes, offs=offs, status=status, off_c=off_c,
                    totals_dev=totals_dev, totals_host=totals_host, event=ev)

    @staticmethod
    def _build_fill(c) -> "HeteroBatch":
        """Phase 2 on the current stream (the same one phase 1 ran on): waits for the totals, sizes and fills
        the node maps and both CSCs."""
        lib = _lib.load()
        dtb, ws, ws_bytes, offs, status, off_c = c["dtb"], c["ws"], c["ws_bytes"], c["offs"], c["status"], c["off_c"]
        dev, B, tbc = dtb.device, dtb.n_graphs, dtb.c_struct
        c["event"].synchronize()
        th = c["totals_host"].tolist()
        status_h = th[5 * (B + 1)]
        if status_h != 0:
            _lib.check(status_h)
        n_word, n_super, n_node, n_edge, n_pair = [th[i * (B + 1) + B] for i in range(5)]
        st = _stream()
        i32 = dict(dtype=torch.int32, device=dev)
        # one int32 arena for every 4-byte array, one byte arena for the rest: 2 allocations instead of 14
        sizes32 = [("word_wid", max(n_word, 1)), ("word_nid", max(n_word, 1)), ("super_nid", max(n_super, 1)),
                   ("super_graph", max(n_super, 1)), ("super_indptr", n_super + 1), ("super_src", max(n_pair, 1)),
                   ("super_eid", max(n_pair, 1)), ("super_extra", max(n_super, 1)), ("word_indptr", n_word + 1),
                   ("word_src", max(n_pair, 1)), ("word_eid", max(n_pair, 1))]
        sizes8 = [("super_type", max(n_super, 1)), ("super_bin", max(n_pair, 1)), ("word_bin", max(n_pair, 1))]
        out = {}
        pad32 = [(n + 3) & ~3 for _, n in sizes32]
        a32 = torch.empty(sum(pad32), **i32)
        for (name, n), part in zip(sizes32, a32.split_with_sizes(pad32)):     # one call -> all the views
            out[name] = part[:n] if part.shape[0] != n else part
        pad8 = [(n + 15) & ~15 for _, n in sizes8]
        a8 = torch.empty(sum(pad8), dtype=torch.uint8, device=dev)
        for (name, n), part in zip(sizes8, a8.split_with_sizes(pad8)):
            out[name] = part[:n] if part.shape[0] != n else part
        out["super_type"] = out["super_type"].view(torch.int8)
        if B == 0:
            out["super_indptr"].zero_()
            out["word_indptr"].zero_()
        goc = _lib.GraphOutC(n_word, n_super, n_pair, 0, off_c,
                             *[out[k].data_ptr() for k in ("word_wid", "word_nid", "super_nid", "super_type",
                                                            "super_graph", "super_indptr", "super_src", "super_bin",
                                                            "super_eid", "super_extra", "word_indptr", "word_src",
                                                            "word_bin", "word_eid")], status.data_ptr())
        _lib.check(lib.hsg_build_fill(C.byref(tbc), C.byref(goc), ws.data_ptr(), ws_bytes, st))
        hb = HeteroBatch(
            n_graphs=B, n_word=n_word, n_super=n_super, n_pair=n_pair,
            word_ptr=offs[0], super_ptr=offs[1], node_ptr=offs[2], edge_ptr=offs[3], pair_ptr=offs[4],
            word_wid=out["word_wid"][:n_word], word_nid=out["word_nid"][:n_word],
            super_nid=out["super_nid"][:n_super], super_type=out["super_type"][:n_super],
            super_graph=out["super_graph"][:n_super], super_extra=out["super_extra"][:n_super],
            super_indptr=out["super_indptr"], super_src=out["super_src"][:n_pair],
            super_bin=out["super_bin"][:n_pair], super_eid=out["super_eid"][:n_pair],
            word_indptr=out["word_indptr"], word_src=out["word_src"][:n_pair], word_bin=out["word_bin"][:n_pair],
            word_eid=out["word_eid"][:n_pair], n_total_nodes=n_node, n_total_edges=n_edge)
        hb._keepalive = (dtb, ws, status, a32, a8, c["totals_dev"])
        hb.labels = dtb.labels
        hb.graph_sent_ptr = dtb.graph_sent_ptr
        if dtb.hdsg and dtb.n_sent > 0:
            # supernode rows of a graph: its sentences, then its documents (dataloader.py:348-363) - no host sync
            base = offs[1][:B][dtb.sent_graph]                                       # first supernode row of the graph
            n_per = (dtb.graph_sent_ptr[1:] - dtb.graph_sent_ptr[:-1])[dtb.sent_graph]
            hb.sent_row = (base + dtb.sent_local[:dtb.n_sent]).int()
            hb.sent_doc_row = (base + n_per + dtb.sent_doc[:dtb.n_sent]).long()
            if dtb.doc_graph is not None:                     # supernode row of every document, in document order
                n_per_g = dtb.graph_sent_ptr[1:] - dtb.graph_sent_ptr[:-1]
                hb.doc_row = (offs[1][:B][dtb.doc_graph] + n_per_g[dtb.doc_graph] + dtb.doc_local).long()
                hb.sent_doc_gidx, hb.doc_graph = dtb.sent_doc_g, dtb.doc_graph32
        return hb

    @staticmethod
    def from_csc_arrays(super_indptr, super_src, super_bin, super_extra, word_indptr, word_src, word_bin,
                        n_graphs=1, super_type=None, super_graph=None, word_extra=None, super_eid=None,
                        word_eid=None, device="cuda") -> "HeteroBatch":
        """Wrap pre-built CSC arrays (stress graph of SURVEY.md §8-d; parity tests against the oracle's arrays)."""
        dev = torch.device(device)

        def t(a, dtype):
            if a is None:
                return None
            return torch.as_tensor(np.ascontiguousarray(a)).to(dev).to(dtype).contiguous()

        n_super = len(super_indptr) - 1
        n_word = len(word_indptr) - 1
        n_pair = int(len(super_src))
        z32 = torch.zeros(1, dtype=torch.int32, device=dev)
        pad = lambda x, dt: x if x is not None and x.numel() > 0 else torch.zeros(1, dtype=dt, device=dev)  # noqa: E731
        return HeteroBatch(
            n_graphs=n_graphs, n_word=n_word, n_super=n_super, n_pair=n_pair,
            word_ptr=z32, super_ptr=z32, node_ptr=z32, edge_ptr=z32, pair_ptr=z32,
            word_wid=torch.zeros(n_word, dtype=torch.int32, device=dev),
            word_nid=torch.arange(n_word, dtype=torch.int32, device=dev),
            super_nid=torch.arange(n_super, dtype=torch.int32, device=dev),
            super_type=t(super_type, torch.int8) if super_type is not None else torch.ones(n_super, dtype=torch.int8, device=dev),
            super_graph=t(super_graph, torch.int32) if super_graph is not None else torch.zeros(n_super, dtype=torch.int32, device=dev),
            super_extra=pad(t(super_extra, torch.int32), torch.int32),
            super_indptr=t(super_indptr, torch.int32), super_src=pad(t(super_src, torch.int32), torch.int32),
            super_bin=pad(t(super_bin, torch.uint8), torch.uint8),
            super_eid=pad(t(super_eid, torch.int32), torch.int32) if super_eid is not None else z32,
            word_indptr=t(word_indptr, torch.int32), word_src=pad(t(word_src, torch.int32), torch.int32),
            word_bin=pad(t(word_bin, torch.uint8), torch.uint8),
            word_eid=pad(t(word_eid, torch.int32), torch.int32) if word_eid is not None else z32,
            word_extra=t(word_extra, torch.int32))


def csc_pair_from_edges(word_row, super_row, bins, n_word, n_super):
    """Host helper for synthetic stress graphs: both CSCs of a list of word<->supernode pairs
    given in DGL insertion order (pair t has edge ids 2t / 2t+1)."""
    word_row = np.asarray(word_row, np.int64)
    super_row = np.asarray(super_row, np.int64)
    bins = np.asarray(bins, np.int64)
    t = np.arange(len(word_row), dtype=np.int64)

    def one(dst, src, n_dst, eid):
        o = np.argsort(dst, kind="stable")
        indptr = np.zeros(n_dst + 1, np.int64)
        np.add.at(indptr, dst + 1, 1)
        return np.cumsum(indptr), src[o], bins[o], eid[o]

    return one(super_row, word_row, n_super, 2 * t), one(word_row, super_row, n_word, 2 * t + 1)


class BuildPipeline:
    """Double-buffered device-side graph builds on a side stream.

    The reference builds its DGL graphs in DataLoader worker processes while the previous batch trains
    (module/dataloader.py:222-268 under torch.utils.data.DataLoader, train.py).  Here the build of batch i+1
    (hsg_build_count, the 24-byte D2H of the totals, hsg_build_fill) runs on its own CUDA stream while the
    kernels of batch i run on the compute stream, so neither the two latency-bound builder kernels (one CTA per
    graph) nor the host read of the totals sit on the critical path:

        pipe.submit(tokens_0); pipe.finish()           # prime
        for i in ...:
            batch = pipe.take()                        # compute stream waits for the fill of batch i (done long ago)
            pipe.submit(tokens_{i+1})                  # phase 1 of the next build, side stream
            ... enqueue forward / backward / optimizer on the compute stream ...
            pipe.finish()                              # totals have arrived: size + phase 2, side stream
    """

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        # the compute stream is sampled once: torch.cuda.current_stream() costs ~12 us per call, the stream context
        # manager three of them, and a training loop does not change its compute stream between steps
        self.main = torch.cuda.current_stream(self.device)
        self._pending = None
        self._ready = None

    def submit(self, source):
        """`source`: a DeviceTokenBatch, or a callable returning one (e.g. the pinned-host -> device upload), run
        on the side stream."""
        if self._pending is not None or self._ready is not None:
            raise RuntimeError("BuildPipeline: previous batch not taken yet")
        self.stream.wait_stream(self.main)
        torch.cuda.set_stream(self.stream)
        try:
            dtb = source() if callable(source) else source
            self._pending = HeteroBatch._build_count(dtb, self.stream)
        finally:
            torch.cuda.set_stream(self.main)

    def finish(self):
        if self._pending is None:
            return
        torch.cuda.set_stream(self.stream)
        try:
            hb = HeteroBatch._build_fill(self._pending)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        finally:
            torch.cuda.set_stream(self.main)
        self._pending = None
        self._ready = (hb, ev)

    def take(self) -> "HeteroBatch":
        if self._ready is None:
            self.finish()
        if self._ready is None:
            raise RuntimeError("BuildPipeline: nothing submitted")
        hb, ev = self._ready
        self._ready = None
        cur = self.main
        cur.wait_event(ev)
        # allocated on the side stream, consumed on the compute stream: tell the caching allocator.  Every field of
        # the batch is a view of one of these few storages.
        bases = [t for t in hb._keepalive if isinstance(t, torch.Tensor)]
        blob = getattr(hb._keepalive[0], "_blob", None)
        if blob is not None:
            bases.append(blob)
        for extra in (hb.sent_row, hb.sent_doc_row, hb.doc_row):
            if extra is not None:
                bases.append(extra)
        for t in bases:
            if t.is_cuda:
                t.record_stream(cur)
        return hb
