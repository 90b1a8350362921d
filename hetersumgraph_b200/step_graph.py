"""One training step of the HSG / HDSG path as a replayed CUDA graph.

The reference's step (train.py:108-135) is: DataLoader workers build the DGL graphs of the next batch while the model
runs `forward -> loss -> backward -> clip -> Adam` on the current one.  `path_model.FusedTrainStep` already issues that
work as ~60 kernel launches from a handful of C calls; at batch 32 the host's enqueue time is then as long as the GPU
time of the step.  `GraphedTrainStep` captures the whole step ONCE per batch shape -

    side branch   H2D of the next batch's token blob -> hsg_build_count -> D2H of its totals -> hsg_build_fill
    main branch   (H2D of sent_feature) -> hsg_embed_gather -> hsg_update_loop_fwd -> hsg_head_fwd / _bwd ->
                  hsg_update_loop_bwd (with its internal weight-gradient side stream) -> [NCCL all-reduce] ->
                  hsg_adam_step_dev (Adam + zero_grad, step number on the device) -> D2H of the loss

- and replays it with one cudaGraphLaunch per step.  Everything a replay touches lives at a fixed address: two
`StaticBatchSlot`s (token blob, builder workspace and outputs at capacity; batch i computes out of slot i % 2 while batch
i+1 is built into the other one), pinned host staging buffers, the graph's private memory pool.  Whatever changes from
step to step is DATA, not a kernel argument: the Adam step number and the dropout step counter are device scalars.

A graph is keyed by (slot parity, sizes of the batch being computed, sizes of the batch being built).  A key that has
not been seen runs eagerly through exactly the same enqueue function (bit-identical results - tested) and is captured
for the next time; `capture=False` always runs eagerly.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .graph import DeviceTokenBatch, HeteroBatch, _ptr


def _raw(stream):
    return stream.cuda_stream


class StaticBatchSlot:
    """Fixed-address device home of one in-flight batch: token blob, builder workspace, offsets / status words and
    every builder output at capacity.  `bind(host)` lays the views of a concrete batch over it; `enqueue_build`
    issues (copy +) count + totals D2H + fill with NO host synchronisation in between (the fill kernel only needs the
    capacities, hsg_graph_out.cap_*); `batch(totals)` wraps the result as a HeteroBatch of the actual sizes."""

    GROW = 1.25

    def __init__(self, device):
        self.device = torch.device(device)
        self.generation = 0             # bumped whenever buffers are re-allocated (captured graphs become invalid)
        self.cap = None
        self.dtb = None
        self.host = None
        self.loaded = None              # host dict whose token blob currently sits in self.blob (resident mode)
        self._bkey = None
        self._batches = {}

    # -- capacity -------------------------------------------------------------------------------------------------
    @staticmethod
    def _need(host):
        m = host["meta"]
        S, L, B = m["S"], m["L"], m["n_graphs"]
        n_tok = S * L + m["n_doc_tok"]
        return dict(blob=host["blob"].numel(), B=B, word=max(n_tok, 1), pair=max(n_tok, 1),
                    sup=max(S + m["n_doc"], 1), sent=max(S, 1), doc=max(m["n_doc"], 1))

    def _alloc(self, need, tbc):
        dev = self.device
        g = self.GROW
        cap = {k: (v if k == "B" else int(v * g) + 16) for k, v in need.items()}
        self.cap = cap
        self.blob = torch.empty(cap["blob"], dtype=torch.uint8, device=dev)
        B = cap["B"]
        self.meta = torch.zeros(5 * (B + 1) + 1, dtype=torch.int32, device=dev)
        self.totals_host = torch.zeros(5 * (B + 1) + 1, dtype=torch.int32).pin_memory()
        sizes32 = [("word_wid", cap["word"]), ("word_nid", cap["word"]), ("super_nid", cap["sup"]),
                   ("super_graph", cap["sup"]), ("super_indptr", cap["sup"] + 1), ("super_src", cap["pair"]),
                   ("super_eid", cap["pair"]), ("super_extra", cap["sup"]), ("word_indptr", cap["word"] + 1),
                   ("word_src", cap["pair"]), ("word_eid", cap["pair"]), ("sent_row", cap["sent"])]
        sizes8 = [("super_type", cap["sup"]), ("super_bin", cap["pair"]), ("word_bin", cap["pair"])]
        pad32 = [(n + 3) & ~3 for _, n in sizes32]
        self.a32 = torch.zeros(sum(pad32), dtype=torch.int32, device=dev)
        self.arr = {name: part for (name, _), part in zip(sizes32, self.a32.split_with_sizes(pad32))}
        pad8 = [(n + 15) & ~15 for _, n in sizes8]
        self.a8 = torch.zeros(sum(pad8), dtype=torch.uint8, device=dev)
        for (name, _), part in zip(sizes8, self.a8.split_with_sizes(pad8)):
            self.arr[name] = part
        self.sent_doc_row = torch.zeros(cap["sent"], dtype=torch.int64, device=dev)      # HDSG maps (int64 rows)
        self.doc_row = torch.zeros(cap["doc"], dtype=torch.int64, device=dev)
        self.ws = None
        self.generation += 1
        self._batches = {}

    def _fits(self, need):
        return self.cap is not None and need["B"] == self.cap["B"] and all(
            need[k] <= self.cap[k] for k in need if k != "B")

    # -- per batch ------------------------------------------------------------------------------------------------
    def bind(self, host, filter_bitmap_dev, vocab_size=None):
        """Lay the views of the batch described by `host` (DeviceTokenBatch.host_buffers) over the slot.  Nothing is
        copied or launched.  Returns the shape key of the BUILD of this batch."""
        if host is self.host and self.dtb is not None:
            return self._bkey                           # same host object re-submitted: views are already in place
        need = self._need(host)
        if not self._fits(need):
            self._alloc(need, None)
        tb = host["tb"]
        self.host = host
        self.loaded = None
        self.dtb = DeviceTokenBatch.upload(tb, self.device, vocab_size, host=host, filter_bitmap_dev=filter_bitmap_dev,
                                           blob_dev=self.blob, copy=False)
        lib = _lib.load()
        ws_bytes = lib.hsg_build_workspace_bytes(C.byref(self.dtb.c_struct))
        if self.ws is None or self.ws.numel() < ws_bytes:
            self.ws = torch.empty(int(ws_bytes * self.GROW) + 256, dtype=torch.uint8, device=self.device)
            self.generation += 1
        m = host["meta"]
        lay = tuple((k, v[0], v[1]) for k, v in host["layout"].items())
        self._bkey = (m["n_graphs"], m["S"], m["L"], m["max_sent"], m["n_doc"], m["n_doc_tok"], m["hdsg"], hash(lay))
        return self._bkey

    def enqueue_build(self, stream, blob_host=None):
        """On `stream` (a torch.cuda.Stream that is current): [H2D of the token blob from pinned staging], phase 1,
        the small D2H of the totals, phase 2 at capacity."""
        lib = _lib.load()
        dtb = self.dtb
        B = dtb.n_graphs
        st = _raw(stream)
        if blob_host is not None:
            dtb._blob.copy_(blob_host[:dtb._blob.numel()], non_blocking=True)
        tbc = dtb.c_struct
        offs = self.meta[:5 * (B + 1)].view(5, B + 1)
        status = self.meta[5 * (B + 1):]
        off_c = _lib.GraphOffsetsC(*[offs.data_ptr() + 4 * i * (B + 1) for i in range(5)])
        _lib.check(lib.hsg_memset(status.data_ptr(), 0, 4, st))
        _lib.check(lib.hsg_build_count(C.byref(tbc), off_c, status.data_ptr(), self.ws.data_ptr(), self.ws.numel(), st))
        cap, a = self.cap, self.arr
        goc = _lib.GraphOutC(cap["word"], cap["sup"], cap["pair"], 0, off_c,
                             *[a[k].data_ptr() for k in ("word_wid", "word_nid", "super_nid", "super_type",
                                                         "super_graph", "super_indptr", "super_src", "super_bin",
                                                         "super_eid", "super_extra", "word_indptr", "word_src",
                                                         "word_bin", "word_eid")], status.data_ptr())
        _lib.check(lib.hsg_build_fill(C.byref(tbc), C.byref(goc), self.ws.data_ptr(), self.ws.numel(), st))
        self._off_c = off_c
        if dtb.hdsg and dtb.n_sent > 0:
            # HDSG row maps (graph.HeteroBatch._build_fill) written into the slot's fixed buffers
            S = dtb.n_sent
            base = offs[1][:B][dtb.sent_graph]
            n_per_g = dtb.graph_sent_ptr[1:] - dtb.graph_sent_ptr[:-1]
            a["sent_row"][:S].copy_(base + dtb.sent_local[:S])
            self.sent_doc_row[:S].copy_(base + n_per_g[dtb.sent_graph] + dtb.sent_doc[:S])
            if dtb.doc_graph is not None:
                nd = dtb.doc_graph.shape[0]
                self.doc_row[:nd].copy_(offs[1][:B][dtb.doc_graph] + n_per_g[dtb.doc_graph] + dtb.doc_local)
        # the totals leave last so that their arrival also means "status is final"
        self.totals_host.copy_(self.meta, non_blocking=True)

    def read_totals(self):
        """(n_word, n_super, n_node, n_edge, n_pair) from the pinned copy; the caller has synchronised on the event
        recorded after enqueue_build."""
        B = self.dtb.n_graphs
        th = self.totals_host.tolist()
        status = th[5 * (B + 1)]
        if status != 0:
            _lib.check(status)
        return tuple(th[i * (B + 1) + B] for i in range(5))

    def batch(self, totals):
        """HeteroBatch views of the actual sizes over the slot (cached per (generation, build key, totals))."""
        key = (self.generation, id(self.host), totals)
        hb = self._batches.get(key)
        if hb is not None:
            return hb
        n_word, n_super, n_node, n_edge, n_pair = totals
        dtb, a = self.dtb, self.arr
        B = dtb.n_graphs
        offs = self.meta[:5 * (B + 1)].view(5, B + 1)
        hb = HeteroBatch(
            n_graphs=B, n_word=n_word, n_super=n_super, n_pair=n_pair,
            word_ptr=offs[0], super_ptr=offs[1], node_ptr=offs[2], edge_ptr=offs[3], pair_ptr=offs[4],
            word_wid=a["word_wid"][:n_word], word_nid=a["word_nid"][:n_word], super_nid=a["super_nid"][:n_super],
            super_type=a["super_type"][:n_super].view(torch.int8), super_graph=a["super_graph"][:n_super],
            super_extra=a["super_extra"][:n_super], super_indptr=a["super_indptr"][:n_super + 1],
            super_src=a["super_src"][:max(n_pair, 1)][:n_pair] if n_pair else a["super_src"][:0],
            super_bin=a["super_bin"][:n_pair], super_eid=a["super_eid"][:n_pair],
            word_indptr=a["word_indptr"][:n_word + 1], word_src=a["word_src"][:n_pair], word_bin=a["word_bin"][:n_pair],
            word_eid=a["word_eid"][:n_pair], n_total_nodes=n_node, n_total_edges=n_edge)
        # zero-size slices of a live buffer keep a valid base pointer, which is what the C side expects for empty sets
        hb._keepalive = (dtb, self)
        hb.labels = dtb.labels
        hb.graph_sent_ptr = dtb.graph_sent_ptr
        if dtb.hdsg and dtb.n_sent > 0:
            S = dtb.n_sent
            hb.sent_row = a["sent_row"][:S]
            hb.sent_doc_row = self.sent_doc_row[:S]
            if dtb.doc_graph is not None:
                hb.doc_row = self.doc_row[:dtb.doc_graph.shape[0]]
                hb.sent_doc_gidx, hb.doc_graph = dtb.sent_doc_g, dtb.doc_graph32
        if len(self._batches) > 64:
            self._batches.clear()
        self._batches[key] = hb
        return hb


class GraphedTrainStep:
    """loss_host, logits, d_sent_feature = step(next_host, sent_feature): one training step of `path_model.HSGPath`
    (forward, the reference's loss, backward, optional all-reduce, Adam with zero_grad) on the batch staged by the
    PREVIOUS call, while `next_host` is uploaded and built for the next call.  See the module docstring.

        gs = GraphedTrainStep(model, opt, filter_bitmap_dev, n_graphs_global)
        gs.prime(host_0)                               # upload + build of the first batch (eager)
        for i in ...:
            loss_h, logits, d_sf = gs.step(host_{i+1}, sent_feature_i)      # sent_feature: device tensor or pinned host
            ... loss_h is a pinned host scalar, valid after gs.sync_loss() (or the next step's return)

    model.loop.fuse_grad_accumulation must be on a dist.FlatGradArena whose arena is `opt.g`."""

    def __init__(self, model, opt, filter_bitmap_dev, n_graphs_global=None, all_reduce=None, capture=True,
                 resident_tokens=False, max_graphs=16, fused_reduce=None):
        from .path_model import FusedTrainStep
        self.model, self.opt = model, opt
        self.dev = opt.p.device
        self.fused = FusedTrainStep(model, n_graphs_global)
        self.bitmap = filter_bitmap_dev
        self.all_reduce = all_reduce
        self.fused_reduce = fused_reduce              # dist.PeerAllReduceAdam: replaces all_reduce + opt.step_dev
        self.capture = capture
        self.resident_tokens = resident_tokens        # True: the token blob is not re-copied per step (bench `value` leg)
        self.max_graphs = max_graphs
        self.slots = [StaticBatchSlot(self.dev), StaticBatchSlot(self.dev)]
        self.stage = [None, None]                     # pinned host staging of each slot's token blob
        self.sf_dev = None
        self.sf_stage = [None, None]
        # sent_feature of the NEXT step uploaded by this step's side branch (step(..., next_sent_feature=...)):
        # pinned staging + device buffer per slot parity, and what each device buffer currently holds
        self.sf_pref = [None, None]
        self.sf_pstage = [None, None]
        self._pref_tag = [None, None]
        self.loss_host = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
        self.side = torch.cuda.Stream(self.dev)
        self.cap_stream = torch.cuda.Stream(self.dev)
        self.ev_totals = [torch.cuda.Event(external=True), torch.cuda.Event(external=True)]
        self.ev_done = [torch.cuda.Event(external=True), torch.cuda.Event(external=True)]
        self.graphs = {}
        self.pool = None
        self.i = 0
        self.cur = None                               # (HeteroBatch, shape key) of the batch to compute next
        self.eager_steps = 0
        self.replays = 0
        self.launches_per_step = None
        # where the build branch forks off the main branch: "start", "after_forward" or "after_head" (see _enqueue)
        import os
        # Default (measured on the 32-graph step, gpurun r02v, three runs each): with the tokens resident the builder starts
        # at the top of the step (0.590 against 0.606 ms); when the branch begins with the H2D copy of the next batch the
        # builder would start ~25 us later and sit on both word-side FFN products of the forward - forking after the
        # forward is then faster end to end (53.4 k against 52.4 k graphs/s from host buffers)
        self.fork_at = os.environ.get("HSG_BUILD_FORK", "start" if resident_tokens else "after_forward")
        if self.fork_at not in ("start", "after_forward", "after_head", "split"):
            raise ValueError("HSG_BUILD_FORK must be start, after_forward, after_head or split")
        # dropout under replay: masks are keyed by the optimizer's device step counter
        model.loop.seed_dev = opt.device_step_counter()

    def _invalidate(self):
        """drop every captured graph (a fixed buffer moved, or the cache is full); their memory pool dies with them"""
        self.graphs.clear()
        self.pool = None

    # -- staging --------------------------------------------------------------------------------------------------
    def _stage_blob(self, slot, host):
        nb = host["blob"].numel()
        buf = self.stage[slot]
        if buf is None or buf.numel() < nb:
            buf = torch.empty(int(nb * StaticBatchSlot.GROW) + 64, dtype=torch.uint8).pin_memory()
            self.stage[slot] = buf
            self._invalidate()
        buf[:nb].copy_(host["blob"])                  # host memcpy into the fixed pinned staging buffer
        return buf

    def _stage_sf(self, sent_feature):
        n = sent_feature.shape[0]
        if self.sf_dev is None or self.sf_dev.shape[0] < n:
            cap = int(n * StaticBatchSlot.GROW) + 8
            self.sf_dev = torch.zeros(cap, sent_feature.shape[1], dtype=torch.float32, device=self.dev)
            self.sf_stage = [torch.zeros(cap, sent_feature.shape[1], dtype=torch.float32).pin_memory() for _ in (0, 1)]
            self._invalidate()
        if sent_feature.is_cuda:
            if sent_feature.data_ptr() != self.sf_dev.data_ptr():
                self.sf_dev[:n].copy_(sent_feature)   # stream-ordered device copy (outside the graph)
            return False
        self.sf_stage[self.i & 1][:n].copy_(sent_feature)   # host memcpy; the H2D is part of the step
        return True

    def _stage_next_sf(self, q, nsf):
        """host memcpy of the next step's sent_feature into the pinned staging buffer of parity q"""
        n, w = nsf.shape
        if self.sf_pref[q] is None or self.sf_pref[q].shape[0] < n or self.sf_pref[q].shape[1] != w:
            cap = int(n * StaticBatchSlot.GROW) + 8
            for k in (0, 1):
                self.sf_pref[k] = torch.zeros(cap, w, dtype=torch.float32, device=self.dev)
                self.sf_pstage[k] = torch.zeros(cap, w, dtype=torch.float32).pin_memory()
            self._pref_tag = [None, None]
            self._invalidate()
        self.sf_pstage[q][:n].copy_(nsf)

    def prime(self, host):
        """Upload and build the first batch eagerly into slot 0."""
        s = 0
        slot = self.slots[s]
        bkey = slot.bind(host, self.bitmap)
        stage = self._stage_blob(s, host)
        main = torch.cuda.current_stream(self.dev)
        slot.enqueue_build(main, stage)
        slot.loaded = host
        main.synchronize()
        totals = slot.read_totals()
        self.cur = (slot.batch(totals), (bkey, totals))
        self.i = 0

    # -- the step -------------------------------------------------------------------------------------------------
    def _enqueue(self, p, batch, n_sf, sf_h2d, build_next, main, use_pref=False, n_next_sf=0):
        """Everything of one step on `main` (current stream) + the side branch; returns (loss, logits, d_sf)."""
        side = self.side
        nxt = self.slots[1 - p]

        def fork_build():
            # side branch: [H2D of the next step's sent_feature], [H2D of its token blob], build.  WHERE it forks off
            # the main branch is a knob (HSG_BUILD_FORK): the builder is one CTA per graph for ~130 us and, started at
            # the top of the step, shares the GPU with the two word-side FFN products of the forward (33 us alone, 50 /
            # 63 us under it - CUPTI timeline r02t).  With resident tokens forking after the forward or after the loss is
            # NOT faster (0.590 ms at the top, 0.606 / 0.605 ms later - gpurun r02v); from host buffers it is (see the
            # default in __init__).  "split": H2D copies at the top, builder after the forward.
            if build_next or n_next_sf:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    if not copied[0]:
                        fork_copy_body()
                    if build_next:
                        nxt.enqueue_build(side, None)
                        self.ev_totals[1 - p].record(side)

        copied = [False]

        def fork_copy_body():
            # (on the side stream) the next step's sent_feature travels first: ev_totals then also covers it
            if n_next_sf:
                self.sf_pref[1 - p][:n_next_sf].copy_(self.sf_pstage[1 - p][:n_next_sf], non_blocking=True)
            if build_next and not self.resident_tokens:
                blob = nxt.dtb._blob
                blob.copy_(self.stage[1 - p][:blob.numel()], non_blocking=True)
            copied[0] = True

        def fork_copy():
            # "split": the H2D copies leave at the top of the step (copy engine, no SM), the builder forks later
            if (build_next and not self.resident_tokens) or n_next_sf:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    fork_copy_body()

        hooks = {}
        if self.fork_at == "start":
            fork_build()
        elif self.fork_at == "split":
            fork_copy()
            hooks["after_forward"] = fork_build
        else:
            hooks[self.fork_at] = fork_build
        sf = self.sf_pref[p][:n_sf] if use_pref else self.sf_dev[:n_sf]
        if sf_h2d:
            sf.copy_(self.sf_stage[p][:n_sf], non_blocking=True)
        loss, logits, d_sf = self.fused(batch, sf, hooks)
        if self.fused_reduce is not None:                  # all-reduce + Adam + zero_grad in one kernel over peer memory
            self.fused_reduce.step()
        else:
            if self.all_reduce is not None:
                self.all_reduce(self.opt.g)
            self.opt.step_dev(zero_grad=True)
        self.loss_host[p].copy_(loss.detach().view(1), non_blocking=True)
        if build_next or n_next_sf:
            main.wait_stream(side)
        self.ev_done[p].record(main)
        return loss, logits, d_sf

    def step(self, next_host, sent_feature, next_sent_feature=None):
        """next_sent_feature (optional, pinned or pageable HOST tensor): the sent_feature the NEXT call will pass; it is
        uploaded by this step's side branch next to the next batch's tokens, so that the next step does not start with
        an H2D copy on its critical path.  It must not be modified before that call, which has to pass the very same
        tensor as `sent_feature`."""
        if self.cur is None:
            raise RuntimeError("GraphedTrainStep: call prime(host) with the first batch before step()")
        p = self.i & 1
        batch, ckey = self.cur
        lib = _lib.load()
        build_next = next_host is not None
        nkey = None
        if build_next:
            nslot = self.slots[1 - p]
            nkey = nslot.bind(next_host, self.bitmap)
            if not self.resident_tokens:
                self._stage_blob(1 - p, next_host)
            elif nslot.loaded is not next_host:       # resident mode: the blob is copied once, outside the step
                nslot.dtb._blob.copy_(next_host["blob"], non_blocking=True)
                nslot.loaded = next_host
        n_sf = sent_feature.shape[0]
        tag = (sent_feature.data_ptr(), tuple(sent_feature.shape))
        use_pref = (not sent_feature.is_cuda) and self._pref_tag[p] == tag
        sf_h2d = False if use_pref else self._stage_sf(sent_feature)
        n_next_sf = 0
        if next_sent_feature is not None and not next_sent_feature.is_cuda and build_next:
            self._stage_next_sf(1 - p, next_sent_feature)
            n_next_sf = int(next_sent_feature.shape[0])
            self._pref_tag[1 - p] = (next_sent_feature.data_ptr(), tuple(next_sent_feature.shape))
        else:
            self._pref_tag[1 - p] = None
        gens = (self.slots[0].generation, self.slots[1].generation)
        key = (p, ckey, nkey, n_sf, sf_h2d, gens, use_pref, n_next_sf)
        main = torch.cuda.current_stream(self.dev)
        ent = self.graphs.get(key) if self.capture else None
        if ent is None:
            l0 = lib.hsg_launch_count()
            if self.capture and self.eager_steps >= 1:
                # capture (no work is executed), then replay
                if len(self.graphs) >= self.max_graphs:
                    self._invalidate()
                g = torch.cuda.CUDAGraph()
                cs = self.cap_stream
                cs.wait_stream(main)
                with torch.cuda.graph(g, pool=self.pool, stream=cs, capture_error_mode="thread_local"):
                    outs = self._enqueue(p, batch, n_sf, sf_h2d, build_next, cs, use_pref, n_next_sf)
                if self.pool is None:
                    self.pool = g.pool()
                main.wait_stream(cs)
                ent = (g, outs, batch)
                self.graphs[key] = ent
                self.launches_per_step = lib.hsg_launch_count() - l0
            else:
                outs = self._enqueue(p, batch, n_sf, sf_h2d, build_next, main, use_pref, n_next_sf)
                self.eager_steps += 1
                self.launches_per_step = lib.hsg_launch_count() - l0
        if ent is not None:
            ent[0].replay()
            outs = ent[1]
            self.replays += 1
        # the next batch's totals: produced early in the step by the side branch
        if build_next:
            self.ev_totals[1 - p].synchronize()
            totals = self.slots[1 - p].read_totals()
            self.cur = (self.slots[1 - p].batch(totals), (nkey, totals))
        else:
            self.cur = None
        self.i += 1
        self._last = p
        return (self.loss_host[p],) + tuple(outs[1:])

    def sync_loss(self):
        """Block until the last step has finished; returns its loss (python float)."""
        self.ev_done[self._last].synchronize()
        return float(self.loss_host[self._last])
