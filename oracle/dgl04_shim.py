"""DGL-0.4-semantics shim (TEST INFRASTRUCTURE - not part of the product path).

The reference (yellow-binary-tree/HeterSumGraph) runs its WSWGAT path on DGL 0.4
(README.md:15).  DGL is not installed in this image and cannot be installed, and
its source is not under /root/reference.  This module restates, in plain
PyTorch-on-CPU, exactly the DGL-0.4 behaviour the reference's call sites depend
on, so that

  * the reference's own module/GAT*.py, HiGraph.py and module/dataloader.py can
    be executed *verbatim* in the build container (tests/golden/make_golden.py),
  * oracle/wswgat_ref.py and oracle/graph_builder_ref.py can be pinned against
    those runs.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import anything under oracle/.

Rules reproduced (reference call site that depends on each rule in brackets):
  * filter_nodes / filter_edges evaluate the predicate over ALL nodes / edges
    and return ascending int64 ids          [module/GATLayer.py:70-71,105-107,143-145]
  * writing a NEW column on a subset of rows creates it initializer(zero)-filled
    for every other row; writes are out-of-place (autograd-safe)
                                            [module/GATLayer.py:73,111,147; HiGraph.py:149-151]
  * apply_edges(f, edges=ids) runs f on exactly those edges [module/GATLayer.py:74,112,148]
  * pull(v, msg, red): ALL in-edges of v (no type filter), degree-bucketed
    mailboxes [n_D, D, .], in-degree-0 nodes skipped [module/GATLayer.py:75,113,149]
  * add_nodes / add_edges / add_edge: ids = insertion order, missing fields are
    initializer-filled                      [module/dataloader.py:214-216,235-263,348-400]
  * batch / unbatch / sum_nodes / predecessors
                                            [module/dataloader.py:480; HiGraph.py:237,248; train.py:118]
"""
import sys
import types

import torch

__all__ = ["DGLGraph", "batch", "unbatch", "sum_nodes", "install"]


def _zero_init(shape, dtype, device=None, id_range=None):
    return torch.zeros(shape, dtype=dtype, device=device)


class _Init(types.SimpleNamespace):
    pass


init = _Init(zero_initializer=_zero_init)


def _to_index(ids):
    """DGL utils.toindex: int / list / tensor -> 1-D int64 tensor."""
    if isinstance(ids, torch.Tensor):
        return ids.reshape(-1).long()
    if isinstance(ids, (list, tuple, range)):
        return torch.as_tensor(list(ids), dtype=torch.int64).reshape(-1)
    return torch.tensor([int(ids)], dtype=torch.int64)


class _Frame:
    """Column store with an initializer for rows never written."""

    def __init__(self, n=0):
        self.n = n
        self.cols = {}
        self.initializer = _zero_init

    def add_rows(self, k, data=None):
        data = data or {}
        new_n = self.n + k
        for key in set(self.cols) | set(data):
            if key in self.cols:
                old = self.cols[key]
                if key in data:
                    add = torch.as_tensor(data[key])
                    if add.dim() == 0:
                        add = add.reshape(1)
                    add = add.to(old.dtype)
                    if add.shape[0] != k:
                        add = add.expand(k, *old.shape[1:])
                else:
                    add = self.initializer((k,) + tuple(old.shape[1:]), old.dtype, old.device)
                self.cols[key] = torch.cat([old, add.reshape((k,) + tuple(old.shape[1:]))], 0)
            else:
                add = torch.as_tensor(data[key])
                if add.dim() == 0:
                    add = add.reshape(1)
                if add.shape[0] != k:
                    add = add.expand(k, *add.shape[1:])
                head = self.initializer((self.n,) + tuple(add.shape[1:]), add.dtype, add.device)
                self.cols[key] = torch.cat([head, add], 0)
        self.n = new_n

    def set_rows(self, ids, key, val):
        val = torch.as_tensor(val) if not isinstance(val, torch.Tensor) else val
        if key not in self.cols:
            base = self.initializer((self.n,) + tuple(val.shape[1:]), val.dtype, val.device)
        else:
            base = self.cols[key]
        self.cols[key] = base.index_copy(0, ids.to(base.device), val.to(base.dtype))

    def set_col(self, key, val):
        val = torch.as_tensor(val) if not isinstance(val, torch.Tensor) else val
        assert val.shape[0] == self.n, (key, val.shape, self.n)
        self.cols[key] = val


class _DataView:
    """g.ndata / g.edata / g.nodes[ids].data / g.edges[ids].data."""

    def __init__(self, frame, ids=None):
        self._f = frame
        self._ids = ids

    def __getitem__(self, key):
        col = self._f.cols[key]
        return col if self._ids is None else col[self._ids.to(col.device)]

    def __setitem__(self, key, val):
        if self._ids is None:
            self._f.set_col(key, val)
        else:
            self._f.set_rows(self._ids, key, val)

    def __contains__(self, key):
        return key in self._f.cols

    def keys(self):
        return self._f.cols.keys()

    def pop(self, key):
        return self._f.cols.pop(key)


class _Indexer:
    def __init__(self, frame):
        self._f = frame

    def __getitem__(self, ids):
        return types.SimpleNamespace(data=_DataView(self._f, _to_index(ids)))


class _NodeBatch:
    def __init__(self, g, ids, mailbox=None):
        self.data = _DataView(g._nf, ids)
        self.mailbox = mailbox
        self._ids = ids

    def nodes(self):
        return self._ids


class _EdgeBatch:
    def __init__(self, g, eids):
        self.src = _DataView(g._nf, g._src_t()[eids])
        self.dst = _DataView(g._nf, g._dst_t()[eids])
        self.data = _DataView(g._ef, eids)


class DGLGraph:
    def __init__(self):
        self._nf = _Frame()
        self._ef = _Frame()
        self._src = []
        self._dst = []
        self._cache = None
        self.batch_size = 1
        self.batch_num_nodes = None
        self.batch_num_edges = None

    # ---- structure -------------------------------------------------------
    def number_of_nodes(self):
        return self._nf.n

    def number_of_edges(self):
        return self._ef.n

    def _src_t(self):
        self._build_cache()
        return self._cache[0]

    def _dst_t(self):
        self._build_cache()
        return self._cache[1]

    def _build_cache(self):
        if self._cache is None:
            self._cache = (torch.as_tensor(self._src, dtype=torch.int64).reshape(-1),
                           torch.as_tensor(self._dst, dtype=torch.int64).reshape(-1))

    def add_nodes(self, num, data=None):
        self._nf.add_rows(int(num), data)

    def add_edges(self, u, v, data=None):
        u, v = _to_index(u), _to_index(v)
        k = max(len(u), len(v))
        if len(u) != k:
            u = u.expand(k)
        if len(v) != k:
            v = v.expand(k)
        self._src.extend(u.tolist())
        self._dst.extend(v.tolist())
        self._cache = None
        self._ef.add_rows(k, data)

    def add_edge(self, u, v, data=None):
        self.add_edges(int(u), int(v), data)

    def set_n_initializer(self, initializer, field=None):
        self._nf.initializer = initializer

    def set_e_initializer(self, initializer, field=None):
        self._ef.initializer = initializer

    def edges_arrays(self):
        return self._src_t(), self._dst_t()

    def predecessors(self, v):
        v = int(v)
        return self._src_t()[self._dst_t() == v]

    def in_degrees(self):
        return torch.bincount(self._dst_t(), minlength=self._nf.n)

    # ---- frames ----------------------------------------------------------
    @property
    def ndata(self):
        return _DataView(self._nf)

    @property
    def edata(self):
        return _DataView(self._ef)

    @property
    def nodes(self):
        return _Indexer(self._nf)

    @property
    def edges(self):
        return _Indexer(self._ef)

    def to(self, device):
        for f in (self._nf, self._ef):
            for k in list(f.cols):
                f.cols[k] = f.cols[k].to(device)
        return self

    # ---- UDF runtime -----------------------------------------------------
    def filter_nodes(self, predicate, nodes=None):
        ids = torch.arange(self._nf.n)
        mask = predicate(_NodeBatch(self, ids))
        return ids[mask.reshape(-1).cpu()]

    def filter_edges(self, predicate, edges=None):
        ids = torch.arange(self._ef.n)
        mask = predicate(_EdgeBatch(self, ids))
        return ids[mask.reshape(-1).cpu()]

    def apply_edges(self, func, edges=None):
        eids = torch.arange(self._ef.n) if edges is None else _to_index(edges)
        out = func(_EdgeBatch(self, eids))
        for k, val in out.items():
            self._ef.set_rows(eids, k, val)

    def pull(self, v, message_func, reduce_func, apply_node_func=None):
        v = _to_index(v)
        if len(v) == 0:
            return
        src, dst = self._src_t(), self._dst_t()
        is_pull = torch.zeros(self._nf.n, dtype=torch.bool)
        is_pull[v] = True
        eids = torch.nonzero(is_pull[dst]).reshape(-1)       # all in-edges, ascending edge id
        if len(eids) == 0:
            return
        msgs = message_func(_EdgeBatch(self, eids))
        e_dst = dst[eids]
        order = torch.argsort(e_dst, stable=True)              # per destination, ascending edge id
        e_dst_sorted = e_dst[order]
        uniq, counts = torch.unique_consecutive(e_dst_sorted, return_counts=True)
        starts = torch.cumsum(counts, 0) - counts
        results = {}
        out_nodes = []
        for deg in torch.unique(counts).tolist():              # degree bucketing
            sel = torch.nonzero(counts == deg).reshape(-1)
            nodes_b = uniq[sel]
            pos = (starts[sel].reshape(-1, 1) + torch.arange(deg).reshape(1, -1)).reshape(-1)
            rows = order[pos]
            mailbox = {k: m[rows].reshape((len(sel), deg) + tuple(m.shape[1:])) for k, m in msgs.items()}
            red = reduce_func(_NodeBatch(self, nodes_b, mailbox=mailbox))
            out_nodes.append(nodes_b)
            for k, val in red.items():
                results.setdefault(k, []).append(val)
        out_nodes = torch.cat(out_nodes)
        for k, vals in results.items():
            self._nf.set_rows(out_nodes, k, torch.cat(vals, 0))


def batch(graph_list):
    bg = DGLGraph()
    n_off = 0
    nn_, ne_ = [], []
    for g in graph_list:
        s, d = g.edges_arrays()
        bg._src.extend((s + n_off).tolist())
        bg._dst.extend((d + n_off).tolist())
        n_off += g.number_of_nodes()
        nn_.append(g.number_of_nodes())
        ne_.append(g.number_of_edges())
    for frame, attr in ((bg._nf, "_nf"), (bg._ef, "_ef")):
        keys = list(getattr(graph_list[0], attr).cols.keys())
        for k in keys:
            frame.cols[k] = torch.cat([getattr(g, attr).cols[k] for g in graph_list], 0)
    bg._nf.n = sum(nn_)
    bg._ef.n = sum(ne_)
    bg.batch_size = len(graph_list)
    bg.batch_num_nodes = nn_
    bg.batch_num_edges = ne_
    return bg


def unbatch(bg):
    out = []
    n_off = e_off = 0
    src, dst = bg.edges_arrays()
    for nn_, ne_ in zip(bg.batch_num_nodes, bg.batch_num_edges):
        g = DGLGraph()
        g._src = (src[e_off:e_off + ne_] - n_off).tolist()
        g._dst = (dst[e_off:e_off + ne_] - n_off).tolist()
        g._nf.n, g._ef.n = nn_, ne_
        for k, col in bg._nf.cols.items():
            g._nf.cols[k] = col[n_off:n_off + nn_]
        for k, col in bg._ef.cols.items():
            g._ef.cols[k] = col[e_off:e_off + ne_]
        out.append(g)
        n_off += nn_
        e_off += ne_
    return out


def sum_nodes(bg, feat):
    col = bg._nf.cols[feat]
    seg = torch.repeat_interleave(torch.arange(bg.batch_size), torch.as_tensor(bg.batch_num_nodes))
    out = torch.zeros((bg.batch_size,) + tuple(col.shape[1:]), dtype=col.dtype, device=col.device)
    return out.index_add(0, seg.to(col.device), col)


def install():
    """Register this shim as `dgl` (and stub nltk stopwords) so /root/reference imports."""
    mod = types.ModuleType("dgl")
    mod.DGLGraph = DGLGraph
    mod.batch = batch
    mod.unbatch = unbatch
    mod.sum_nodes = sum_nodes
    mod.init = init
    data = types.ModuleType("dgl.data")
    utils = types.ModuleType("dgl.data.utils")
    utils.save_graphs = lambda *a, **k: None
    utils.load_graphs = lambda *a, **k: ([], {})
    data.utils = utils
    mod.data = data
    sys.modules["dgl"] = mod
    sys.modules["dgl.data"] = data
    sys.modules["dgl.data.utils"] = utils
    if "nltk" not in sys.modules:
        nltk = types.ModuleType("nltk")
        corpus = types.ModuleType("nltk.corpus")
        corpus.stopwords = types.SimpleNamespace(words=lambda lang: [])
        nltk.corpus = corpus
        nltk.FreqDist = dict
        sys.modules["nltk"] = nltk
        sys.modules["nltk.corpus"] = corpus
    return mod
