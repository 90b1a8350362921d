"""Synthetic CNN/DM-, NYT50- and Multi-News-shaped document batches.

Token-level generator, so the graph builder is exercised exactly like the
reference's loader (module/dataloader.py:201-268, 328-406): every example is a
list of padded sentences plus per-sentence (and per-document) TF-IDF
dictionaries, computed with the formula of the reference's preprocessing
(script/calw2sTFIDF.py:28-33: sklearn CountVectorizer + TfidfTransformer, i.e.
idf = ln((1+n)/(1+df)) + 1, L2-normalised rows).

Host-side ETL only (out of scope for kernels): `pack_token_batch`
turns examples into the flat int arrays the device-side builder consumes.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

VOCAB_SIZE = 50000
SENT_MAX_LEN = 100
DOC_MAX_TIMESTEPS = 50
PAD_ID, UNK_ID = 0, 1           # module/vocabulary.py:44-47
N_STOP = 200                    # stop-word surrogate: ranks 1..200
N_LOW_TFIDF = 5000              # dataloader.py:167-182 low-TFIDF cut

SHAPES = {
    # name: (mean sents, std sents, min sents, max sents, median tokens/sent)
    "cnndm": (30.0, 12.0, 3, 50, 25.0),
    "nyt50": (40.0, 12.0, 5, 50, 30.0),
    "multinews": (45.0, 10.0, 8, 50, 28.0),
    "tiny": (8.0, 3.0, 3, 12, 12.0),          # test fixtures only
}


def filter_ids(vocab_size: int = VOCAB_SIZE) -> np.ndarray:
    """ids removed from the word-node set: PAD, 200 most frequent, 5000 rarest ranks."""
    n_rank = vocab_size - 4
    ids = [PAD_ID] + [r + 3 for r in range(1, N_STOP + 1)] + [r + 3 for r in range(n_rank - N_LOW_TFIDF + 1, n_rank + 1)]
    return np.asarray(ids, np.int64)


def filter_bitmap(vocab_size: int = VOCAB_SIZE) -> np.ndarray:
    bm = np.zeros((vocab_size + 31) // 32, np.uint32)
    ids = filter_ids(vocab_size)
    np.bitwise_or.at(bm, ids >> 5, (np.uint32(1) << (ids & 31).astype(np.uint32)))
    return bm


def is_tfidf_key(wid) -> np.ndarray:
    """Surrogate for "the token survives sklearn's tokenizer" (dataloader.py:251)."""
    wid = np.asarray(wid)
    return (wid % 16 != 0) & (wid != UNK_ID) & (wid != PAD_ID)


@dataclass
class DocExample:
    sents: np.ndarray                                   # int32 [N, L] padded with PAD_ID
    w2s: List[Dict[int, float]]                         # per sentence {wid: tfidf}
    labels: np.ndarray                                  # int64 [N] (label matrix summed over steps)
    doc_len: Optional[List[int]] = None                 # HDSG: sentences per document
    doc_tokens: Optional[List[List[int]]] = None        # HDSG: unpadded token ids per document
    w2d: Optional[List[Dict[int, float]]] = None        # HDSG: per document {wid: tfidf}

    @property
    def n_sent(self) -> int:
        return int(self.sents.shape[0])


def _tfidf(docs: List[np.ndarray]) -> List[Dict[int, float]]:
    """sklearn TfidfTransformer(norm='l2', smooth_idf=True) over `docs` (arrays of key token ids)."""
    n = len(docs)
    uniq = [np.unique(d, return_counts=True) for d in docs]
    df: Dict[int, int] = {}
    for u, _ in uniq:
        for w in u.tolist():
            df[w] = df.get(w, 0) + 1
    out = []
    for u, c in uniq:
        if len(u) == 0:
            out.append({})
            continue
        idf = np.log((1.0 + n) / (1.0 + np.asarray([df[w] for w in u.tolist()], np.float64))) + 1.0
        v = c.astype(np.float64) * idf
        v = v / np.sqrt((v * v).sum())
        out.append({int(w): float(x) for w, x in zip(u.tolist(), v.tolist())})
    return out


class _Sampler:
    def __init__(self, rng, vocab_size):
        self.rng = rng
        self.n_rank = vocab_size - 4
        w = 1.0 / np.arange(1, self.n_rank + 1, dtype=np.float64)
        self.cdf = np.cumsum(w / w.sum())

    def zipf(self, n):
        r = np.searchsorted(self.cdf, self.rng.random(n), side="left") + 1
        return np.minimum(r, self.n_rank) + 3

    def topic(self):
        return self.rng.integers(201, 20001, size=150) + 3


def make_example(rng, sampler, shape: str, hdsg: bool = False, sent_max_len: int = SENT_MAX_LEN,
                 unk_prob: float = 0.01) -> DocExample:
    mean, std, lo, hi, lmed = SHAPES[shape]
    n_s = int(np.clip(np.rint(rng.normal(mean, std)), lo, hi))
    lens = np.clip(np.rint(rng.lognormal(np.log(lmed), 0.5, size=n_s)), 3, sent_max_len).astype(np.int64)
    topic = sampler.topic()
    sents = np.zeros((n_s, sent_max_len), np.int32)
    raw = []
    for i in range(n_s):
        L = int(lens[i])
        tok = sampler.zipf(L)
        use_topic = rng.random(L) < 0.4
        tok = np.where(use_topic, topic[rng.integers(0, 150, size=L)], tok)
        tok = np.where(rng.random(L) < unk_prob, UNK_ID, tok)
        sents[i, :L] = tok
        raw.append(tok.astype(np.int64))
    w2s = _tfidf([t[is_tfidf_key(t)] for t in raw])
    labels = np.zeros(n_s, np.int64)
    labels[rng.choice(n_s, size=min(3, n_s), replace=False)] = 1
    ex = DocExample(sents=sents, w2s=w2s, labels=labels)
    if hdsg:
        n_docs = int(rng.integers(2, 6))
        n_docs = min(n_docs, n_s)
        base, rem = divmod(n_s, n_docs)
        ex.doc_len = [base + (1 if j < rem else 0) for j in range(n_docs)]
        ex.doc_tokens, cur = [], 0
        for dl in ex.doc_len:
            ex.doc_tokens.append(np.concatenate(raw[cur:cur + dl]).tolist())
            cur += dl
        ex.w2d = _tfidf([np.asarray(t, np.int64)[is_tfidf_key(np.asarray(t, np.int64))] for t in ex.doc_tokens])
    return ex


def make_examples(n_graphs: int, shape: str = "cnndm", seed: int = 0, hdsg: bool = False,
                  vocab_size: int = VOCAB_SIZE, sent_max_len: int = SENT_MAX_LEN) -> List[DocExample]:
    rng = np.random.default_rng(seed)
    sampler = _Sampler(rng, vocab_size)
    return [make_example(rng, sampler, shape, hdsg, sent_max_len) for _ in range(n_graphs)]


@dataclass
class TokenBatch:
    """Flat host arrays consumed by the device-side graph builder (hsg_build_batch).

    Graphs are already in batch order (stable sort by #sentences descending,
    reference dataloader.py:479).  `sent_bin[s,t]` is the TF-IDF bin
    round_half_even(9*tfidf) (dataloader.py:253) of token (s,t) when its word is a
    TF-IDF key of sentence s, else -1.
    """
    hdsg: bool
    order: List[int]
    tokens: np.ndarray            # int32 [S, L]
    sent_bin: np.ndarray          # int8  [S, L]
    graph_sent_ptr: np.ndarray    # int32 [B+1]
    filter_bitmap: np.ndarray     # uint32 [ceil(V/32)]
    labels: np.ndarray            # int64 [S]
    # HDSG only
    graph_doc_ptr: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))   # int32 [B+1]
    sent_doc: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))        # int32 [S] local doc index
    doc_tok_ptr: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))     # int32 [D+1]
    doc_tokens: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))      # int32 [T]
    doc_bin: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int8))          # int8  [T]

    @property
    def n_graphs(self) -> int:
        return len(self.graph_sent_ptr) - 1


def _bins_for(tokens: np.ndarray, table: Dict[int, float]) -> np.ndarray:
    out = np.full(tokens.shape, -1, np.int8)
    if table:
        keys = np.fromiter(table.keys(), np.int64, len(table))
        vals = np.rint(np.fromiter(table.values(), np.float64, len(table)) * 9.0).astype(np.int8)  # half-to-even
        o = np.argsort(keys)
        keys, vals = keys[o], vals[o]
        pos = np.searchsorted(keys, tokens)
        pos = np.minimum(pos, len(keys) - 1)
        hit = keys[pos] == tokens
        out[hit] = vals[pos[hit]]
    return out


def stable_desc_order(lengths) -> np.ndarray:
    return np.argsort(-np.asarray(lengths, np.int64), kind="stable")


def map_sent2doc(doc_len: List[int], n_sent: int) -> Dict[int, int]:
    """MultiExampleSet.MapSent2Doc (dataloader.py:314-326) incl. its early return."""
    out, no = {}, 0
    for i, dl in enumerate(doc_len):
        for _ in range(dl):
            out[no] = i
            no += 1
            if no > n_sent:
                return out
    return out


def bitmap_of(ids, vocab_size: int) -> np.ndarray:
    """bit set = filtered id (the builder's filter_bitmap input)."""
    ids = np.asarray(sorted(set(int(i) for i in ids if 0 <= int(i) < vocab_size)), np.int64)
    bm = np.zeros((vocab_size + 31) // 32, np.uint32)
    np.bitwise_or.at(bm, ids >> 5, (np.uint32(1) << (ids & 31).astype(np.uint32)))
    return bm


def pack_token_batch(examples: List[DocExample], hdsg: bool = False, vocab_size: int = VOCAB_SIZE,
                     order: Optional[List[int]] = None, doc_max_timesteps: int = DOC_MAX_TIMESTEPS,
                     filter_ids_list=None) -> TokenBatch:
    """filter_ids_list: the dataset's filter ids (dataloader.py:167-182); None = the synthetic generator's set."""
    if order is None:
        order = stable_desc_order([min(e.n_sent, doc_max_timesteps) for e in examples]).tolist()
    toks, bins, labels, sent_doc = [], [], [], []
    gsp, gdp, dtp = [0], [0], [0]
    dtoks, dbins = [], []
    for idx in order:
        e = examples[idx]
        n = min(e.n_sent, doc_max_timesteps)              # dataloader.py:278,416
        t = np.ascontiguousarray(e.sents[:n])
        toks.append(t)
        bins.append(np.stack([_bins_for(t[i], e.w2s[i]) for i in range(n)]) if n else np.zeros((0, t.shape[1]), np.int8))
        labels.append(e.labels[:n])
        gsp.append(gsp[-1] + n)
        if hdsg:
            s2d = map_sent2doc(e.doc_len, n)
            n_docs = len(set(s2d.values()))
            sent_doc.append(np.asarray([s2d[i] for i in range(n)], np.int32))
            for j in range(n_docs):
                dt = np.asarray(e.doc_tokens[j], np.int32)
                dtoks.append(dt)
                dbins.append(_bins_for(dt, e.w2d[j]))
                dtp.append(dtp[-1] + len(dt))
            gdp.append(gdp[-1] + n_docs)
    L = examples[0].sents.shape[1]
    tb = TokenBatch(
        hdsg=hdsg, order=[int(i) for i in order],
        tokens=np.concatenate(toks).astype(np.int32) if toks else np.zeros((0, L), np.int32),
        sent_bin=np.concatenate(bins).astype(np.int8) if bins else np.zeros((0, L), np.int8),
        graph_sent_ptr=np.asarray(gsp, np.int32),
        filter_bitmap=filter_bitmap(vocab_size) if filter_ids_list is None else bitmap_of(filter_ids_list, vocab_size),
        labels=np.concatenate(labels).astype(np.int64) if labels else np.zeros(0, np.int64))
    if hdsg:
        tb.graph_doc_ptr = np.asarray(gdp, np.int32)
        tb.sent_doc = np.concatenate(sent_doc).astype(np.int32)
        tb.doc_tok_ptr = np.asarray(dtp, np.int32)
        tb.doc_tokens = np.concatenate(dtoks).astype(np.int32) if dtoks else np.zeros(0, np.int32)
        tb.doc_bin = np.concatenate(dbins).astype(np.int8) if dbins else np.zeros(0, np.int8)
    return tb


def stress_edges(n_word: int = 262144, n_super: int = 32768, n_edges: int = 1048576, seed: int = 4,
                 extra: int = 64):
    """One big bipartite graph: supernode in-degree ~ Poisson(E/Ns).

    Returns (word_row[E], super_row[E], bin[E], extra_cnt[Ns]) for the active
    word<->supernode pairs, in DGL insertion order (supernode-major).
    """
    rng = np.random.default_rng(seed)
    deg = rng.poisson(n_edges / n_super, size=n_super).astype(np.int64)
    diff = n_edges - int(deg.sum())
    while diff != 0:                                      # fix the total to exactly n_edges
        k = min(abs(diff), n_super)
        idx = rng.choice(n_super, size=k, replace=False)
        if diff > 0:
            deg[idx] += 1
            diff -= k
        else:
            ok = idx[deg[idx] > 0]
            deg[ok] -= 1
            diff += len(ok)
    sup = np.repeat(np.arange(n_super, dtype=np.int64), deg)
    word = rng.integers(0, n_word, size=n_edges, dtype=np.int64)
    bins = rng.integers(0, 10, size=n_edges, dtype=np.int64)
    return word, sup, bins, np.full(n_super, extra, np.int64)
