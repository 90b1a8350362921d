"""Host ETL (hetersumgraph_b200/etl.py) against the reference's own ExampleSet / MultiExampleSet run on the DGL-0.4
shim: same files in, and the graph the reference builds must equal the graph the oracle builder derives from the
ETL's DocExample (node ids, word ids, edge lists, TF-IDF boxes, labels)."""
import json
import os
import sys

import numpy as np
import pytest

from hetersumgraph_b200 import etl
from hetersumgraph_b200 import synthetic as syn

STOP = ["the", "a", "of", "and", "is", "in", "to"]
WORDS = ["w%d" % i for i in range(60)]


def _write_files(tmp, multi):
    rng = np.random.default_rng(7)
    vocab_path = os.path.join(tmp, "vocab")
    with open(vocab_path, "w") as f:
        for w in STOP + WORDS[:50] + [",", "."]:
            f.write("%s\t%d\n" % (w, 10))
    with open(os.path.join(tmp, "filter_word.txt"), "w") as f:
        for w in ["w3", "w7", "zzz-not-in-vocab", "w11"]:
            f.write(w + "\n")

    def sent():
        n = int(rng.integers(3, 12))
        toks = [str(rng.choice(STOP + WORDS + [",", "."])) for _ in range(n)]
        if rng.random() < 0.5:
            toks[0] = toks[0].upper() if toks[0].startswith("w") else toks[0]    # lower-casing at lookup
        return " ".join(toks)

    data, w2s, w2d = [], [], []
    for _ in range(5):
        if multi:
            docs = [[sent() for _ in range(int(rng.integers(1, 4)))] for _ in range(int(rng.integers(2, 4)))]
            sents = [s for d in docs for s in d]
            text = docs
        else:
            sents = [sent() for _ in range(int(rng.integers(2, 9)))]
            text = sents
        label = sorted(rng.choice(len(sents), size=min(2, len(sents)), replace=False).tolist())
        data.append({"text": text, "summary": ["x"], "label": label})
        # TF-IDF files keyed by WORD (lower-cased as sklearn does); some words missing, some OOV keys
        w2s.append({str(i): {w.lower(): float(rng.random()) for w in s.split() if rng.random() < 0.8 and len(w) > 1}
                    for i, s in enumerate(sents)})
        if multi:
            w2d.append({str(j): {w.lower(): float(rng.random()) for s in d for w in s.split() if rng.random() < 0.7 and len(w) > 1}
                        for j, d in enumerate(docs)})
    for name, rows in (("data.jsonl", data), ("w2s.jsonl", w2s), ("w2d.jsonl", w2d)):
        with open(os.path.join(tmp, name), "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    return vocab_path


def _graph_arrays_from_example(ex, filt, hdsg):
    from oracle import graph_builder_ref as gb
    if hdsg:
        return gb.create_graph_hdsg(ex.doc_len, ex.sents.tolist(), ex.doc_tokens, ex.w2s, ex.w2d, filt)
    return gb.create_graph_hsg(ex.sents.tolist(), ex.w2s, filt)


@pytest.mark.parametrize("multi", [False, True])
def test_etl_matches_live_reference_dataset(tmp_path, multi):
    if not os.path.isdir("/root/reference/module"):
        pytest.skip("reference checkout not present (GPU box)")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg                       # installs the shim, imports module.dataloader as refdl
    vocab_path = _write_files(str(tmp_path), multi)
    p = lambda n: os.path.join(str(tmp_path), n)   # noqa: E731
    mg.refdl.FILTERWORD = list(STOP) + list(etl.PUNCTUATIONS)      # what stopwords.words('english') + punctuations is
    sys.path.insert(0, "/root/reference")
    from module.vocabulary import Vocab as RefVocab
    rv = RefVocab(vocab_path, 1000)
    mine_v = etl.Vocab(vocab_path, 1000)
    assert rv.size() == mine_v.size() and all(rv.word2id(w) == mine_v.word2id(w) for w in STOP + WORDS + ["[PAD]", "q"])
    L, DM = 8, 6
    if multi:
        ref = mg.refdl.MultiExampleSet(p("data.jsonl"), rv, DM, L, p("filter_word.txt"), p("w2s.jsonl"), p("w2d.jsonl"))
        mine = etl.JsonlDataset(p("data.jsonl"), mine_v, DM, L, p("filter_word.txt"), p("w2s.jsonl"), STOP, p("w2d.jsonl"))
    else:
        ref = mg.refdl.ExampleSet(p("data.jsonl"), rv, DM, L, p("filter_word.txt"), p("w2s.jsonl"))
        mine = etl.JsonlDataset(p("data.jsonl"), mine_v, DM, L, p("filter_word.txt"), p("w2s.jsonl"), STOP)
    assert sorted(set(ref.filterids)) == sorted(set(mine.filter_ids))
    filt = set(mine.filter_ids)
    for i in range(len(mine)):
        G, _ = ref[i]
        want = mg.shim_to_arrays(G)
        ex = mine[i]
        got = _graph_arrays_from_example(ex, filt, multi)
        for k in ("src", "dst", "unit", "ndtype", "wid", "tffrac", "etype"):
            assert np.array_equal(getattr(got, k), getattr(want, k)), (i, k)
        snode = np.nonzero(want.ndtype == 1)[0]
        ref_label = G.ndata["label"][snode].sum(-1).numpy()                 # train.py:115
        assert np.array_equal(ex.labels, ref_label), i
    tb = mine.collate(list(range(len(mine))))
    n_sent = [min(len([s for d in e["text"] for s in d]) if multi else len(e["text"]), DM) for e in mine.examples]
    assert tb.order == syn.stable_desc_order(n_sent).tolist()              # graph_collate_fn order
    assert tb.tokens.shape == (sum(n_sent), L)


def test_etl_label_and_padding_rules(tmp_path):
    vocab_path = _write_files(str(tmp_path), False)
    v = etl.Vocab(vocab_path, 0)
    e = {"text": ["w1 W2 w3 w4 w5 w6 w7 w8 w9 w10", "w1", "qqq w2"], "label": [2, 0, 1, 2]}
    w2s = {"0": {"w1": 0.5, "w2": 1.0, "w10": 0.3}, "1": {}, "2": {"qqq": 0.9, "w2": 0.25}}
    ex = etl.make_doc_example(e, v, sent_max_len=4, doc_max_timesteps=2, w2s=w2s)
    assert ex.sents.shape == (2, 4)                                        # first doc_max sentences, cut / padded to 4
    assert ex.sents[0].tolist() == [v.word2id("w1"), v.word2id("w2"), v.word2id("w3"), v.word2id("w4")]
    assert ex.sents[1].tolist() == [v.word2id("w1"), 0, 0, 0]
    assert ex.labels.tolist() == [1, 0]                                    # only the first doc_max label steps count
    assert ex.w2s[0] == {v.word2id("w1"): 0.5, v.word2id("w2"): 1.0}       # w10 was truncated away; [UNK] never matches
