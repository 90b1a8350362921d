"""Host ETL: the reference's on-disk formats -> the flat TokenBatch the device-side builder consumes
(SURVEY.md §8-f rank 4).  Pure host code (no torch, no DGL, no NLTK).

Follows, line by line in behaviour,
  module/vocabulary.py:33-93          Vocab: [PAD]=0, [UNK]=1, [START]=2, [STOP]=3, then "<word>\\t<freq>" lines
  module/dataloader.py:56-137         Example / Example2: whitespace tokens, lower-cased lookup, pad/truncate to
                                      sent_max_len, label matrix; Example2.enc_doc_input = UNtruncated ids per document
  module/dataloader.py:142-183        ExampleSet.__init__: filter ids = stop words + punctuation + [PAD] + the first
                                      5 001 in-vocabulary words of the low-TF-IDF file
  module/dataloader.py:270-283,408-423  __getitem__: first doc_max_timesteps sentences, label matrix cut to
                                      [:doc_max, :doc_max], per-example TF-IDF dictionaries keyed by WORD
  module/dataloader.py:472-481        graph_collate_fn: stable sort by #sentences descending  (in pack_token_batch)

The reference takes its stop words from nltk.corpus.stopwords (dataloader.py:48), which is not shipped here: pass the
same list as `stopwords`.
"""
import json
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from .synthetic import DocExample, TokenBatch, pack_token_batch

PAD_TOKEN, UNKNOWN_TOKEN, START_DECODING, STOP_DECODING = "[PAD]", "[UNK]", "[START]", "[STOP]"
PUNCTUATIONS = [',', '.', ':', ';', '?', '(', ')', '[', ']', '&', '!', '*', '@', '#', '$', '%', "''", "'", '`', '``',
                '-', '--', '|', '\\/']          # dataloader.py:49-50


class Vocab:
    """module/vocabulary.py:33-93 (same ids, same max_size cut, duplicate lines skipped)."""

    def __init__(self, vocab_file: str, max_size: int):
        self._w2i: Dict[str, int] = {}
        self._i2w: Dict[int, str] = {}
        for w in (PAD_TOKEN, UNKNOWN_TOKEN, START_DECODING, STOP_DECODING):
            self._add(w)
        with open(vocab_file, "r", encoding="utf8") as f:
            for line in f:
                w = line.split("\t")[0]
                if w in (UNKNOWN_TOKEN, PAD_TOKEN, START_DECODING, STOP_DECODING):
                    raise Exception("[UNK], [PAD], [START] and [STOP] shouldn't be in the vocab file, but %s is" % w)
                if w in self._w2i:
                    continue
                self._add(w)
                if max_size != 0 and len(self._w2i) >= max_size:
                    break

    def _add(self, w):
        i = len(self._w2i)
        self._w2i[w] = i
        self._i2w[i] = w

    def word2id(self, word: str) -> int:
        return self._w2i.get(word, self._w2i[UNKNOWN_TOKEN])

    def id2word(self, word_id: int) -> str:
        if word_id not in self._i2w:
            raise ValueError("Id not found in vocab: %d" % word_id)
        return self._i2w[word_id]

    def size(self) -> int:
        return len(self._w2i)


def read_jsonl(path: str) -> List[dict]:
    with open(path, encoding="utf-8") as f:
        return [json.loads(line) for line in f]


def build_filter_ids(vocab: Vocab, filter_word_path: str, stopwords: Iterable[str]) -> List[int]:
    """dataloader.py:167-182."""
    words = list(stopwords) + PUNCTUATIONS
    ids = [vocab.word2id(w.lower()) for w in words]
    ids.append(vocab.word2id(PAD_TOKEN))
    n = 0
    unk = vocab.word2id(UNKNOWN_TOKEN)
    with open(filter_word_path, encoding="utf-8") as f:
        for line in f:
            w = line.strip()
            if vocab.word2id(w) != unk:
                ids.append(vocab.word2id(w))
                n += 1
            if n > 5000:
                break
    return ids


def _by_id(padded_ids: Sequence[int], tfw: Dict[str, float], vocab: Vocab) -> Dict[int, float]:
    """{word: tfidf} -> {wid: tfidf} the way CreateGraph looks it up: id -> id2word(id) -> key (dataloader.py:251-252),
    so out-of-vocabulary words ([UNK]) never match."""
    out = {}
    for wid in dict.fromkeys(int(w) for w in padded_ids):
        w = vocab.id2word(wid)
        if w in tfw:
            out[wid] = float(tfw[w])
    return out


def make_doc_example(e: dict, vocab: Vocab, sent_max_len: int, doc_max_timesteps: int, w2s: Dict[str, Dict[str, float]],
                     w2d: Optional[Dict[str, Dict[str, float]]] = None) -> DocExample:
    """One JSONL record ({"text": [...] or [[...], ...], "label": [...]}) -> DocExample (Example / Example2 +
    ExampleSet.__getitem__ / MultiExampleSet.__getitem__)."""
    text = e["text"]
    multi = isinstance(text, list) and len(text) > 0 and isinstance(text[0], list)
    sents = [s for doc in text for s in doc] if multi else list(text)
    ids_raw = [[vocab.word2id(w.lower()) for w in s.split()] for s in sents]
    pad = vocab.word2id(PAD_TOKEN)
    padded = [(r[:sent_max_len] + [pad] * max(0, sent_max_len - len(r))) for r in ids_raw]
    n_all = len(sents)
    n = min(n_all, doc_max_timesteps)
    label = list(e.get("label", []))
    lab = np.zeros(n_all, np.int64)
    for step, sent_no in enumerate(label):                 # label_matrix[label[j], j] = 1, columns cut at doc_max
        if step < doc_max_timesteps:
            lab[sent_no] += 1
    arr = np.asarray(padded[:n], np.int32).reshape(n, sent_max_len)
    ex = DocExample(sents=arr, w2s=[_by_id(padded[i], w2s[str(i)], vocab) for i in range(n)], labels=lab[:n])
    if multi:
        cur, doc_len, doc_tokens = 0, [], []
        for doc in text:
            if len(doc) == 0:
                continue
            doc_len.append(len(doc))
            doc_tokens.append([t for r in ids_raw[cur:cur + len(doc)] for t in r])     # Example2.enc_doc_input
            cur += len(doc)
        ex.doc_len, ex.doc_tokens = doc_len, doc_tokens
        ex.w2d = [_by_id(doc_tokens[j], (w2d or {}).get(str(j), {}), vocab) for j in range(len(doc_tokens))]
    return ex


class JsonlDataset:
    """ExampleSet / MultiExampleSet without the graph construction (that happens on the device): item i is the
    DocExample of record i; collate() is graph_collate_fn up to the TokenBatch."""

    def __init__(self, data_path: str, vocab: Vocab, doc_max_timesteps: int, sent_max_len: int, filter_word_path: str,
                 w2s_path: str, stopwords: Iterable[str], w2d_path: Optional[str] = None):
        self.vocab, self.doc_max_timesteps, self.sent_max_len = vocab, doc_max_timesteps, sent_max_len
        self.examples = read_jsonl(data_path)
        self.filter_ids = build_filter_ids(vocab, filter_word_path, stopwords)
        self.w2s = read_jsonl(w2s_path)
        self.w2d = read_jsonl(w2d_path) if w2d_path else None
        self.hdsg = w2d_path is not None

    def __len__(self):
        return len(self.examples)

    def __getitem__(self, i: int) -> DocExample:
        return make_doc_example(self.examples[i], self.vocab, self.sent_max_len, self.doc_max_timesteps, self.w2s[i],
                                self.w2d[i] if self.w2d is not None else None)

    def collate(self, indices: Sequence[int]) -> TokenBatch:
        exs = [self[i] for i in indices]
        return pack_token_batch(exs, hdsg=self.hdsg, vocab_size=self.vocab.size(),
                                doc_max_timesteps=self.doc_max_timesteps, filter_ids_list=self.filter_ids)
