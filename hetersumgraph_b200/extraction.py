"""Sentence extraction and label metrics at evaluation time - the counterpart of SLTester.evaluation / getMetric /
ngram_blocking (Tester.py:88-184, tools/utils.py:45-55) without the ROUGE / text-file side (out of scope).

  m == 0           prediction = argmax over the two classes                          Tester.py:120-122
  m  > 0           the m sentences with the largest class-1 logit per graph          Tester.py:127-128  (hsg_topm, device)
  blocking         greedy selection in score order, skipping a sentence that repeats an n-gram of an already selected
                   one (incl. the reference's `range(len(pieces) - n_win)` bound, which never looks at the last n-gram)
                                                                                     Tester.py:155-184  (host: text)
  counters         pred / true / match_true / match / total -> accuracy, precision, recall, F      Tester.py:132-137,
                                                                                     tools/utils.py:45-55
"""
from typing import List, Optional, Sequence

import torch

from .functional import topm as _topm
from .graph import HeteroBatch


def ngram_blocking(sents: Sequence[str], order: Sequence[int], n_win: int, k: int) -> List[int]:
    """Tester.py:155-184 given the sentences' descending-score order (`p_sent.sort(descending=True)[1]`)."""
    seen = set()
    picked: List[int] = []
    for idx in order:
        pieces = sents[idx].split()
        grams = []
        overlap = False
        for i in range(len(pieces) - n_win):              # sic: the last n-gram is never formed (Tester.py:171)
            g = " ".join(pieces[i:i + n_win])
            if g in seen:
                overlap = True
                break
            grams.append(g)
        if not overlap:
            picked.append(int(idx))
            seen.update(grams)
            if len(picked) >= k:
                break
    return picked


def eval_label(match_true, pred, true, total, match):
    """tools/utils.py:45-55 (0.0 for every metric when a denominator is zero)."""
    try:
        accu = float(match) / float(total)
        precision = float(match_true) / float(pred)
        recall = float(match_true) / float(true)
        f = 2 * precision * recall / (precision + recall)
    except ZeroDivisionError:
        accu, precision, recall, f = 0.0, 0.0, 0.0, 0.0
    return accu, precision, recall, f


class SentenceExtractor:
    """Accumulates what SLTester accumulates over the evaluation batches (Tester.py:75-152)."""

    def __init__(self, m: int, blocking_win: int = 3):
        self.m, self.blocking_win = m, blocking_win
        self.pred = self.true = self.match = self.match_true = 0
        self.total_sentence_num = self.example_num = self.batch_number = 0
        self.running_loss = 0.0
        self.extracts: List[List[int]] = []

    def evaluate(self, graph: HeteroBatch, logits: torch.Tensor, loss: Optional[float] = None,
                 sents: Optional[Sequence[Sequence[str]]] = None, blocking: bool = False) -> List[List[int]]:
        """logits [n sentences, 2] in batched sentence order (model.forward(graph)); sents: per graph (batch order) the
        original sentence strings, needed for blocking only.  Returns the selected local sentence indices per graph."""
        self.batch_number += 1
        if loss is not None:
            self.running_loss += float(loss)
        B = graph.n_graphs
        gptr = graph.graph_sent_ptr
        counts = (gptr[1:] - gptr[:-1]).tolist()
        n_sent = logits.shape[0]
        prediction = torch.zeros(n_sent, dtype=torch.int64, device=logits.device)
        if self.m == 0:
            prediction = (logits[:, 1] > logits[:, 0]).long()          # p_sent.max(1)[1]: ties go to class 0
            flat = prediction.tolist()
            offs = gptr.tolist()
            picked = [[i for i in range(counts[g]) if flat[offs[g] + i]] for g in range(B)]
        elif blocking:
            if sents is None:
                raise ValueError("n-gram blocking needs the original sentence strings of every graph")
            width = max(counts) if counts else 0
            order = _topm(logits, gptr, max(width, 1)).tolist()        # full descending order per graph, -1 padded
            picked = [ngram_blocking(sents[g], [i for i in order[g] if i >= 0], self.blocking_win, min(self.m, counts[g]))
                      for g in range(B)]
        else:
            top = _topm(logits, gptr, self.m)                          # [B, m] int32, -1 padded (Tester.py:128)
            picked = [[i for i in row if i >= 0] for row in top.tolist()]
        if self.m != 0:
            offs = gptr.tolist()
            sel = [offs[g] + i for g in range(B) for i in picked[g]]
            if sel:
                prediction[torch.tensor(sel, dtype=torch.int64, device=logits.device)] = 1
        label = graph.labels
        stats = torch.stack([prediction.sum(), label.sum(), ((prediction == label) & (prediction == 1)).sum(),
                             (prediction == label).sum()]).tolist()     # one D2H for the four counters
        self.pred += stats[0]
        self.true += stats[1]
        self.match_true += stats[2]
        self.match += stats[3]
        self.total_sentence_num += n_sent
        self.example_num += B
        self.extracts.extend(picked)
        return picked

    @property
    def running_avg_loss(self):
        return self.running_loss / self.batch_number

    def get_metric(self):
        """(accuracy, precision, recall, F) - SLTester.getMetric."""
        return eval_label(self.match_true, self.pred, self.true, self.total_sentence_num, self.match)
