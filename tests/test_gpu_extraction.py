"""GPU test (-m gpu) of extraction.SentenceExtractor against a literal restatement of SLTester.evaluation's selection and
counters (Tester.py:96-137): argmax mode (m == 0), top-m (bit-exact indices, hsg_topm), n-gram blocking."""
import numpy as np
import pytest
import torch

import hetersumgraph_b200 as hb
from hetersumgraph_b200 import synthetic as syn
from hetersumgraph_b200.extraction import SentenceExtractor, eval_label, ngram_blocking

pytestmark = pytest.mark.gpu


def restated_evaluation(logits, labels, ptr, m, sents=None, n_win=3):
    """per graph, as the reference does after dgl.unbatch (Tester.py:105-137)"""
    pred = true = match = match_true = 0
    extracts = []
    for g in range(len(ptr) - 1):
        p_sent = logits[ptr[g]:ptr[g + 1]]
        label = labels[ptr[g]:ptr[g + 1]]
        N = p_sent.shape[0]
        if m == 0:
            prediction = p_sent.max(1)[1]
            pred_idx = torch.arange(N)[prediction != 0].long()
        else:
            if sents is not None:
                order = p_sent[:, 1].sort(descending=True)[1].tolist()
                pred_idx = torch.LongTensor(ngram_blocking(sents[g], order, n_win, min(m, N)))
            else:
                pred_idx = torch.topk(p_sent[:, 1], min(m, N))[1]
            prediction = torch.zeros(N).long()
            prediction[pred_idx] = 1
        extracts.append(pred_idx.tolist())
        pred += int(prediction.sum())
        true += int(label.sum())
        match_true += int(((prediction == label) & (prediction == 1)).sum())
        match += int((prediction == label).sum())
    return extracts, (match_true, pred, true, logits.shape[0], match)


@pytest.mark.parametrize("m", [0, 3, 60])
@pytest.mark.parametrize("blocking", [False, True])
def test_sentence_extractor_matches_restated_evaluation(m, blocking):
    if m == 0 and blocking:
        pytest.skip("blocking only applies to m > 0")
    exs = syn.make_examples(6, "cnndm", seed=8)
    tb = syn.pack_token_batch(exs)
    batch = hb.HeteroBatch.from_token_batch(tb, "cuda")
    g = torch.Generator().manual_seed(m + 10 * blocking)
    logits = torch.randn(tb.tokens.shape[0], 2, generator=g)
    ptr = tb.graph_sent_ptr.tolist()
    rng = np.random.default_rng(1)
    words = ["w%d" % i for i in range(40)]
    sents = [[" ".join(rng.choice(words, size=int(rng.integers(0, 12)))) for _ in range(ptr[i + 1] - ptr[i])]
             for i in range(len(ptr) - 1)] if blocking else None
    ref_extracts, ref_counts = restated_evaluation(logits, torch.from_numpy(tb.labels), ptr, m, sents)
    ex = SentenceExtractor(m)
    got = ex.evaluate(batch, logits.cuda(), loss=1.5, sents=sents, blocking=blocking)
    assert got == ref_extracts                      # bit-exact indices, same order
    assert (ex.match_true, ex.pred, ex.true, ex.total_sentence_num, ex.match) == ref_counts
    assert ex.get_metric() == eval_label(*ref_counts)
    assert ex.example_num == tb.n_graphs and ex.running_avg_loss == 1.5
