"""Times the sentence-encoder stages on a synthetic batch (CUDA events, L2 flushed between iterations).

    python profiles/encoder_bench.py [n_graphs] [iters]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hetersumgraph_b200 import _lib  # noqa: E402
from hetersumgraph_b200 import synthetic as syn  # noqa: E402
from hetersumgraph_b200.encoder import EncoderPlan, SentenceEncoder  # noqa: E402


def timed(fn, iters, flush):
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    torch.manual_seed(0)
    exs = syn.make_examples(n_graphs, "cnndm", seed=0)
    tb = syn.pack_token_batch(exs)
    embed = torch.nn.Embedding(50000, 300, padding_idx=0)
    embed.weight.requires_grad_(False)
    enc = SentenceEncoder(embed, lstm_dropout=0.0).cuda()
    plan = EncoderPlan.from_token_batch(tb, "cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    S, L = tb.tokens.shape
    lib = _lib.load()

    def fwd_bwd():
        sf = enc(plan)
        sf.backward(cot)

    def ngram_only():
        ng = enc.ngram(plan)
        ng.backward(cot_ng)

    cot = torch.randn(S, 64, device="cuda")
    cot_ng = torch.randn(S, 300, device="cuda")
    for _ in range(3):
        fwd_bwd()
        ngram_only()
    torch.cuda.synchronize()
    res = {"graphs": n_graphs, "sentences": S, "compact_rows": plan.n_rows, "padded_rows": S * L,
           "ngram_fwd_ms": timed(lambda: enc.ngram(plan), iters, flush),
           "ngram_fwd_bwd_ms": timed(ngram_only, iters, flush),
           "encoder_fwd_bwd_ms": timed(fwd_bwd, iters, flush)}
    # per-kernel split of one encoder forward+backward (library event slots)
    lib.hsg_profile_reset()
    lib.hsg_profile_enable(1)
    for _ in range(5):
        flush.zero_()
        fwd_bwd()
    torch.cuda.synchronize()
    res["slots_ms_per_step"] = {k: round(v[1] / 5, 4) for k, v in _lib.profile_snapshot().items()}
    lib.hsg_profile_enable(0)
    ref_flops = S * sum((L + 1 - h) * h for h in range(2, 8)) * 50 * 300 * 2
    done_flops = plan.n_rows * 300 * 2100 * 2
    res["conv_fwd_reference_gflop"] = ref_flops / 1e9
    res["conv_fwd_executed_gflop"] = done_flops / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
