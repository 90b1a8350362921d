"""Edge kernels alone on a data-parallel shard (default 2 048 CNN/DM-shaped graphs), L2 flushed, CUDA events.

    python profiles/edge_bench.py [n_graphs] [iters]          (GPU box)
Prints one JSON line per (kernel, layer, row mapping) with the fraction of the measured HBM peak (algorithmic bytes,
hetersumgraph_b200/accounting.py).  The row-mapping knobs of the library are swept so one call compares the variants.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hetersumgraph_b200 import _lib  # noqa: E402


def main():
    n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    from hetersumgraph_b200 import synthetic as syn
    from hetersumgraph_b200.graph import DeviceTokenBatch, HeteroBatch
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    pk = bench.peaks()
    lib = _lib.load()
    _lib.require_device()
    exs = syn.make_examples(n_graphs, "cnndm", seed=3)
    tb = syn.pack_token_batch(exs)
    batch = HeteroBatch.build(DeviceTokenBatch.upload(tb, dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    variants = [("auto", -1, -1, -1, -1), ("shared-row fwd, shared-row bwd", 0, 0, 0, 0),
                ("row-parallel fwd+bwd", 1, 1, 0, 0), ("async-gather bwd", 0, 0, 0, 1)]
    only = os.environ.get("EDGE_BENCH_VARIANTS")
    if only:
        variants = [v for v in variants if v[0].split()[0] in only.split(",")]
    for name, fr, br, blk, asy in variants:
        lib.hsg_set_edge_bwd_async(asy)
        lib.hsg_set_edge_fwd_rowpar(fr)
        lib.hsg_set_edge_rowpar(br)
        lib.hsg_set_edge_blockrow(blk)
        for r in bench._time_edge_kernels(batch, "%d cnndm graphs" % n_graphs, dev, pk, flush, iters):
            print(json.dumps({"mapping": name, "kernel": r["kernel"], "layer": r["layer"], "us": round(r["ms"] * 1e3, 1),
                              "MB": round(r["algorithmic_MB"], 1), "frac": round(r["frac_of_hbm_peak"], 3)}))
    lib.hsg_set_edge_bwd_async(-1)
    lib.hsg_set_edge_fwd_rowpar(-1)
    lib.hsg_set_edge_rowpar(-1)
    lib.hsg_set_edge_blockrow(-1)


if __name__ == "__main__":
    main()
