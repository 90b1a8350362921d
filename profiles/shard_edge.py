"""Edge kernels of a data-parallel shard (default 2 048 CNN/DM-shaped graphs) alone, for ncu:
ncu --set full -k regex:edge -c 12 python profiles/shard_edge.py [n_graphs] [W2S|S2W]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from hetersumgraph_b200 import _lib, synthetic as syn
    from hetersumgraph_b200.graph import DeviceTokenBatch, HeteroBatch
    n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.require_device()
    exs = syn.make_examples(n_graphs, "cnndm", seed=3)
    dtb = DeviceTokenBatch.upload(syn.pack_token_batch(exs), dev)
    batch = HeteroBatch.build(dtb)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pk = bench.peaks() if hasattr(bench, "peaks") else None
    rows = bench._time_edge_kernels(batch, "%d cnndm graphs" % n_graphs, dev, pk, flush, int(os.environ.get("ITERS", "3")))
    print(json.dumps(rows, indent=0))


if __name__ == "__main__":
    main()
