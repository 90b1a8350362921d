"""ms per replayed training step (bench.py's `value` leg: 32 cnndm graphs, resident inputs, L2 flushed between steps)
for a sweep of library knobs.  python profiles/step_sweep.py side_ctas=0,128,104,80 tn_min_rows=256,768"""
import itertools
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SETTERS = {"side_ctas": "hsg_set_side_ctas", "tn_min_rows": "hsg_set_tn_min_rows", "pdl": "hsg_set_pdl",
           "overlap": "hsg_set_bwd_overlap", "gemm_mode": "hsg_set_gemm_mode", "tn_item_rows": "hsg_set_tn_item_rows",
           "small_flops": "hsg_set_gemm_small_flops"}


def main():
    from hetersumgraph_b200 import _lib, synthetic as syn
    from hetersumgraph_b200.dist import FlatGradArena
    from hetersumgraph_b200.functional import FusedAdam
    from hetersumgraph_b200.graph import DeviceTokenBatch
    from hetersumgraph_b200.path_model import HSGPath
    from hetersumgraph_b200.step_graph import GraphedTrainStep
    sweep = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        sweep[k] = [int(x) for x in v.split(",")]
    n_graphs = int(os.environ.get("SW_GRAPHS", "32"))
    steps = int(os.environ.get("SW_STEPS", "30"))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.require_device()
    lib = _lib.load()
    exs = syn.make_examples(n_graphs, "cnndm", seed=0)
    tb = syn.pack_token_batch(exs)
    torch.manual_seed(1234)
    model = HSGPath(n_iter=1).to(dev)
    arena = FlatGradArena(model.parameters(), flatten_params=True)
    model.loop.fuse_grad_accumulation = True
    opt = FusedAdam(arena.flat_param.data, arena.flat, lr=5e-4)
    host, _ = DeviceTokenBatch.host_buffers(tb)
    bitmap = torch.from_numpy(tb.filter_bitmap.view(np.int32).copy()).to(dev)
    gs = GraphedTrainStep(model, opt, bitmap, n_graphs, None, capture=True, resident_tokens=True)
    gs.prime(host)
    sf = torch.randn(int(tb.tokens.shape[0]), 64, device=dev)
    gs._stage_sf(sf)
    sfr = gs.sf_dev[:sf.shape[0]]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    keys = list(sweep)
    for combo in itertools.product(*[sweep[k] for k in keys]):
        for k, v in zip(keys, combo):
            _lib.check(getattr(lib, SETTERS[k])(v))
        gs._invalidate()
        for _ in range(8):
            gs.step(host, sfr)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.zero_()
            a.record()
            gs.step(host, sfr)
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        print(json.dumps({"cfg": dict(zip(keys, combo)), "ms_mean": sum(ms) / len(ms), "ms_median": ms[len(ms) // 2],
                          "ms_min": ms[0], "replays": gs.replays}), flush=True)


if __name__ == "__main__":
    main()
