"""Kernel timeline of the replayed training step (CUPTI through torch.profiler): start / duration / stream of every
kernel of a few replays of bench.py's `value` leg.  Run on the GPU box:  python profiles/timeline.py > gpurun_out/tl.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from hetersumgraph_b200 import _lib, synthetic as syn
    from hetersumgraph_b200.dist import FlatGradArena
    from hetersumgraph_b200.functional import FusedAdam
    from hetersumgraph_b200.graph import DeviceTokenBatch
    from hetersumgraph_b200.path_model import HSGPath
    from hetersumgraph_b200.step_graph import GraphedTrainStep
    n_graphs = int(os.environ.get("TL_GRAPHS", "32"))
    capture = os.environ.get("TL_EAGER", "0") != "1"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.require_device()
    exs = syn.make_examples(n_graphs, "cnndm", seed=0)
    tb = syn.pack_token_batch(exs)
    torch.manual_seed(1234)
    model = HSGPath(n_iter=1).to(dev)
    arena = FlatGradArena(model.parameters(), flatten_params=True)
    model.loop.fuse_grad_accumulation = True
    opt = FusedAdam(arena.flat_param.data, arena.flat, lr=5e-4)
    host, _ = DeviceTokenBatch.host_buffers(tb)
    bitmap = torch.from_numpy(tb.filter_bitmap.view(np.int32).copy()).to(dev)
    gs = GraphedTrainStep(model, opt, bitmap, n_graphs, None, capture=capture, resident_tokens=True)
    gs.prime(host)
    sf = torch.randn(int(tb.tokens.shape[0]), 64, device=dev)
    gs._stage_sf(sf)
    sfr = gs.sf_dev[:sf.shape[0]]
    for _ in range(12):
        gs.step(host, sfr)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(4):
            gs.step(host, sfr)
        torch.cuda.synchronize()
    evs = []
    for e in prof.profiler.kineto_results.events():
        if str(e.device_type()).endswith("CUDA"):
            evs.append((e.start_ns() / 1e3, e.duration_ns() / 1e3, e.name()[:70], int(e.device_resource_id())))
    evs.sort()
    t0 = evs[0][0] if evs else 0
    out = [{"t_us": round(s - t0, 2), "dur_us": round(d, 2), "name": n, "stream": st} for s, d, n, st in evs]
    print(json.dumps({"graphs": n_graphs, "capture": capture, "events": out}))


if __name__ == "__main__":
    main()
