"""Where the step of a data-parallel shard goes: per-slot CUDA-event times (hsg_profile_*) of the update loop fwd+bwd +
loss on n CNN/DM-shaped graphs (default 2 048), next to the wall time of the step between two events.
python profiles/shard_profile.py [n_graphs] > gpurun_out/shard_profile.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from hetersumgraph_b200 import _lib, synthetic as syn
    from hetersumgraph_b200.graph import DeviceTokenBatch, HeteroBatch
    from hetersumgraph_b200.path_model import HSGPath, graph_loss
    n_graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _lib.require_device()
    lib = _lib.load()
    exs = syn.make_examples(n_graphs, "cnndm", seed=3)
    tb = syn.pack_token_batch(exs)
    dtb = DeviceTokenBatch.upload(tb, dev)
    torch.manual_seed(1234)
    model = HSGPath(n_iter=1).to(dev)
    sf = torch.randn(tb.tokens.shape[0], 64, device=dev)

    def step():
        b = HeteroBatch.build(dtb)
        loss = graph_loss(b, model(b, sf.detach().requires_grad_(True)), b.labels)
        for p in model.parameters():
            p.grad = None
        loss.backward()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    iters = 5
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    lib.hsg_profile_reset()
    lib.hsg_profile_enable(1)
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    lib.hsg_profile_enable(0)
    slots = {k: {"launches_per_step": v[0] / iters, "ms_per_step": round(v[1] / iters, 4)}
             for k, v in _lib.profile_snapshot().items()}
    tot = sum(v["ms_per_step"] for v in slots.values())
    print(json.dumps({"n_graphs": n_graphs, "step_ms": ms, "sum_of_slots_ms": tot,
                      "slots": dict(sorted(slots.items(), key=lambda kv: -kv[1]["ms_per_step"]))}, indent=1))


if __name__ == "__main__":
    main()
