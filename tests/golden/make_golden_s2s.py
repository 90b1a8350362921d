"""Golden vectors for the S2S layer type (module/GAT.py:38-39,50-52; SGATLayer module/GATLayer.py:49-78;
MultiHeadSGATLayer module/GATStackLayer.py:27-44) - the reference's own classes run verbatim on the DGL-0.4 shim.
The reference never instantiates this layer type (HiGraph.py:57-76); it is covered for completeness.

    python tests/golden/make_golden_s2s.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the shim, imports the reference)

from hetersumgraph_b200 import synthetic as syn  # noqa: E402
from oracle import fixtures as fx  # noqa: E402
from oracle import graph_builder_ref as gb  # noqa: E402


def s2s_golden(name, exs, hdsg, hid, nh, ffn_h, seed):
    filt = set(syn.filter_ids().tolist())
    graphs = [(mg.ref_graph_hdsg if hdsg else mg.ref_graph_hsg)(e, filt) for e in exs]
    order = gb.stable_desc_order([e.n_sent for e in exs]).tolist()
    BG = mg.shim.batch([graphs[i] for i in order])
    ga = mg.shim_to_arrays(BG)
    torch.manual_seed(seed)
    layer = mg.WSWGAT(hid, hid, nh, 0.1, ffn_h, 0.1, 50, "S2S").eval()
    with torch.no_grad():
        layer.ffn.layer_norm.weight.uniform_(0.5, 1.5)
        layer.ffn.layer_norm.bias.uniform_(-0.2, 0.2)
    ns = int((ga.unit == 1).sum())
    s = torch.randn(ns, hid, requires_grad=True)
    cs = torch.randn(ns, hid)
    out_s = layer(BG, s, s)
    (out_s * cs).sum().backward()
    out = {}
    out.update(fx.examples_to_arrays(exs, "ex"))
    out.update(fx.graph_to_arrays(ga, "g_"))
    out["order"] = np.asarray(order, np.int64)
    out["dims"] = np.asarray([hid, nh, ffn_h, int(hdsg)], np.int64)
    out["in_s"], out["cs"], out["out_s"], out["grad_in_s"] = s.detach().numpy(), cs.numpy(), out_s.detach().numpy(), s.grad.numpy()
    for k, v in layer.state_dict(keep_vars=True).items():
        out["p:" + k] = v.detach().numpy()
        out["gp:" + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "Ns", ns, "out abs max", float(out_s.abs().max()), [k for k in out if k.startswith("p:")][:4])


if __name__ == "__main__":
    exs = syn.make_examples(3, "tiny", seed=61)
    exs.append(mg.edge_case_examples()[0])
    s2s_golden("s2s_hsg.npz", exs, False, 64, 8, 512, 4321)
    exd = syn.make_examples(3, "tiny", seed=62, hdsg=True)
    exd.append(mg.edge_case_examples()[1])
    s2s_golden("s2s_hdsg.npz", exd, True, 64, 8, 512, 4322)
