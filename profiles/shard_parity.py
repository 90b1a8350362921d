"""Shard-vs-whole gradient parity emulated on ONE GPU (sum of the shard arenas against the whole batch), with the per-parameter
attribution of the differences: python profiles/shard_parity.py N_GLOBAL WORLD [recompute mode] [gemm mode] [small-product flops].
Used to find that a shard and the whole batch fall on different sides of the small-product threshold (dist.gradient_parity)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from hetersumgraph_b200 import _lib, synthetic as syn
from hetersumgraph_b200.dist import FlatGradArena, shard_indices
from hetersumgraph_b200.graph import HeteroBatch
from hetersumgraph_b200.path_model import FusedTrainStep, HSGPath
lib = _lib.load()
dev = torch.device("cuda", 0)
n_global = int(sys.argv[1]); world = int(sys.argv[2]); mode = int(sys.argv[3]) if len(sys.argv) > 3 else -1
gm = sys.argv[4] if len(sys.argv) > 4 else None
if gm:
    import hetersumgraph_b200 as hb
    hb.set_gemm_mode(gm)
lib.hsg_set_edge_recompute(mode)
if len(sys.argv) > 5: lib.hsg_set_gemm_small_flops(float(sys.argv[5]))
exs_all = syn.make_examples(n_global, "cnndm", seed=3)
sf_all = torch.randn(sum(e.n_sent for e in exs_all) + 8, 64, generator=torch.Generator().manual_seed(11))
offs = np.concatenate([[0], np.cumsum([e.n_sent for e in exs_all])])
names = None
def run(exs_sub, idxs):
    global names
    tbs = syn.pack_token_batch(exs_sub)
    order = list(tbs.order)
    rows = np.concatenate([np.arange(offs[idxs[j]], offs[idxs[j]] + exs_sub[j].n_sent) for j in order])
    batch = HeteroBatch.from_token_batch(tbs, dev)
    torch.manual_seed(1234)
    m = HSGPath(n_iter=1).to(dev)
    ar = FlatGradArena(m.parameters(), flatten_params=True)
    m.loop.fuse_grad_accumulation = True
    FusedTrainStep(m, n_global)(batch, sf_all[rows].to(dev))
    names = [(k, p.numel()) for k, p in m.named_parameters() if p.requires_grad]
    return ar.flat.clone()
sh = shard_indices([e.n_sent for e in exs_all], [float(sum(len(x) for x in e.w2s)) for e in exs_all], world)
g = None
for r in range(world):
    gr = run([exs_all[i] for i in sh[r]], sh[r])
    g = gr if g is None else g + gr
full = run(exs_all, list(range(n_global)))
print("n_global", n_global, "world", world, "mode", mode, "err", float((g - full).abs().max() / full.abs().max()))
o = 0
fm = float(full.abs().max())
for k, n in names:
    seg_a, seg_b = g[o:o + n], full[o:o + n]
    e = float((seg_a - seg_b).abs().max())
    if e / fm > 1e-6:
        d = (seg_a - seg_b).abs()
        print("  %-40s err/arena_max %.2e  own max %.3e  n>1e-6: %d of %d  argmax %d" % (k, e / fm, float(seg_b.abs().max()), int((d / fm > 1e-6).sum()), n, int(d.argmax())))
    o += n
