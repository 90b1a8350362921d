import sys, os, torch
sys.path.insert(0, ".")
import hetersumgraph_b200 as hb
from hetersumgraph_b200 import _lib
from hetersumgraph_b200.functional import gemm_nt, gemm_nn, gemm_tn
lib = _lib.load()
hb.set_gemm_mode(os.environ.get("MODE", "tf32x3"))
M, N, K = 4099, 300, 512
inter = os.environ.get("INTER", "nn")
fails = 0
iters = int(os.environ.get("ITERS", "300"))
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda")
Bn = torch.randn(K, N, device="cuda")
R = torch.randn(M, N, device="cuda")
for it in range(iters):
    if inter == "nn":
        o = gemm_nn(A, Bn, R=R, epi=4)
        del o
    A2 = torch.randn(M, N, device="cuda")
    if os.environ.get("COLSUM", "1") == "1":
        Ct, cs = gemm_tn(A2, A, want_colsum=True)
    elif os.environ.get("COLSUM") == "2":
        Ct, cs = gemm_tn(A2, A, want_colsum=True)
        Ct, cs = gemm_tn(A2, A, want_colsum=False)
    else:
        Ct, cs = gemm_tn(A2, A, want_colsum=False)
    ref = A2.double().t() @ A.double()
    d = (Ct.double() - ref).abs()
    if float(d.max()) > 1e-2 * float(ref.abs().max()):
        bad = d > 1e-2 * float(ref.abs().max())
        rows, cols = bad.any(1).nonzero().flatten(), bad.any(0).nonzero().flatten()
        fails += 1
        if fails <= 5:
            print("it", it, "nbad", int(bad.sum()), "rows", rows[:2].tolist(), rows[-1:].tolist(), len(rows), "cols", cols[:2].tolist(), cols[-1:].tolist(), len(cols))
print("MODE", os.environ.get("MODE", "tf32x3"), "INTER", inter, "PDL", os.environ.get("HSG_PDL"), "PAIR", os.environ.get("HSG_GEMM_PAIR"), "fails", fails, "of", iters)
