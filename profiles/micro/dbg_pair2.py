import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import hetersumgraph_b200 as hb
from hetersumgraph_b200 import _lib
from hetersumgraph_b200.functional import gemm_nt, gemm_nn, gemm_tn
lib = _lib.load()
def nerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
shapes = [(1000, 72, 300), (257, 312, 64), (129, 512, 64), (1, 64, 512), (333, 50, 30), (4099, 300, 512), (2000, 512, 300), (77, 16, 48)]
for mode in ["fp32", "tf32x3", "tf32"]:
    hb.set_gemm_mode(mode)
    for (M, N, K) in shapes:
        torch.manual_seed(0)
        A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
        bias = torch.randn(N, device="cuda"); R = torch.randn(M, N, device="cuda")
        ref = (A.double() @ B.double().t())
        outs = [("plain", gemm_nt(A, B), ref), ("relu", gemm_nt(A, B, bias=bias, epi=3), torch.relu(ref + bias.double())),
                ("add", gemm_nt(A, B, bias=bias, R=R, epi=5), ref + bias.double() + R.double())]
        Bn = torch.randn(K, N, device="cuda")
        refn = A.double() @ Bn.double()
        outs += [("nn", gemm_nn(A, Bn), refn), ("nnmask", gemm_nn(A, Bn, R=R, epi=8), torch.where(R > 0, refn, torch.zeros_like(refn))),
                 ("nnadd", gemm_nn(A, Bn, R=R, epi=4), refn + R.double())]
        A2 = torch.randn(M, N, device="cuda")
        Ct, cs = gemm_tn(A2, A, want_colsum=True)
        Ct2, none = gemm_tn(A2, A, want_colsum=False)
        for name, o, r in outs:
            e = nerr(o, r)
            if e > (3e-2 if mode == "tf32" else 1e-5):
                bad = (o.double() - r).abs() > 1e-2 * float(r.abs().max())
                rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
                print("FAIL", mode, M, N, K, name, e, "nbad", int(bad.sum()), "rows", rows[:6].tolist(), rows[-3:].tolist(), "cols", cols[:6].tolist(), cols[-3:].tolist(), len(rows), len(cols))
print("done")
