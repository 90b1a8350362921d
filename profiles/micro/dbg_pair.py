import sys, torch
sys.path.insert(0, ".")
import hetersumgraph_b200 as hb
from hetersumgraph_b200 import _lib
from hetersumgraph_b200.functional import gemm_nt, gemm_nn
lib = _lib.load()
hb.set_gemm_mode("tf32x3")
for trial in range(3):
    for (M, N, K) in [(1000, 72, 300), (4099, 300, 512), (2000, 512, 300)]:
        torch.manual_seed(0)
        A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
        bias = torch.randn(N, device="cuda"); R = torch.randn(M, N, device="cuda")
        ref = (A.double() @ B.double().t())
        out = gemm_nt(A, B)
        d = (out.double() - ref).abs()
        bad = d > 1e-3
        rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
        print(trial, M, N, K, "bad", int(bad.sum()), "rows", rows[:4].tolist(), rows[-4:].tolist() if len(rows) else [], "cols", cols[:6].tolist(), cols[-6:].tolist() if len(cols) else [], len(rows), len(cols))
        out2 = gemm_nt(A, B, bias=bias, epi=3)
        d = (out2.double() - torch.relu(ref + bias.double())).abs(); print("   bias/relu bad", int((d > 1e-3).sum()))
        out3 = gemm_nt(A, B, bias=bias, R=R, epi=5)
        d = (out3.double() - (ref + bias.double() + R.double())).abs(); print("   bias/add bad", int((d > 1e-3).sum()))
