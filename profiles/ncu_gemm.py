"""One FFN-1 launch (M=11817, N=512, K=300, bias+ReLU epilogue, 3xTF32) with a CLEAN, cold L2 for an ncu --set full
capture: the operands are written, then L2 is evicted by READING a 512 MB buffer (clean lines only, so nothing of an
earlier kernel is written back during the measured launch - round 1's traffic figure contained 55 MB of such
write-backs).  Run as
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -o OUT \
      python profiles/ncu_gemm.py [mode]"""
import sys

import torch

sys.path.insert(0, ".")
import hetersumgraph_b200 as hb  # noqa: E402
from hetersumgraph_b200.functional import gemm_nt  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
    hb.set_gemm_mode(mode)
    M, N, K = 11817, 512, 300
    torch.manual_seed(0)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    big = torch.ones(128 << 20, dtype=torch.float32, device="cuda")      # 512 MB
    for _ in range(2):
        gemm_nt(A, B, bias=bias, epi=3, out=out)
    torch.cuda.synchronize()
    float(big.sum())                                                      # read-only sweep: L2 now holds clean lines of `big`
    torch.cuda.synchronize()
    gemm_nt(A, B, bias=bias, epi=3, out=out)                              # <- the captured launch (third of this kernel)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
