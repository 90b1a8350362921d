"""Pair (cta_group::2) against single-CTA tcgen05 kernel per product shape of the 32-graph step, back to back (warm L2).
python profiles/gemm_shape_bench.py"""
import json
import sys

import torch

sys.path.insert(0, ".")
from hetersumgraph_b200 import _lib  # noqa: E402
from hetersumgraph_b200.functional import gemm_nn, gemm_nt  # noqa: E402


def timed(fn, it=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3


def main():
    lib = _lib.load()
    _lib.require_device()
    out = []
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 11817
    for kind, N, K in (("nt", 72, 300), ("nn", 300, 72), ("nt", 512, 300), ("nt", 300, 512), ("nn", 512, 300), ("nn", 300, 512)):
        A = torch.randn(M, K, device="cuda")
        if kind == "nt":
            B = torch.randn(N, K, device="cuda")
            C = torch.empty(M, N, device="cuda")
            fn = lambda: gemm_nt(A, B, out=C)  # noqa: E731
        else:
            B = torch.randn(K, N, device="cuda")
            fn = lambda: gemm_nn(A, B)  # noqa: E731
        row = {"kind": kind, "M": M, "N": N, "K": K}
        for pair in (1, 0):
            lib.hsg_set_gemm_pair(pair)
            row["pair_us" if pair else "single_us"] = round(timed(fn), 2)
        lib.hsg_set_gemm_pair(1)
        out.append(row)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
