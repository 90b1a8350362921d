"""us per launch of the tensor-core products at the shapes of the 32-graph step (and a 2 048-graph shard), L2 flushed
before every launch, CUDA events, single-CTA vs CTA-pair kernel.  python profiles/gemm_bench.py [M ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import hetersumgraph_b200 as hb
    from hetersumgraph_b200 import _lib
    from hetersumgraph_b200.functional import gemm_nn, gemm_nt
    lib = _lib.load()
    _lib.require_device()
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    Ms = [int(x) for x in sys.argv[1:]] or [11817, 731003]
    peak = 1375.9
    for M in Ms:
        shapes = [("nt", M, 512, 300, 3), ("nt", M, 300, 512, 5), ("nn", M, 512, 300, 8), ("nn", M, 300, 512, 4),
                  ("nt", M, 72, 300, 0), ("nn", M, 300, 72, 0)]
        for kind, m, n, k, epi in shapes:
            A = torch.randn(m, k, device=dev)
            B = torch.randn(n, k, device=dev) * 0.05 if kind == "nt" else torch.randn(k, n, device=dev) * 0.05
            bias = torch.randn(n, device=dev)
            R = torch.randn(m, n, device=dev)
            row = {"kind": kind, "M": m, "N": n, "K": k}
            for mode in ("tf32x3", "tf32"):
                hb.set_gemm_mode(mode)
                for pair in (0, 1):
                    _lib.check(lib.hsg_set_gemm_pair(pair))

                    def fn():
                        if kind == "nt":
                            gemm_nt(A, B, bias=bias if epi & 1 else None, R=R if epi & 4 else None, epi=epi)
                        else:
                            gemm_nn(A, B, R=R if epi & 12 else None, epi=epi)
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    its = 10 if m > 100000 else 20
                    tot = 0.0
                    for _ in range(its):
                        flush.zero_()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        fn()
                        b.record()
                        torch.cuda.synchronize()
                        tot += a.elapsed_time(b)
                    us = tot / its * 1e3
                    tf = 2.0 * m * n * k / (us * 1e-6) / 1e12
                    row["%s_%s_us" % (mode, "pair" if pair else "single")] = round(us, 2)
                    row["%s_%s_frac" % (mode, "pair" if pair else "single")] = round(tf / peak, 4)
            hb.set_gemm_mode("tf32x3")
            _lib.check(lib.hsg_set_gemm_pair(1))
            print(json.dumps(row), flush=True)
            del A, B, R


if __name__ == "__main__":
    main()
