"""GEMM time vs K at fixed M, N: the slope is the steady-state cost per 32-deep k-block, the intercept the per-tile
epilogue / fill cost.  python profiles/gemm_kscale.py"""
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
    M = int(os.environ.get("KS_M", "151552"))          # 1184 row tiles = 8 per SM
    for N in (128, 512):
        for kind in ("nt", "nn"):
            for mode in ("tf32", "tf32x3"):
                for pair in (0, 1):
                    res = {}
                    for K in (64, 320, 640, 1280, 2560):
                        A = torch.randn(M, K, device=dev)
                        B = torch.randn(N, K, device=dev) if kind == "nt" else torch.randn(K, N, device=dev)
                        out = torch.empty(M, N, device=dev)
                        hb.set_gemm_mode(mode)
                        _lib.check(lib.hsg_set_gemm_pair(pair))
                        fn = (lambda: gemm_nt(A, B, out=out)) if kind == "nt" else (lambda: gemm_nn(A, B))
                        for _ in range(2):
                            fn()
                        torch.cuda.synchronize()
                        tot = 0.0
                        for _ in range(5):
                            flush.zero_()
                            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            a.record()
                            fn()
                            b.record()
                            torch.cuda.synchronize()
                            tot += a.elapsed_time(b)
                        res[K] = tot / 5 * 1e3
                        del A, B, out
                    tiles_per_cta = (M / 128) * (N / 128) / 148.0
                    kb = lambda K: K / 32.0      # noqa: E731
                    slope_us = (res[2560] - res[640]) / (kb(2560) - kb(640)) / tiles_per_cta
                    icpt_us = (res[640] - slope_us * kb(640) * tiles_per_cta) / tiles_per_cta
                    print(json.dumps({"N": N, "kind": kind, "mode": mode, "pair": pair,
                                      "us": {k: round(v, 1) for k, v in res.items()},
                                      "cycles_per_kblock_per_128x128": round(slope_us * 1965, 0),
                                      "cycles_per_tile_fixed": round(icpt_us * 1965, 0)}), flush=True)
    hb.set_gemm_mode("tf32x3")
    _lib.check(lib.hsg_set_gemm_pair(1))


if __name__ == "__main__":
    main()
