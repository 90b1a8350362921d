"""Pipeline trace of CTA 0 of the tcgen05 GEMM (hsg_gemm_trace) on the dominant shape, both arithmetic modes.

Events (hsg_gemm_tc.cu): 1 TMA issue of k-block it, 2 MMA thread saw the stage ready, 3 MMAs of the k-block issued +
committed, 4 converters saw the stage land, 5 converters done, 6 epilogue start of tile tl, 7 epilogue end of tile tl.
Prints per-k-block and per-tile cycle deltas so the bound (TMA latency / conversion / MMA / epilogue) can be read off.
Usage (GPU box): python profiles/gemm_trace.py [M N K]
"""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from hetersumgraph_b200 import _lib  # noqa: E402


def run(M, N, K, mode, epi_bias=True):
    lib = _lib.load()
    _lib.require_device()
    _lib.set_gemm_mode(mode)
    lib.hsg_set_gemm_small_flops(C.c_double(0.0))
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    Cm = torch.empty(M, N, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def call():
        _lib.check(lib.hsg_gemm_nt(M, N, K, A.data_ptr(), K, B.data_ptr(), K, Cm.data_ptr(), N, bias.data_ptr(), None, 0,
                                   (_lib.EPI_BIAS | _lib.EPI_RELU) if epi_bias else 0, C.c_void_p(s)))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    # timing, L2 flushed, one launch at a time
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    # warm, back to back
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        call()
    e1.record()
    torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) * 1e3 / 20
    lib.hsg_gemm_trace(1, None, 0)
    flush.zero_()
    torch.cuda.synchronize()
    call()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (3 * 2048))()
    n = lib.hsg_gemm_trace(-1, buf, 2048)
    lib.hsg_gemm_trace(0, None, 0)
    ev = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 3)[:n].astype(np.int64)
    tr = {}
    for e, it, clk in ev:
        if clk:
            tr[(int(e), int(it))] = int(clk)
    t0 = min(tr.values())
    out = {"M": M, "N": N, "K": K, "mode": mode, "cold_us_median": float(np.median(ts)), "cold_us_min": float(min(ts)),
           "back_to_back_us": warm}
    nkb = (K + 31) // 32
    its = sorted(it for (e, it) in tr if e == 1)
    rows = []
    for it in its:
        r = {"it": it}
        for e, name in ((1, "tma"), (4, "land"), (5, "conv"), (2, "mma_go"), (3, "mma_iss")):
            if (e, it) in tr:
                r[name] = tr[(e, it)] - t0
        rows.append(r)
    out["kblocks"] = rows
    tiles = []
    for tl in range(0, 8):
        if (6, tl) in tr:
            tiles.append({"tile": tl, "epi_start": tr[(6, tl)] - t0, "epi_end": tr.get((7, tl), 0) - t0})
    out["tiles"] = tiles
    out["nkb_per_tile"] = nkb
    return out


if __name__ == "__main__":
    shape = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else [11817, 512, 300]
    for mode in ("tf32x3", "tf32"):
        o = run(*shape, mode)
        print(json.dumps({k: v for k, v in o.items() if k not in ("kblocks",)}))
        prev = None
        for r in o["kblocks"]:
            d = (r.get("tma", 0) - prev) if prev is not None else 0
            prev = r.get("tma", 0)
            print("  it %3d tma %7d (+%5d) land +%5d conv +%5d mma_go +%5d mma_iss +%5d" % (
                r["it"], r.get("tma", -1), d, r.get("land", 0) - r.get("tma", 0), r.get("conv", 0) - r.get("land", 0),
                r.get("mma_go", 0) - r.get("tma", 0), r.get("mma_iss", 0) - r.get("mma_go", 0)))
