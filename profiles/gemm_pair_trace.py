"""Pipeline trace of the first cluster of the CTA-pair GEMM.  Events: 1 TMA issue, 4 converters saw the stage land,
5 converters done (before the arrive), 6 after the arrive, 2 MMA thread saw `ready`, 3 MMAs issued + committed
(+10: the peer CTA's events).  python profiles/gemm_pair_trace.py [M N K mode]"""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from hetersumgraph_b200 import _lib  # noqa: E402


def main():
    M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (151552, 512, 1280)
    mode = sys.argv[4] if len(sys.argv) > 4 else "tf32"
    lib = _lib.load()
    _lib.require_device()
    _lib.set_gemm_mode(mode)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    Cm = torch.empty(M, N, device="cuda")
    s = torch.cuda.current_stream().cuda_stream

    def call():
        _lib.check(lib.hsg_gemm_nt(M, N, K, A.data_ptr(), K, B.data_ptr(), K, Cm.data_ptr(), N, None, None, 0, 0, C.c_void_p(s)))
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    lib.hsg_gemm_pair_trace(1, None, 0)
    call()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (3 * 4096))()
    n = lib.hsg_gemm_pair_trace(-1, buf, 4096)
    lib.hsg_gemm_pair_trace(0, None, 0)
    ev = np.frombuffer(buf, dtype=np.uint64).reshape(-1, 3)[:n].astype(np.int64)
    tr = {(int(e), int(it)): int(c) for e, it, c in ev if c}
    t0 = min(tr.values())
    names = {1: "tma0", 11: "tma1", 4: "land0", 14: "land1", 5: "cdone0", 15: "cdone1", 6: "arr0", 16: "arr1", 2: "mma_go",
             3: "mma_commit", 7: "epi_go0", 8: "epi_end0", 17: "epi_go1", 18: "epi_end1"}
    rows = []
    for it in range(0, 48):
        r = {"it": it}
        for e, nm in names.items():
            if (e, it) in tr:
                r[nm] = tr[(e, it)] - t0
        rows.append(r)
    for r in rows:
        print(json.dumps(r))
    # per-CTA spans: are all CTA pairs co-resident?
    sp = (C.c_ulonglong * (3 * 160))()
    lib.hsg_gemm_pair_trace(-2, sp, 3 * 160)
    a = np.frombuffer(sp, dtype=np.uint64).reshape(-1, 3).astype(np.int64)
    a = a[a[:, 0] > 0]
    if len(a):
        t0 = a[:, 0].min()
        print(json.dumps({"max_active_clusters": lib.hsg_gemm_pair_trace(-3, None, 0), "ctas": len(a),
                          "start_us": [round((x - t0) / 1e3, 1) for x in a[:, 0].tolist()],
                          "end_us": [round((x - t0) / 1e3, 1) for x in a[:, 1].tolist()],
                          "smid": a[:, 2].tolist()}))


if __name__ == "__main__":
    main()
