"""Fused small-node-set FFN (hsg_ffn_rows_fwd / _bwd) against the three-kernel path (hsg_gemm_nt x2 + hsg_layernorm_fwd /
hsg_layernorm_bwd + hsg_gemm_nn x2): time per direction, back to back, and the max difference of the results.
python profiles/ffn_rows_bench.py [n F d_hid]"""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
from hetersumgraph_b200 import _lib  # noqa: E402


def main():
    n, F, Dh = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1009, 64, 512)
    lib = _lib.load()
    _lib.require_device()
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(n, F, device=dev)
    w1, b1 = torch.randn(Dh, F, device=dev) * 0.1, torch.randn(Dh, device=dev) * 0.1
    w2, b2 = torch.randn(F, Dh, device=dev) * 0.1, torch.randn(F, device=dev) * 0.1
    gamma, beta = torch.rand(F, device=dev) + 0.5, torch.randn(F, device=dev) * 0.1
    dy = torch.randn(n, F, device=dev)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    ws = torch.empty(lib.hsg_layernorm_bwd_workspace_bytes(n, F) + 1024, dtype=torch.uint8, device=dev)

    def bufs():
        return dict(hdn=torch.empty(n, Dh, device=dev), r=torch.empty(n, F, device=dev), y=torch.empty(n, F, device=dev),
                    st=torch.empty(n, 2, device=dev), dr=torch.empty(n, F, device=dev), dhp=torch.empty(n, Dh, device=dev),
                    dx=torch.empty(n, F, device=dev), dg=torch.zeros(F, device=dev), db=torch.zeros(F, device=dev))
    a, b = bufs(), bufs()

    def fused_fwd():
        _lib.check(lib.hsg_ffn_rows_fwd(n, F, Dh, p(x), p(w1), p(b1), p(w2), p(b2), p(gamma), p(beta), p(a["hdn"]), p(a["r"]),
                                        p(a["y"]), p(a["st"]), s))

    def fused_bwd():
        _lib.check(lib.hsg_ffn_rows_bwd(n, F, Dh, p(dy), p(a["r"]), p(a["st"]), p(gamma), p(a["hdn"]), p(w1), p(w2), p(a["dr"]),
                                        p(a["dhp"]), p(a["dx"]), p(a["dg"]), p(a["db"]), 0, p(ws), ws.numel(), s))

    def split_fwd():
        _lib.check(lib.hsg_gemm_nt(n, Dh, F, p(x), F, p(w1), F, p(b["hdn"]), Dh, p(b1), None, 0, 3, s))
        _lib.check(lib.hsg_gemm_nt(n, F, Dh, p(b["hdn"]), Dh, p(w2), Dh, p(b["r"]), F, p(b2), p(x), F, 5, s))
        _lib.check(lib.hsg_layernorm_fwd(n, F, p(b["r"]), p(gamma), p(beta), p(b["y"]), p(b["st"]), s))

    def split_bwd():
        _lib.check(lib.hsg_layernorm_bwd(n, F, p(dy), p(b["r"]), p(b["st"]), p(gamma), p(b["dr"]), p(b["dg"]), p(b["db"]), p(ws),
                                         ws.numel(), s))
        _lib.check(lib.hsg_gemm_nn(n, Dh, F, p(b["dr"]), F, p(w2), Dh, p(b["dhp"]), Dh, p(b["hdn"]), Dh, 8, s))
        _lib.check(lib.hsg_gemm_nn(n, F, Dh, p(b["dhp"]), Dh, p(w1), F, p(b["dx"]), F, p(b["dr"]), F, 4, s))

    def timed(fn, it=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it * 1e3

    out = {"n": n, "F": F, "d_hid": Dh, "ok": int(lib.hsg_ffn_rows_ok(n, F, Dh))}
    out["fused_fwd_us"], out["split_fwd_us"] = timed(fused_fwd), timed(split_fwd)
    out["fused_bwd_us"], out["split_bwd_us"] = timed(fused_bwd), timed(split_bwd)
    err = lambda u, v: float((u - v).abs().max() / (v.abs().max() + 1e-30))  # noqa: E731
    out["max_rel_diff"] = {k: err(a[k], b[k]) for k in ("hdn", "r", "y", "st", "dr", "dhp", "dx", "dg", "db")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
