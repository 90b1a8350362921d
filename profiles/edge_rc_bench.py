"""Edge kernels of the S2W layer on a data-parallel shard (n cnndm graphs): the path that saves sh (forward storing sh,
bwd-prep reading it) against the recomputing path (forward without sh, hsg_edge_bwd_prep_rc), each kernel timed alone with CUDA events, L2 flushed before every launch.  Fractions are of the measured HBM peak over
SURVEY.md 8(d)'s B_fwd / B_bwd.

    python profiles/edge_seg_bench.py [--graphs 2048] [--iters 10] [--ncu KERNEL]   (--ncu: run one kernel only)
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--shape", default="cnndm")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import hetersumgraph_b200 as hb
    from hetersumgraph_b200 import _lib, accounting
    from hetersumgraph_b200 import synthetic as syn
    from hetersumgraph_b200.functional import _Workspace
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    hbm = None
    for key in ("hbm_gbs", "hbm_gbps"):
        if key in peaks:
            hbm = peaks[key]
            break
    if hbm is None:
        hbm = 6545.9
    exs = syn.make_examples(a.graphs, a.shape, seed=3)
    batch = hb.HeteroBatch.from_token_batch(syn.pack_token_batch(exs))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    H, d = 6, 50
    csc, csc_t = batch.csc("S2W")
    F = H * d
    fp, ldz = _lib.edge_layout(H, d)
    E = csc.n_edges
    zp = torch.randn(csc.n_src, ldz, device=dev)
    q = torch.randn(10, H, device=dev)
    origin = torch.randn(csc.n_dst, F, device=dev)
    sh = torch.empty(csc.n_dst, F, device=dev)
    x = torch.empty(csc.n_dst, F, device=dev)
    stat = torch.empty(csc.n_dst, 3 * H, device=dev)
    g = torch.empty(csc.n_dst, fp, device=dev)
    dzp = torch.empty(csc.n_src, ldz, device=dev)
    dq = torch.empty(10, H, device=dev)
    ws = _Workspace.get(lib.hsg_edge_bwd_workspace_bytes(H), dev, "edge")
    fns = {
        "edge_fwd": lambda: _lib.check(lib.hsg_edge_fwd(C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(),
                                                        origin.data_ptr(), sh.data_ptr(), x.data_ptr(), stat.data_ptr(), st)),
        "edge_fwd_no_sh": lambda: _lib.check(lib.hsg_edge_fwd(C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(),
                                                              origin.data_ptr(), None, x.data_ptr(), stat.data_ptr(), st)),
        "edge_fwd_no_sh_general": lambda: (lib.hsg_set_edge_fwd_lowdeg(0), _lib.check(lib.hsg_edge_fwd(
            C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(), origin.data_ptr(), None, x.data_ptr(),
            stat.data_ptr(), st)), lib.hsg_set_edge_fwd_lowdeg(-1)),
        "edge_bwd_prep": lambda: _lib.check(lib.hsg_edge_bwd_prep(csc.n_dst, H, d, origin.data_ptr(), None, sh.data_ptr(),
                                                                  g.data_ptr(), stat.data_ptr(), st)),
        "edge_bwd": lambda: _lib.check(lib.hsg_edge_bwd(C.byref(csc_t), H, d, zp.data_ptr(), ldz, q.data_ptr(),
                                                        g.data_ptr(), stat.data_ptr(), dzp.data_ptr(), dq.data_ptr(),
                                                        ws.data_ptr(), ws.numel(), st)),
        "edge_bwd_prep_rc": lambda: _lib.check(lib.hsg_edge_bwd_prep_rc(C.byref(csc), H, d, zp.data_ptr(), ldz,
                                                                        q.data_ptr(), origin.data_ptr(), g.data_ptr(),
                                                                        stat.data_ptr(), st))}
    b_fwd = accounting.edge_fwd_bytes_survey(E, csc.n_src, csc.n_dst, H, d)
    b_bwd = accounting.edge_bwd_bytes_survey(E, csc.n_src, csc.n_dst, H, d)
    out = {"graphs": a.graphs, "n_dst": csc.n_dst, "n_src": csc.n_src, "pairs": E,
           "survey_fwd_MB": b_fwd / 1e6, "survey_bwd_MB": b_bwd / 1e6, "hbm_peak": hbm, "ms": {}}
    fns["edge_fwd"]()
    for name, fn in fns.items():
        if a.only and name != a.only:
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        out["ms"][name] = tot / a.iters
    ms = out["ms"]
    fr = lambda b, t: b / (t * 1e-3) / 1e9 / hbm  # noqa: E731
    out["frac_survey"] = {}
    for name in ("edge_fwd", "edge_fwd_no_sh", "edge_fwd_no_sh_general"):
        if name in ms:
            out["frac_survey"][name] = fr(b_fwd, ms[name])
    if "edge_bwd" in ms and "edge_bwd_prep" in ms:
        out["frac_survey"]["edge_bwd_prep+edge_bwd"] = fr(b_bwd, ms["edge_bwd"] + ms["edge_bwd_prep"])
    if "edge_bwd" in ms and "edge_bwd_prep_rc" in ms:
        out["frac_survey"]["edge_bwd_prep_rc+edge_bwd"] = fr(b_bwd, ms["edge_bwd"] + ms["edge_bwd_prep_rc"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
