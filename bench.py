#!/usr/bin/env python
"""bench.py - HSG WSWGAT-path training throughput (graphs/s, fwd+bwd) on B200.

    python bench.py --gpus N --steps K --warmup W            # sm_100a path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm (oracle port of the DGL path)

One "step" = one pass of the hot path over one batch: device-side graph build (K0) from token arrays,
word-embedding lookup, WSWGAT update loop forward (W2S, n_iter x (S2W, W2S)), classifier + the
reference's loss (train.py:114-119), backward, (N>1: one NCCL all-reduce of the flat gradient), fused
Adam step.  The sentence encoder (CNN+BiLSTM, out of the hot-path scope) is replaced by a fixed
random `sent_feature` input.  Workload at N=1: BASELINE.json configs[1] - 32 CNN/DM-shaped graphs,
HSG, n_iter=1, seed 0.  For N>1 every rank gets its own 32-graph shard (weak scaling, sharded by graph).
Dropout is 0 in both arms (parity mode, the reference's semantics are defined deterministically there).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (shape, hdsg, n_iter, seed, BASELINE.json config index)
    "cnndm": ("cnndm", False, 1, 0, 1),
    "nyt50": ("nyt50", False, 3, 1, 2),
    "multinews": ("multinews", True, 1, 2, 3),
}
UNIT = "graphs/s"
METRIC = "HSG train graphs/sec (fwd+bwd), WSWGAT path"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cnndm", choices=list(WORKLOADS))
    ap.add_argument("--graphs-per-gpu", type=int, default=32)
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling (BASELINE.json configs[4]): total graphs per step, split evenly over the GPUs")
    ap.add_argument("--no-stress", action="store_true", help="skip the stress-graph edge-kernel roofline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encoder", action="store_true", help="skip the sentence-encoder leg (SURVEY 8-f rank 1)")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--stress-only", action="store_true", help="run only the stress-graph edge-kernel leg (for ncu)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--no-scaling-legs", action="store_true", help="skip the strong_4096 / weak_1024_per_gpu / grad_parity legs")
    ap.add_argument("--stress-scale", type=int, default=4)
    ap.add_argument("--stress-iters", type=int, default=10)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


def make_workload(args, rank):
    from hetersumgraph_b200 import synthetic as syn
    shape, hdsg, n_iter, seed, cfg_idx = WORKLOADS[args.workload]
    if args.global_batch > 0:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        args.graphs_per_gpu = max(1, args.global_batch // world)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # ONE seeded global batch of graphs_per_gpu x N graphs, dealt to the ranks the way the data-parallel reference
        # would (dist.shard_indices: module/dataloader.py:479 order + snake deal balanced by edge count); its first
        # graphs_per_gpu graphs are the N = 1 batch
        from hetersumgraph_b200.dist import shard_indices
        exs_all = syn.make_examples(args.graphs_per_gpu * world, shape, seed=seed, hdsg=hdsg)
        sh = shard_indices([e.n_sent for e in exs_all], [float(sum(len(x) for x in e.w2s)) for e in exs_all], world)
        exs = [exs_all[i] for i in sh[rank]]
    else:
        exs = syn.make_examples(args.graphs_per_gpu, shape, seed=seed, hdsg=hdsg)
    tb = syn.pack_token_batch(exs, hdsg=hdsg)
    return exs, tb, hdsg, n_iter, cfg_idx


# ------------------------------------------------------------------------------------------------
# reference arm: the oracle port of the reference's CPU path (per head, degree-bucketed like DGL 0.4)
# ------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self, exs, tb, hdsg, n_iter, closed_form=False, device=None):
        from hetersumgraph_b200 import synthetic as syn
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        from hetersumgraph_b200.path_model import HSGPath
        from oracle import graph_builder_ref as gb
        self.closed_form = closed_form
        filt = set(syn.filter_ids().tolist())
        t0 = time.time()
        if hdsg:
            graphs = [gb.create_graph_hdsg(e.doc_len, e.sents.tolist(), e.doc_tokens, e.w2s, e.w2d, filt) for e in exs]
        else:
            graphs = [gb.create_graph_hsg(e.sents.tolist(), e.w2s, filt) for e in exs]
        self.g, _ = gb.collate(graphs, order=tb.order)
        self.builder_s = time.time() - t0
        self.csc = gb.derive_csc(self.g)
        torch.manual_seed(1234)
        m = HSGPath(n_iter=n_iter, hdsg=False)
        self.params = {k[len("loop."):]: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()
                       if k.startswith("loop.")}
        self.wh_w = m.wh.weight.detach().clone().requires_grad_(True)
        self.wh_b = m.wh.bias.detach().clone().requires_grad_(True)
        self.embed = m._embed.weight.detach().clone()
        self.n_iter = n_iter
        self.wid = torch.from_numpy(self.g.wid[self.csc["wnode_id"]])
        ns = int((self.g.unit == 1).sum())
        gen = torch.Generator().manual_seed(7)
        self.sent_feature = torch.randn(ns, 64, generator=gen).requires_grad_(True)
        self.sent_rows = torch.from_numpy(np.nonzero(self.g.ndtype[self.csc["snode_id"]] == 1)[0])
        self.labels = torch.from_numpy(tb.labels)
        self.n_graphs = tb.n_graphs
        if self.device.type != "cpu":            # stock PyTorch on the GPU (B-cf-gpu): everything resident on the device
            d = self.device
            mv = lambda t: t.detach().to(d).requires_grad_(t.requires_grad)   # noqa: E731
            self.params = {k: mv(v) for k, v in self.params.items()}
            self.wh_w, self.wh_b, self.sent_feature = mv(self.wh_w), mv(self.wh_b), mv(self.sent_feature)
            self.embed, self.wid, self.sent_rows, self.labels = (t.to(d) for t in (self.embed, self.wid, self.sent_rows,
                                                                                   self.labels))
            self.csc = {k: (torch.as_tensor(v).to(d) if isinstance(v, np.ndarray) else v) for k, v in self.csc.items()}
        self.opt = torch.optim.Adam(list(self.params.values()) + [self.wh_w, self.wh_b], lr=5e-4)

    def step(self, sync=True):
        from oracle import closed_form as cf
        from oracle import wswgat_ref as wr
        with torch.device(self.device):
            self.opt.zero_grad(set_to_none=True)
            self.sent_feature.grad = None
            wfeat = self.embed[self.wid]
            if self.closed_form:
                _, ss = cf.update_loop_cf(self.csc, wfeat, self.sent_feature, self.params, self.n_iter)
            else:
                _, ss = wr.update_loop(self.g, wfeat, self.sent_feature, self.params, self.n_iter)
            logits = ss[self.sent_rows] @ self.wh_w.t() + self.wh_b
            loss = torch.nn.functional.cross_entropy(logits, self.labels, reduction="sum") / self.n_graphs
            loss.backward()
            self.opt.step()
        return float(loss.detach()) if sync else loss


def time_cpu(ref, steps, warmup):
    for _ in range(warmup):
        ref.step()
    t0 = time.time()
    for _ in range(steps):
        ref.step()
    return (time.time() - t0) / max(steps, 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    exs, tb, hdsg, n_iter, cfg_idx = make_workload(args, 0)
    ref = CpuReference(exs, tb, hdsg, n_iter)
    sec = time_cpu(ref, args.steps, args.warmup)
    value = tb.n_graphs / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args, cfg_idx, n_iter),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "the whole %d-graph batch per step (graph prebuilt; reference CreateGraph restatement took %.2f s once)"
                                   % (tb.n_graphs, ref.builder_s),
                         "what": "oracle/wswgat_ref.py: the reference's per-head, degree-bucketed DGL-0.4 execution restated on torch CPU (real DGL not installable)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_name(args, cfg_idx):
    return ("BASELINE.json configs[%d]: %s WSWGAT update loop fwd+bwd, %d %s-shaped graphs per GPU, seed-0 synthetic "
            "tokens, random-init embeddings, sentence-encoder output replaced by a fixed random sent_feature (the "
            "whole model incl. the encoder: key with_sentence_encoder)"
            % (cfg_idx, "HDSG" if args.workload == "multinews" else "HSG", args.graphs_per_gpu, args.workload))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons of ONE GPU, sampled in-process through NVML every `period` seconds (a Python thread
    calling three NVML getters: no child process, no nvidia-smi start-up on the host cores the ranks share).  Falls back
    to `nvidia-smi -lms 200` when pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period=0.1):
        self.index, self.period = index, period
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc = self.thread = None
        self.stop_flag = threading.Event()
        self.mode = None

    def _nvml_loop(self):
        import pynvml as nv
        h = self.handle
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def _smi_read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.strip().split(",")]
            if len(f) < 6:
                continue
            try:
                self.sm.append(float(f[0]))
                self.mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                parts = [x for x in vis.split(",") if x.strip() != ""]
                if idx < len(parts) and parts[idx].strip().isdigit():
                    idx = int(parts[idx])
            self.handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "nvidia-smi"
            self.thread = threading.Thread(target=self._smi_read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        self.stop_flag.set()
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)
        sm = self.sm
        # "under load": samples at or above 60 % of the highest clock seen (idle gaps between legs clock down)
        hi = [x for x in sm if x >= 0.6 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(hi)) if hi else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(sm), "samples_under_load": len(hi),
                "sampler": self.mode + (" thread, %.0f ms period" % (self.period * 1e3) if self.mode == "nvml" else " -lms 200")}


def base_config(args, cfg_idx, n_iter):
    """the `config` object BOTH arms print (identical keys and values, so the driver can compare them)"""
    return {"workload": workload_name(args, cfg_idx), "graphs_per_gpu": args.graphs_per_gpu, "n_iter": n_iter,
            "dropout": 0.0}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import hetersumgraph_b200 as hb
    from hetersumgraph_b200 import _lib, accounting
    from hetersumgraph_b200.dist import FlatGradArena
    from hetersumgraph_b200.functional import FusedAdam
    from hetersumgraph_b200.graph import BuildPipeline, DeviceTokenBatch
    from hetersumgraph_b200.path_model import FusedTrainStep, HSGPath
    from hetersumgraph_b200.step_graph import GraphedTrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()
    lib = _lib.load()
    pk = peaks()

    exs, tb, hdsg, n_iter, cfg_idx = make_workload(args, rank)
    n_graphs_global = tb.n_graphs * world
    torch.manual_seed(1234)
    model = HSGPath(n_iter=n_iter, hdsg=hdsg).to(dev)
    arena = FlatGradArena(model.parameters(), flatten_params=True)         # contiguous gradient + parameter arenas
    flat = arena.flat
    model.loop.fuse_grad_accumulation = True     # kernels add parameter gradients straight into the arena views
    opt = FusedAdam(arena.flat_param.data, flat, lr=5e-4)                 # torch.optim.Adam semantics, one kernel over the flat arena

    host, h2d_tok_bytes = DeviceTokenBatch.host_buffers(tb)
    bitmap_dev = torch.from_numpy(tb.filter_bitmap.view(np.int32).copy()).to(dev)
    n_sent_rows = int(tb.tokens.shape[0])
    gen = torch.Generator().manual_seed(7 + rank)
    sf_host = torch.randn(n_sent_rows, 64, generator=gen).pin_memory()

    all_reduce = (lambda t: dist.all_reduce(t)) if dist is not None else None
    # N > 1: gradient all-reduce + Adam + zero_grad as ONE kernel over NVLink peer memory (dist.PeerAllReduceAdam);
    # NCCL all-reduce + Adam kernel when symmetric memory is not available (or HSG_PEER_REDUCE=0)
    fused_reduce, reduce_how = None, ("nccl all_reduce + adam kernel" if dist is not None else "single GPU: adam kernel")
    if dist is not None and os.environ.get("HSG_PEER_REDUCE", "1") != "0":
        from hetersumgraph_b200.dist import PeerAllReduceAdam
        try:
            if PeerAllReduceAdam.available():
                fused_reduce = PeerAllReduceAdam(opt)
                reduce_how = "fused peer-memory all-reduce + adam kernel (hsg_allreduce_adam_step)"
        except Exception as e:           # every rank takes the same branch: the constructor is collective
            reduce_how += " (peer path unavailable: %s)" % repr(e)[:120]
    use_graph = not args.no_graph
    # `value` leg: token blob and sent_feature resident in HBM before the timed region; `e2e` leg: both come from
    # pinned host memory every step (H2D inside the step), the loss goes back to pinned host memory every step
    gs_res = GraphedTrainStep(model, opt, bitmap_dev, n_graphs_global, all_reduce, capture=use_graph, resident_tokens=True,
                              fused_reduce=fused_reduce)
    gs_e2e = GraphedTrainStep(model, opt, bitmap_dev, n_graphs_global, all_reduce, capture=use_graph, resident_tokens=False,
                              fused_reduce=fused_reduce)
    gs_res.prime(host)
    gs_e2e.prime(host)
    gs_res._stage_sf(sf_host.to(dev))
    sf_res = gs_res.sf_dev[:n_sent_rows]

    def step_resident():
        return gs_res.step(host, sf_res)

    def step_e2e():
        out = gs_e2e.step(host, sf_host, next_sent_feature=sf_host)     # inputs of step i+1 travel during step i
        return out

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)           # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, prewarm=0):
        for _ in range(prewarm + warmup):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.zero_()                                                    # L2 flush, outside the timed events
            a.record()
            fn()
            b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    def wall(fn, steps):
        """host wall clock around `steps` back-to-back calls, one device sync at the end, no L2 flush"""
        for _ in range(5):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        t = torch.tensor([sec], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    PREWARM = 6      # untimed steps per leg before the driver's --warmup: first step eager, then one capture per slot parity
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_step = timed(step_resident, args.steps, args.warmup, PREWARM)
    ms_e2e = timed(step_e2e, args.steps, args.warmup, PREWARM)
    loss_last = gs_e2e.sync_loss()
    wall_steps = max(500, args.steps)
    sec_wall = wall(step_e2e, wall_steps)
    sec_wall_res = wall(step_resident, wall_steps)
    clk = clocks.stop() if rank == 0 else None
    graph_launches = gs_res.launches_per_step
    replays = (gs_res.replays, gs_e2e.replays)

    # ---- the same step WITHOUT graph replay (round-1 path: FusedTrainStep enqueued from Python every step) ----
    pipe = BuildPipeline(dev)
    dtb = DeviceTokenBatch.upload(tb, dev, host=host, filter_bitmap_dev=bitmap_dev)
    fused_step = FusedTrainStep(model, n_graphs_global)
    sf_dev = sf_host.to(dev)

    def step_eager():
        batch = pipe.take()
        pipe.submit(dtb)
        loss, _logits, _d = fused_step(batch, sf_dev)
        if dist is not None:
            dist.all_reduce(flat)
        opt.step_dev(zero_grad=True)
        pipe.finish()
        return loss, batch

    pipe.submit(dtb)
    pipe.finish()
    l0 = lib.hsg_launch_count()
    ms_eager = timed(step_eager, args.steps, 10)
    eager_launches = (lib.hsg_launch_count() - l0) // (args.steps + 10)
    # informational: the same step with single-pass TF32 products; NOT the headline
    hb.set_gemm_mode("tf32")
    ms_fast = timed(step_eager, args.steps, 5)
    # bf16 projection mode (BASELINE.json north_star "bf16 projections <= 2e-2": tcgen05 kind::f16 on bf16-rounded
    # operands, fp32 TMEM accumulation): the same graph-replayed step, reported NEXT TO the fp32-parity headline
    hb.set_gemm_mode("bf16")
    ms_bf16_eager = timed(step_eager, args.steps, 5)
    gs_bf = GraphedTrainStep(model, opt, bitmap_dev, n_graphs_global, all_reduce, capture=use_graph, resident_tokens=True,
                             fused_reduce=fused_reduce)
    gs_bf.prime(host)
    gs_bf._stage_sf(sf_dev)
    sf_bf = gs_bf.sf_dev[:n_sent_rows]
    ms_bf16 = timed(lambda: gs_bf.step(host, sf_bf), args.steps, args.warmup, PREWARM)
    gs_bf._invalidate()
    hb.set_gemm_mode("tf32x3")

    # ---- per-kernel CUDA-event timing of the same step (roofline leg) ----
    _, batch = step_eager()
    torch.cuda.synchronize()
    lib.hsg_set_bwd_overlap(0)      # per-kernel events are only meaningful when the kernels do not share the GPU
    lib.hsg_profile_reset()
    lib.hsg_profile_enable(1)
    for _ in range(args.steps):
        flush.zero_()
        step_eager()
    torch.cuda.synchronize()
    lib.hsg_profile_enable(0)
    lib.hsg_set_bwd_overlap(1)
    prof = _lib.profile_snapshot()
    acct = dict(accounting.step_accounting(batch.n_word, batch.n_super, batch.n_pair, n_iter))
    bb = accounting.builder_bytes(n_sent_rows, int(tb.tokens.shape[1]), batch.n_pair, batch.n_word, batch.n_super)
    acct["build_count"] = (0, bb // 2, 1)
    acct["build_fill"] = (0, bb - bb // 2, 1)
    acct["build_scan"] = (0, 10 * (tb.n_graphs + 1) * 4, 1)
    acct["adam"] = (0, flat.numel() * 4 * 7, 1)
    kernels = []
    for name, (cnt, ms) in prof.items():
        per_step_ms = ms / args.steps
        fl, by, _n = acct.get(name, (0, 0, 0))
        row = {"kernel": name, "launches_per_step": cnt / args.steps, "ms_per_step": per_step_ms,
               "share": None, "flops": fl, "bytes": by}
        if fl > 0:
            row.update(bound="tensor", achieved=fl / (per_step_ms * 1e-3) / 1e12, peak=pk["tf"], unit="TFLOP/s")
        elif by > 0:
            row.update(bound="hbm", achieved=by / (per_step_ms * 1e-3) / 1e9, peak=pk["hbm"], unit="GB/s")
        if "achieved" in row:
            row["frac"] = row["achieved"] / row["peak"]
        kernels.append(row)
    tot = sum(k["ms_per_step"] for k in kernels) or 1.0
    for k in kernels:
        k["share"] = k["ms_per_step"] / tot
    kernels.sort(key=lambda k: -k["ms_per_step"])
    step_bytes = sum(v[1] for v in acct.values())
    step_flops = sum(v[0] for v in acct.values())
    # ---- roofline of the dominant kernel: the tcgen05 GEMM at its largest launch shape (FFN-1 on the word nodes,
    # GATLayer.py:38) timed live, one launch at a time, L2 flushed, CUDA events on the launching stream ----
    from hetersumgraph_b200.functional import gemm_nt
    from hetersumgraph_b200._lib import EPI_BIAS, EPI_RELU
    Mw, Kw, Nh = batch.n_word, 300, 512
    xa = torch.randn(Mw, Kw, device=dev)
    wb = torch.randn(Nh, Kw, device=dev) * 0.05
    bb_ = torch.randn(Nh, device=dev)
    outb = torch.empty(Mw, Nh, device=dev)
    for _ in range(3):
        gemm_nt(xa, wb, bias=bb_, epi=EPI_BIAS | EPI_RELU, out=outb)
    torch.cuda.synchronize()
    tsum = 0.0
    n_it = 20
    for _ in range(n_it):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gemm_nt(xa, wb, bias=bb_, epi=EPI_BIAS | EPI_RELU, out=outb)
        e1.record()
        torch.cuda.synchronize()
        tsum += e0.elapsed_time(e1)
    ms_gemm_single = tsum / n_it
    # (b) the same launch back to back over rotating operand sets larger than L2 (5 x (A 14 MB + C 24 MB) = 192 MB):
    # two events around 40 launches.  A single launch between two events after an L2 flush (a) also contains the
    # launch latency of an idle stream (~8 us here: the kernel's own first-CTA-start -> last-CTA-end span, read from
    # %globaltimer inside the kernel, is 32.5 us where (a) says 40); inside the step the launches queue behind each
    # other (graph replay + programmatic dependent launch), which is what (b) reproduces.
    n_sets, n_rot = 5, 40
    xas = [torch.randn(Mw, Kw, device=dev) for _ in range(n_sets)]
    outs_ = [torch.empty(Mw, Nh, device=dev) for _ in range(n_sets)]
    for i in range(n_sets):
        gemm_nt(xas[i], wb, bias=bb_, epi=EPI_BIAS | EPI_RELU, out=outs_[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_rot):
        gemm_nt(xas[i % n_sets], wb, bias=bb_, epi=EPI_BIAS | EPI_RELU, out=outs_[i % n_sets])
    e1.record()
    torch.cuda.synchronize()
    ms_gemm = e0.elapsed_time(e1) / n_rot
    # (c) GPU-side span of one launch (what ncu's gpu__time_duration reports), L2 flushed
    import ctypes as _C
    us_span = None
    try:
        lib.hsg_gemm_pair_trace(1, None, 0)
        flush.zero_()
        gemm_nt(xa, wb, bias=bb_, epi=EPI_BIAS | EPI_RELU, out=outb)
        torch.cuda.synchronize()
        sp = (_C.c_ulonglong * (3 * 160))()
        lib.hsg_gemm_pair_trace(-2, sp, 3 * 160)
        lib.hsg_gemm_pair_trace(0, None, 0)
        spn = np.frombuffer(sp, dtype=np.uint64).reshape(-1, 3).astype(np.int64)
        spn = spn[spn[:, 0] > 0]
        if len(spn):
            us_span = float(spn[:, 1].max() - spn[:, 0].min()) / 1e3
    except Exception:
        us_span = None
    del xas, outs_
    fl_gemm = 2.0 * Mw * Kw * Nh
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("gemm_nt_ffn1_words", {}).get("bytes")
    gemm_share = sum(k["share"] for k in kernels if k["kernel"].startswith("gemm"))
    roofline = {"kernel": "gemm_tc2_kernel<false> (hsg_gemm_nt: CTA-pair tcgen05 cta_group::2 kind::tf32, 3-product "
                          "fp32-parity mode), FFN-1 on the word nodes M=%d N=%d K=%d" % (Mw, Nh, Kw),
                "bound": "tensor", "achieved": fl_gemm / (ms_gemm * 1e-3) / 1e12, "peak": pk["tf"], "unit": "TFLOP/s",
                "frac": fl_gemm / (ms_gemm * 1e-3) / 1e12 / pk["tf"], "traffic": traffic, "peak_source": pk["src"],
                "us_per_launch": ms_gemm * 1e3, "us_single_launch_after_l2_flush": ms_gemm_single * 1e3,
                "us_gpu_span_one_launch_l2_flushed": us_span, "algorithmic_flops_per_launch": fl_gemm,
                "algorithmic_bytes_per_launch": 4.0 * (Mw * Kw + Nh * Kw + Mw * Nh),
                "share_of_kernel_time_all_gemm_slots": gemm_share,
                "note": "peak = measured sustained bf16 cuBLAS TFLOP/s; the fp32-parity scheme issues 3 TF32 MMAs per "
                        "product, so its ceiling is TF32-peak/3 (about 0.27 of this peak); "
                        "us_per_launch: 40 launches back to back between two CUDA events over 5 rotating operand "
                        "sets (192 MB > 126 MB L2); us_single_launch_after_l2_flush: one launch between two events after "
                        "a 256 MiB L2 flush (contains the launch latency of an idle stream); "
                        "us_gpu_span_one_launch_l2_flushed: first CTA start -> last CTA end from %globaltimer inside the "
                        "kernel (the figure ncu reports as gpu__time_duration)"}
    del xa, wb, bb_, outb

    # ---- multi-GPU legs of BASELINE.json configs[4]: ONE global batch dealt by dist.shard_indices ----
    scaling_legs = grad_par = None
    if not args.no_scaling_legs:
        scaling_legs, grad_par = scaling_legs_run(args, dev, rank, world, dist, timed)

    if rank != 0:
        _shutdown(dist, (gs_res, gs_e2e))
        return

    stress = None
    if not args.no_stress and world == 1:
        stress = []
        for sc in (1, 4, 16):
            stress += stress_leg(dev, pk, sc, args.stress_iters, flush)
    large = None
    if not args.no_stress and world == 1:
        large = large_shard_leg(dev, pk)
    encoder = None
    if not args.no_encoder and world == 1:
        encoder = encoder_leg(dev, pk, exs, tb, n_iter, args.steps, hdsg=hdsg)
    stock = None
    if not args.no_cpu_baseline and world == 1:
        stock = stock_pytorch_gpu_leg(dev, exs, tb, hdsg, n_iter, flush, args.steps)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        ref = CpuReference(exs, tb, hdsg, n_iter)
        sec = time_cpu(ref, args.cpu_steps, 1)
        cpu = {"value": tb.n_graphs / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "the whole %d-graph batch, %d timed steps after 1 warm-up (%.2f s/step); graph prebuilt "
                         "(reference CreateGraph restatement: %.2f s once)" % (tb.n_graphs, args.cpu_steps, sec, ref.builder_s)}
        refc = CpuReference(exs, tb, hdsg, n_iter, closed_form=True)
        secc = time_cpu(refc, args.cpu_steps, 1)
        cpu["closed_form_graphs_per_s"] = tb.n_graphs / secc
        cpu["reference_verbatim_on_shim"] = verbatim_reference_leg(exs, tb, hdsg, n_iter)

    d2h = 4 + 4 * (5 * (tb.n_graphs + 1) + 1)
    line = {
        "metric": METRIC, "value": n_graphs_global / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args, cfg_idx, n_iter),
        "details": {"l2": "flushed between steps (256 MiB memset outside the timed events)",
                    "step": ("one cudaGraphLaunch per step (step_graph.GraphedTrainStep): build of batch i+1 on a forked "
                             "branch + embedding gather, update loop fwd, loss, update loop bwd, gradient reduce + "
                             "Adam + zero_grad of batch i") if use_graph else "eager enqueue (--no-graph)",
                    "gradient_reduce": reduce_how,
                    "batch": ("one seeded global batch of %d graphs dealt to the %d ranks by dist.shard_indices "
                              "(dataloader.py:479 order + snake deal)" % (n_graphs_global, world)) if world > 1
                             else "one seeded batch",
                    "graph_replays_resident_e2e": replays,
                    "prewarm_steps_per_leg": PREWARM,
                    "build": "device-side (K0), double-buffered slots: batch i+1 is built while batch i computes; exactly one build per timed step",
                    "gemm_mode": "tf32x3 (tcgen05 kind::tf32, hi/lo split: fp32-parity mode)",
                    "rank0_sizes": {"word_nodes": batch.n_word, "supernodes": batch.n_super, "pairs_per_direction": batch.n_pair,
                                    "dgl_edges": batch.n_total_edges},
                    "edges_per_s": 2 * batch.n_pair * (1 + 2 * n_iter) * world / (ms_step * 1e-3),
                    "step_algorithmic_MB": step_bytes / 1e6, "step_algorithmic_GFLOP": step_flops / 1e9,
                    "step_frac_of_hbm_peak": step_bytes / (ms_step * 1e-3) / 1e9 / pk["hbm"],
                    "step_frac_of_tensor_peak": step_flops / (ms_step * 1e-3) / 1e12 / pk["tf"],
                    "last_loss": loss_last},
        "e2e": {"value": n_graphs_global / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h2d_tok_bytes + sf_host.numel() * 4), "d2h_bytes_per_step": d2h,
                "d2h": "loss -> pinned host memory every step (inside the step), read by the host after the next step; "
                       "builder totals of the next batch read every step",
                "wall_clock": {"steps": wall_steps, "seconds": sec_wall, "ms_per_step": sec_wall / wall_steps * 1e3,
                               "graphs_per_s": n_graphs_global * wall_steps / sec_wall,
                               "what": "time.perf_counter around %d back-to-back e2e steps (host memcpy into pinned staging, "
                                       "H2D, step, D2H), one device sync at the end, no L2 flush" % wall_steps},
                "wall_clock_resident": {"steps": wall_steps, "seconds": sec_wall_res,
                                        "graphs_per_s": n_graphs_global * wall_steps / sec_wall_res}},
        "gpu_launches": int(graph_launches if graph_launches is not None else eager_launches),
        "eager_path": {"ms_per_step": ms_eager, "graphs_per_s": n_graphs_global / (ms_eager * 1e-3),
                       "launches_per_step": int(eager_launches),
                       "what": "the same step enqueued from Python every step (FusedTrainStep + BuildPipeline), resident inputs"},
        "clocks": clk, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "stock_pytorch_gpu": stock,
        "edge_kernels_stress": stress, "large_shard": large, "with_sentence_encoder": encoder,
        "scaling_legs": scaling_legs, "grad_parity": grad_par,
        "bf16_mode": {"graphs_per_s": n_graphs_global / (ms_bf16 * 1e-3), "ms_per_step": ms_bf16,
                      "eager_ms_per_step": ms_bf16_eager,
                      "note": "hsg_set_gemm_mode(3): every tensor-core product (fc / FFN linears and their input- and "
                              "weight-gradient products) as ONE tcgen05.mma kind::f16 on bf16-rounded operands with fp32 "
                              "accumulation; tolerance class 2e-2 forward / 5e-2 gradients (tests); activations and "
                              "weights stay fp32 in HBM; secondary figure, never the headline"},
        "single_pass_tf32_mode": {"graphs_per_s": n_graphs_global / (ms_fast * 1e-3), "ms_per_step": ms_fast,
                                  "note": "informational (eager path), tolerance class 2e-2; not the headline"},
    }
    print(json.dumps(line))
    _shutdown(dist, (gs_res, gs_e2e))


def _shutdown(dist, graphed_steps):
    """End of a rank.  The captured step graphs contain ncclAllReduce nodes: they are destroyed and the device is drained
    BEFORE the process group goes away (tearing the communicator down under live graphs left both ranks hanging in
    destroy_process_group on the 2-GPU box: r02m), and the teardown itself is bounded - a rank that cannot finish it
    within 20 s exits anyway, its line is already printed."""
    sys.stdout.flush()
    for gs in graphed_steps:
        gs._invalidate()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    if dist is None:
        return
    dist.barrier()
    torch.cuda.synchronize()
    import threading
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(20.0)
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def scaling_legs_run(args, dev, rank, world, dist, timed):
    """BASELINE.json configs[4] in a driver-visible form: ONE seeded global batch, sorted and dealt to the ranks by
    dist.shard_indices (dataloader.py:479-480 ordering + snake deal), each rank computes its shard with the loss scaled
    by 1/B_global, one all-reduce of the arena, Adam.  strong: 4 096 graphs per step in total; weak: 1 024 per GPU.
    grad_parity: the all-reduced gradient arena against rank 0 processing the whole (smaller) global batch alone."""
    from hetersumgraph_b200 import synthetic as syn
    from hetersumgraph_b200.dist import FlatGradArena, shard_indices
    from hetersumgraph_b200.functional import FusedAdam
    from hetersumgraph_b200.graph import HeteroBatch
    from hetersumgraph_b200.path_model import FusedTrainStep, HSGPath

    def model_pair():
        torch.manual_seed(1234)
        m = HSGPath(n_iter=1).to(dev)
        ar = FlatGradArena(m.parameters(), flatten_params=True)
        m.loop.fuse_grad_accumulation = True
        return m, ar, FusedAdam(ar.flat_param.data, ar.flat, lr=5e-4)

    def shard_of(exs_all, r, w):
        sh = shard_indices([e.n_sent for e in exs_all], [float(sum(len(x) for x in e.w2s)) for e in exs_all], w)
        return [exs_all[i] for i in sh[r]]

    base = syn.make_examples(512, "cnndm", seed=3)     # the global batch tiles 512 distinct graphs (generation cost)

    def leg(n_global, steps):
        exs_all = [base[i % len(base)] for i in range(n_global)]
        mine = shard_of(exs_all, rank, world)
        tbs = syn.pack_token_batch(mine)
        batch = HeteroBatch.from_token_batch(tbs, dev)
        m, ar, op = model_pair()
        step = FusedTrainStep(m, n_global)
        sf = torch.randn(int(tbs.tokens.shape[0]), 64, device=dev, generator=torch.Generator(device=dev).manual_seed(5 + rank))

        def fn():
            step(batch, sf)
            if dist is not None:
                dist.all_reduce(ar.flat)
            op.step_dev(zero_grad=True)
        ms = timed(fn, steps, 3)
        out = {"global_graphs": n_global, "graphs_this_rank": len(mine), "word_nodes_this_rank": batch.n_word,
               "ms_per_step": ms, "graphs_per_s": n_global / (ms * 1e-3),
               "what": "token arrays and the built batch resident; update loop fwd+bwd + loss + all-reduce + Adam per step"}
        del batch, m, ar, op, step, sf
        torch.cuda.empty_cache()
        return out

    legs = {"strong_4096": leg(4096, 5), "weak_1024_per_gpu": leg(1024 * world, 5)}

    # gradient parity on a 256-graph global batch (base[:256] is the same seeded batch)
    from hetersumgraph_b200.dist import gradient_parity
    par = gradient_parity(rank, world, dev, (lambda t: dist.all_reduce(t)) if dist is not None else None, 256, seed=3)
    return legs, par


def stock_pytorch_gpu_leg(dev, exs, tb, hdsg, n_iter, flush, steps):
    """B-cf-gpu of BASELINE.md: the closed-form restatement (oracle/closed_form.py: index_add / scatter softmax in stock
    PyTorch, no custom kernels) on the same B200 and batch - what eager PyTorch does with this path on this GPU."""
    ref = CpuReference(exs, tb, hdsg, n_iter, closed_form=True, device=dev)
    for _ in range(3):
        ref.step()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ref.step(sync=False)
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    ms = tot / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.step(sync=False)
    torch.cuda.synchronize()
    ms_wall = (time.perf_counter() - t0) / steps * 1e3
    return {"what": "oracle/closed_form.py (vectorised torch restatement, fp32, TF32 off) fwd+bwd+torch.optim.Adam on the B200, "
                    "stock PyTorch kernels only; graph prebuilt on the host",
            "ms_per_step_events": ms, "ms_per_step_wall": ms_wall, "graphs_per_s": tb.n_graphs / (max(ms, ms_wall) * 1e-3)}


def verbatim_reference_leg(exs, tb, hdsg, n_iter, steps=1):
    """B-ref of BASELINE.md: the reference's own WSWGAT modules (module/GAT.py:45-59) UNMODIFIED on the DGL-0.4 shim,
    1 thread and all threads.  Only where /root/reference exists (the build container); None on the GPU box."""
    if not os.path.isdir("/root/reference/module"):
        return None
    try:
        from oracle import fixtures as fx
        return fx.time_verbatim_reference(exs, tb, hdsg, n_iter, steps)
    except Exception as e:          # test infrastructure only: never fail the bench line
        return {"unavailable": repr(e)[:200]}


def encoder_leg(dev, pk, exs, tb, n_iter, steps, cpu_steps=1, hdsg=False):
    """the sentence encoder (n-gram CNN + BiLSTM + projections) in front of the path, same batch.
    (i) encoder alone fwd+bwd, (ii) encoder -> update loop -> loss, backward through both (autograd path, parameter
    gradients of every stage), (iii) the oracle restatement of the encoder on the host cores, (iv) per-kernel split."""
    from hetersumgraph_b200 import _lib
    from hetersumgraph_b200.encoder import EncoderPlan, SentenceEncoder
    from hetersumgraph_b200.graph import DeviceTokenBatch, HeteroBatch
    from hetersumgraph_b200.path_model import HSGPath, fused_loss
    lib = _lib.load()
    torch.manual_seed(1234)
    model = HSGPath(n_iter=n_iter, hdsg=hdsg).to(dev)
    enc = SentenceEncoder(model._embed, lstm_dropout=0.0).to(dev)
    dtb = DeviceTokenBatch.upload(tb, dev)
    batch = HeteroBatch.build(dtb)
    S, L = tb.tokens.shape
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cot = torch.randn(S, 64, device=dev)
    params = [q for q in list(enc.parameters()) + list(model.parameters()) if q.requires_grad]

    def enc_only():
        plan = EncoderPlan.from_token_batch(tb, dev, tokens_dev=dtb.tokens)      # host plan + one small H2D per step
        enc(plan).backward(cot)

    def full():
        plan = EncoderPlan.from_token_batch(tb, dev, tokens_dev=dtb.tokens)
        for q in params:
            q.grad = None
        loss, _ = fused_loss(model, batch, enc(plan))
        loss.backward()
        return loss

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / steps

    ms_enc = timed(enc_only)
    ms_full = timed(full)

    # ---- the whole model through the public drop-in class, one TRAINING step from host buffers: pinned-host token blob
    # -> H2D -> device graph build -> HSumGraph.forward (encoder, embedding, update loop, classifier) -> the reference's
    # loss (train.py:114-119) -> backward -> Adam over the flat arena; the loss goes back to pinned host memory ----
    import types

    import hetersumgraph_b200 as hb
    from hetersumgraph_b200.dist import FlatGradArena
    from hetersumgraph_b200.functional import FusedAdam
    hps = types.SimpleNamespace(n_iter=n_iter, word_emb_dim=300, sent_max_len=L, doc_max_timesteps=50, n_feature_size=128,
                                hidden_size=64, lstm_hidden_state=128, lstm_layers=2, bidirectional=True, n_head=8,
                                atten_dropout_prob=0.0, ffn_inner_hidden_size=512, ffn_dropout_prob=0.0,
                                feat_embed_size=50, cuda=True)
    torch.manual_seed(1234)
    embed = torch.nn.Embedding(50000, 300, padding_idx=0)
    embed.weight.requires_grad_(False)
    whole = (hb.HSumDocGraph if hdsg else hb.HSumGraph)(hps, embed).to(dev)
    whole.lstm.dropout = 0.0
    arena = FlatGradArena(whole.parameters(), flatten_params=True)
    whole.fuse_grad_accumulation = True    # kernels add parameter gradients straight into the arena views
    opt = FusedAdam(arena.flat_param.data, arena.flat, lr=5e-4)
    host, h2d_bytes = DeviceTokenBatch.host_buffers(tb)
    bitmap_dev = torch.from_numpy(tb.filter_bitmap.view(np.int32).copy()).to(dev)
    loss_host = torch.zeros(1).pin_memory()
    from hetersumgraph_b200.graph import BuildPipeline
    pipe = BuildPipeline(dev)      # upload + build of batch i+1 on a side stream while batch i trains (one build per step)

    def upload():
        return DeviceTokenBatch.upload(tb, dev, host=host, filter_bitmap_dev=bitmap_dev)

    def train_step():
        b = pipe.take()
        pipe.submit(upload)
        arena.flat.zero_()
        loss, _logits = whole.loss(b)                 # forward + the reference's loss (train.py:113-119)
        loss.backward()
        opt.step()
        loss_host.copy_(loss.detach().view(1), non_blocking=True)
        pipe.finish()
        return loss

    pipe.submit(upload)
    pipe.finish()
    ms_train = timed(train_step)
    lib.hsg_profile_reset()
    lib.hsg_profile_enable(1)
    for _ in range(steps):
        flush.zero_()
        enc_only()
    torch.cuda.synchronize()
    lib.hsg_profile_enable(0)
    slots = {k: {"launches_per_step": v[0] / steps, "ms_per_step": v[1] / steps} for k, v in _lib.profile_snapshot().items()}
    plan = EncoderPlan.from_token_batch(tb, dev, tokens_dev=dtb.tokens)
    useful = S * sum((L + 1 - h) * h for h in range(2, 8)) * 50 * 300 * 2.0      # what the reference's Conv2d executes
    executed = plan.n_rows * 300.0 * 2.0 * (312 * 2 + 260 * 2 + 156 * 2 + 52)
    out = {"what": "sentEncoder + cnn_proj + BiLSTM + lstm_proj + n_feature_proj (Encoder.py:56-76, HiGraph.py:112-161, :96) "
                   "fwd+bwd on the same %d-graph batch; L2 flushed; inter-layer LSTM dropout 0 in both arms" % tb.n_graphs,
           "sentences": S, "compact_rows": plan.n_rows, "padded_rows": S * L,
           "encoder_fwd_bwd_ms": ms_enc, "encoder_graphs_per_s": tb.n_graphs / (ms_enc * 1e-3),
           "encoder_plus_path_fwd_bwd_ms": ms_full, "encoder_plus_path_graphs_per_s": tb.n_graphs / (ms_full * 1e-3),
           "whole_model_train_step": {
               "what": "hetersumgraph_b200.HSumGraph / HSumDocGraph (drop-ins for HiGraph.HSumGraph / HSumDocGraph) from HOST buffers: H2D of the token "
                       "blob + device graph build of the NEXT batch on a side stream (one per step), forward, loss, "
                       "backward, Adam; loss copied to pinned host memory every step",
               "ms_per_step": ms_train, "graphs_per_s": tb.n_graphs / (ms_train * 1e-3),
               "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 4 + 4 * (5 * (tb.n_graphs + 1) + 1),
               "trainable_parameters": int(arena.flat.numel())},
           "conv_fwd_gflop_reference": useful / 1e9, "conv_fwd_gflop_executed": executed / 1e9, "kernels": slots}
    # host-cores arm: the oracle restatement of the same stage (oracle/encoder_ref.py), forward + backward
    from oracle import encoder_ref as er
    frozen = ("ngram_enc.embed.weight", "sent_pos_embed.weight", "ngram_enc.position_embedding.weight")
    sd = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    p = {k: (v.clone().requires_grad_(True) if k not in frozen else v) for k, v in sd.items()}
    cot_c = cot.cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    for _ in range(cpu_steps):
        sf, _ = er.sent_feature(tb.tokens, tb.graph_sent_ptr, p)
        sf.backward(cot_c)
    sec = (time.perf_counter() - t0) / cpu_steps
    out["cpu_baseline"] = {"value": tb.n_graphs / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                           "sample": "the whole %d-graph batch, %d step(s), %.2f s/step" % (tb.n_graphs, cpu_steps, sec)}
    return out


def _time_edge_kernels(batch, label, dev, pk, flush, iters):
    """the three edge kernels of both layer types on `batch`, each timed alone with CUDA events (L2 flushed)."""
    import ctypes as C

    from hetersumgraph_b200 import _lib, accounting
    from hetersumgraph_b200.functional import _Workspace
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for kind, H, d in (("W2S", 8, 8), ("S2W", 6, 50)):
        csc, csc_t = batch.csc(kind)
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
                                                            origin.data_ptr(), sh.data_ptr(), x.data_ptr(),
                                                            stat.data_ptr(), st)),
            "edge_bwd_prep": lambda: _lib.check(lib.hsg_edge_bwd_prep(csc.n_dst, H, d, origin.data_ptr(), None,
                                                                      sh.data_ptr(), g.data_ptr(), stat.data_ptr(), st)),
            "edge_bwd": lambda: _lib.check(lib.hsg_edge_bwd(C.byref(csc_t), H, d, zp.data_ptr(), ldz, q.data_ptr(),
                                                            g.data_ptr(), stat.data_ptr(), dzp.data_ptr(), dq.data_ptr(),
                                                            ws.data_ptr(), ws.numel(), st))}
        nb = {"edge_fwd": accounting.edge_fwd_bytes(E, csc.n_src, csc.n_dst, H, d),
              "edge_bwd_prep": accounting.edge_bwd_prep_bytes(csc.n_dst, H, d),
              "edge_bwd": accounting.edge_bwd_bytes(E, csc.n_src, csc.n_dst, H, d)}
        if lib.hsg_edge_bwd_prep_rc_ok(H, d, ldz):
            # the recomputing path (csrc/hsg_edge_rc.cu; what the update loop takes from 65 536 destination rows on):
            # the forward does not store sh, the backward prep recomputes it from (m, den) and the source rows
            fns["edge_fwd_no_sh"] = lambda: _lib.check(lib.hsg_edge_fwd(
                C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(), origin.data_ptr(), None, x.data_ptr(),
                stat.data_ptr(), st))
            fns["edge_bwd_prep_rc"] = lambda: _lib.check(lib.hsg_edge_bwd_prep_rc(
                C.byref(csc), H, d, zp.data_ptr(), ldz, q.data_ptr(), origin.data_ptr(), g.data_ptr(), stat.data_ptr(),
                st))
            nb["edge_fwd_no_sh"] = accounting.edge_fwd_bytes_survey(E, csc.n_src, csc.n_dst, H, d)
            nb["edge_bwd_prep_rc"] = accounting.edge_bwd_prep_rc_bytes(E, csc.n_src, csc.n_dst, H, d)
        out += _edge_rows(fns, nb, label, kind, H, d, E, csc, pk, flush, iters)
        del zp, origin, sh, x, g, dzp
    return out


def _edge_rows(fns, nb, label, kind, H, d, E, csc, pk, flush, iters):
    """time each edge kernel alone (CUDA events, L2 flushed before every launch) -> rows with two fractions of the HBM
    peak: `frac_of_hbm_peak` over the bytes the kernel really moves (incl. the saved `sh`), and `frac_survey` over
    SURVEY.md 8(d)'s B_fwd / B_bwd (the backward figure covers prep + the source-centric pass together)."""
    from hetersumgraph_b200 import accounting
    ms_of = {}
    for name, fn in fns.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        ms_of[name] = tot / iters
    rows = []
    b_fwd = accounting.edge_fwd_bytes_survey(E, csc.n_src, csc.n_dst, H, d)
    b_bwd = accounting.edge_bwd_bytes_survey(E, csc.n_src, csc.n_dst, H, d)
    for name, ms in ms_of.items():
        gbs = nb[name] / (ms * 1e-3) / 1e9
        row = {"input": label, "kernel": name, "layer": kind, "heads": H, "head_dim": d, "pairs": E,
               "n_src": csc.n_src, "n_dst": csc.n_dst, "ms": ms, "algorithmic_MB": nb[name] / 1e6,
               "GBps": gbs, "frac_of_hbm_peak": gbs / pk["hbm"], "l2": "flushed before every launch"}
        if name in ("edge_fwd", "edge_fwd_no_sh"):
            row["survey_MB"] = b_fwd / 1e6
            row["frac_survey"] = b_fwd / (ms * 1e-3) / 1e9 / pk["hbm"]
        rows.append(row)
    if "edge_bwd" in ms_of:
        for prep in ("edge_bwd_prep", "edge_bwd_prep_rc"):
            if prep in ms_of:
                ms_pair = ms_of["edge_bwd"] + ms_of[prep]
                rows.append({"input": label, "kernel": prep + "+edge_bwd", "layer": kind, "heads": H, "head_dim": d,
                             "pairs": E, "ms": ms_pair, "survey_MB": b_bwd / 1e6,
                             "frac_survey": b_bwd / (ms_pair * 1e-3) / 1e9 / pk["hbm"]})
    return rows


def large_shard_leg(dev, pk, n_graphs=2048, iters=10):
    """What a data-parallel shard looks like (config 5): n_graphs CNN/DM-shaped graphs on one GPU.  Edge kernels
    alone (HBM roofline) and the whole fwd+bwd step (graphs/s when the GPU, not the host, is the limit)."""
    import hetersumgraph_b200 as hb
    from hetersumgraph_b200 import synthetic as syn
    from hetersumgraph_b200.graph import DeviceTokenBatch, HeteroBatch
    from hetersumgraph_b200.path_model import HSGPath, graph_loss
    exs = syn.make_examples(n_graphs, "cnndm", seed=3)
    tb = syn.pack_token_batch(exs)
    dtb = DeviceTokenBatch.upload(tb, dev)
    batch = HeteroBatch.build(dtb)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    edges = _time_edge_kernels(batch, "%d cnndm graphs" % n_graphs, dev, pk, flush, iters)
    torch.manual_seed(1234)
    model = HSGPath(n_iter=1).to(dev)
    sf = torch.randn(tb.tokens.shape[0], 64, device=dev)

    def step():
        b = HeteroBatch.build(dtb)
        loss = graph_loss(b, model(b, sf.detach().requires_grad_(True)), b.labels)
        for p in model.parameters():
            p.grad = None
        loss.backward()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    return {"n_graphs": n_graphs, "word_nodes": batch.n_word, "supernodes": batch.n_super, "pairs": batch.n_pair,
            "step_ms": ms, "graphs_per_s": n_graphs / (ms * 1e-3), "edge_kernels": edges,
            "what": "device build + update loop fwd+bwd + loss (no optimizer), inputs resident"}


def stress_leg(dev, pk, scale=4, iters=10, flush=None):
    """Edge kernels on the stress graph of SURVEY 8-d (x1, x4, x16), each timed alone with CUDA events, L2 flushed before
    every launch."""
    import hetersumgraph_b200 as hb
    from hetersumgraph_b200 import synthetic as syn
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_word, n_super, n_edges = 262144 * scale, 32768 * scale, 1048576 * scale
    word, sup, bins, extra = syn.stress_edges(n_word, n_super, n_edges, seed=4, extra=64)
    (sip, ssrc, sbin, _), (wip, wsrc, wbin, _) = hb.csc_pair_from_edges(word, sup, bins, n_word, n_super)
    batch = hb.HeteroBatch.from_csc_arrays(sip, ssrc, sbin, extra, wip, wsrc, wbin, device=dev)
    out = _time_edge_kernels(batch, "stress graph x%d" % scale, dev, pk, flush, iters)
    # what the kernels REALLY move on a random graph whose gathered set exceeds L2 (ncu capture, profiles/traffic.json):
    # every gather is a DRAM miss, so SURVEY 8(d)'s compulsory count (each source row once) is out of reach there
    try:
        seen = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("stress_x%d" % scale, {})
    except (OSError, ValueError):
        seen = {}
    for row in out:
        cap = seen.get("%s/%s" % (row.get("layer"), row.get("kernel")))
        if cap:
            nbytes = cap["dram_read"] + cap["dram_write"]
            row["ncu_dram_MB"] = nbytes / 1e6
            row["frac_of_hbm_peak_ncu_bytes"] = nbytes / (row["ms"] * 1e-3) / 1e9 / pk["hbm"]
    del batch
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    a = parse()
    if a.stress_only:
        torch.cuda.set_device(0)
        print(json.dumps({"edge_kernels_stress": stress_leg(torch.device("cuda", 0), peaks(), a.stress_scale, a.stress_iters)}))
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
