"""CPU tests of the host side: C-ABI library loads and exports every declared symbol,
state_dict compatibility with the reference's per-head keys, error behaviour without a GPU,
synthetic generator / packer invariants."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import hetersumgraph_b200 as hb
from hetersumgraph_b200 import _lib
from hetersumgraph_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "hsg_b200.h")).read()
    declared = set(re.findall(r"\b(hsg_[a-z0-9_]+)\s*\(", header))
    declared -= {"hsg_status"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "libhsg_b200.so does not export %s" % name
    assert set(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.hsg_version() == 1
    assert lib.hsg_strerror(0) == b"ok"
    assert b"workspace" in lib.hsg_strerror(-4)
    assert lib.hsg_profile_num_slots() > 0
    # struct layouts agree with the header (pointer-size sanity)
    assert ctypes.sizeof(_lib.CscC) == 16 + 4 * 8
    assert ctypes.sizeof(_lib.TokenBatchC) == 32 + 9 * 8
    assert ctypes.sizeof(_lib.GraphOutC) == 16 + 5 * 8 + 15 * 8


def test_ctypes_struct_layouts_match_the_library():
    """Every argument structure of include/hsg_b200.h, in declaration order, against the ctypes mirror in _lib.py: the
    sizes the compiled library reports (hsg_abi_sizeof) are the sizes the binding passes, and the header declares exactly
    these structures."""
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "hsg_b200.h")).read()
    declared = [n for n in re.findall(r"^\}\s*(hsg_[a-z0-9_]+);", header, flags=re.M) if n != "hsg_status"]
    mirror = [_lib.TokenBatchC, _lib.GraphOffsetsC, _lib.CscC, _lib.GraphOutC, _lib.WswgatFwdArgsC, _lib.WswgatBwdArgsC,
              _lib.LayerParamsC, _lib.LayerGradsC, _lib.LoopArgsC, _lib.LoopPlanC, _lib.LoopBwdArgsC, _lib.HeadArgsC,
              _lib.S2SGraphC, _lib.DocMapC]
    assert declared == ["hsg_token_batch", "hsg_graph_offsets", "hsg_csc", "hsg_graph_out", "hsg_wswgat_fwd_args",
                        "hsg_wswgat_bwd_args", "hsg_layer_params", "hsg_layer_grads", "hsg_loop_args", "hsg_loop_plan",
                        "hsg_loop_bwd_args", "hsg_head_args", "hsg_s2s_graph", "hsg_doc_map"]
    for i, (name, cls) in enumerate(zip(declared, mirror)):
        assert lib.hsg_abi_sizeof(i) == ctypes.sizeof(cls), (name, lib.hsg_abi_sizeof(i), ctypes.sizeof(cls))
    assert lib.hsg_abi_sizeof(len(mirror)) == 0


def test_argument_validation_without_gpu():
    lib = _lib.load()
    assert lib.hsg_gemm_nt(4, 4, 4, None, 4, None, 4, None, 4, None, None, 0, 0, None) == -1
    assert lib.hsg_layernorm_fwd(4, 6, 1, 1, 1, 1, 1, None) == -2          # D % 4 != 0
    csc = _lib.CscC(4, 4, 0, 0, 16, 16, 16, None)
    assert lib.hsg_edge_fwd(ctypes.byref(csc), 7, 8, 16, 72, 16, None, 16, None, 16, None) == -2   # (7,8) not instantiated
    fp, ldz = _lib.edge_layout(6, 50)
    assert (fp, ldz) == (300, 312) and _lib.edge_layout(8, 8) == (64, 72)
    perm = sorted(lib.hsg_edge_perm(6, 50, c) for c in range(300))
    assert perm == list(range(300))                     # a bijection onto [0, fp) for the default S2W shape
    assert lib.hsg_edge_bwd_workspace_bytes(8) > 0 and lib.hsg_gemm_tn_workspace_bytes(1000, 64, 64) > 0
    # sentence encoder / LSTM entry points (no device work on these argument errors)
    assert lib.hsg_enc_gather(4, 100, 300, 40, None, None, None, None, None, None, None) == -1      # null pointers
    assert lib.hsg_enc_gather(4, 5, 300, 40, 16, 16, 16, 16, 16, 16, None) == -1                     # L < 7
    assert lib.hsg_enc_gather(4, 100, 302, 40, 16, 16, 16, 16, 16, 16, None) == -2                   # D % 4 != 0
    assert lib.hsg_enc_conv_wgrad_workspace_bytes(1009, 300) > 0
    assert lib.hsg_lstm_fwd(2, 130, 2, 16, 16, 16, 16, 16, 16, 16, 16, 16, None) == -2               # H > 128
    assert lib.hsg_lstm_fwd(2, 128, 3, 16, 16, 16, 16, 16, 16, 16, 16, 16, None) == -1               # ndir not 1 / 2
    assert lib.hsg_lstm_bwd(0, 128, 2, None, None, None, None, None, None, None) == 0                # empty batch
    tok = np.zeros((2, 8), np.int32)
    out = np.zeros(16, np.int32)
    assert lib.hsg_enc_plan_host(2, 8, tok.ctypes.data, 1, np.asarray([0, 3], np.int32).ctypes.data, out.ctypes.data,
                                 out.ctypes.data + 16, out.ctypes.data + 32) == -1                   # ptr does not cover S


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_cuda():
    m = hb.WSWGAT(300, 64, 8, 0.0, 512, 0.0, 50, "W2S")
    batch = hb.HeteroBatch.from_csc_arrays(np.array([0, 1]), np.array([0]), np.array([3]), np.array([0]),
                                           np.array([0, 1]), np.array([0]), np.array([3]), device="cpu")
    batch.set_tfidf_embedding(torch.randn(10, 50))
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        m(batch, torch.randn(1, 300), torch.randn(1, 64))


def test_state_dict_matches_reference_keys_and_round_trips():
    z = dict(np.load(os.path.join(GOLD, "wswgat_hsg_default.npz")))
    ref_sd = {k[2:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("p:")}
    m = hb.WSWGATUpdateLoop()
    sd = m.state_dict()
    assert set(sd.keys()) == set(ref_sd.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref_sd[k].shape), k
    res = m.load_state_dict(ref_sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    # packed layout: head k occupies rows [k*d, (k+1)*d)
    assert torch.equal(m.word2sent.layer.fc_weight[8:16], ref_sd["word2sent.layer.heads.1.fc.weight"])
    assert torch.equal(m.sent2word.layer.attn_fc_weight[5:6], ref_sd["sent2word.layer.heads.5.attn_fc.weight"])
    assert sum(p.numel() for p in m.word2sent.parameters()) == 88832
    assert sum(p.numel() for p in m.sent2word.parameters()) == 344012
    bad = dict(ref_sd)
    bad.pop("word2sent.layer.heads.0.fc.weight")
    with pytest.raises(RuntimeError):
        hb.WSWGATUpdateLoop().load_state_dict(bad, strict=True)


def test_reference_error_behaviour():
    with pytest.raises(NotImplementedError, match="has not been implemented"):   # module/GAT.py:41
        hb.WSWGAT(64, 64, 8, 0.1, 512, 0.1, 50, "X2Y")
    m = hb.WSWGAT(300, 64, 8, 0.1, 512, 0.1, 50, "W2S").train()
    batch = hb.HeteroBatch.from_csc_arrays(np.array([0, 1]), np.array([0]), np.array([3]), np.array([0]),
                                           np.array([0, 1]), np.array([0]), np.array([3]), device="cpu")
    batch.set_tfidf_embedding(torch.randn(10, 50))
    # training-mode dropout is implemented in the sm_100a kernels: without a GPU the call fails loudly (no fallback)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        m(batch, torch.randn(1, 300), torch.randn(1, 64))
    # the bare sub-layers carry their training-mode dropout on their own too (GATStackLayer.py:56, GATLayer.py:41):
    # same kernels, so without a GPU they fail just as loudly
    with pytest.raises(RuntimeError, match="no CUDA device"):
        m.layer(batch, torch.randn(1, 300))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        m.ffn(torch.randn(1, 1, 64))
    # the never-instantiated S2S layer type keeps its input dropout unimplemented (HiGraph.py:57-76 builds no S2S)
    s2s = hb.MultiHeadSGATLayer(64, 8, 8, 0.1).train()
    with pytest.raises(NotImplementedError, match="dropout"):
        s2s(batch, torch.randn(1, 64))


def test_generator_shapes_and_determinism():
    a = syn.make_examples(4, "cnndm", seed=0)
    b = syn.make_examples(4, "cnndm", seed=0)
    for x, y in zip(a, b):
        assert np.array_equal(x.sents, y.sents) and x.w2s == y.w2s
        assert 3 <= x.n_sent <= 50 and x.sents.shape[1] == 100
        assert all(0.0 < v <= 1.0 for d in x.w2s for v in d.values())
    tb = syn.pack_token_batch(a)
    assert tb.tokens.dtype == np.int32 and tb.sent_bin.dtype == np.int8
    assert tb.graph_sent_ptr[-1] == tb.tokens.shape[0]
    lens = np.diff(tb.graph_sent_ptr)
    assert np.all(lens[:-1] >= lens[1:])                  # sorted by #sentences, descending (dataloader.py:479)
    assert tb.sent_bin.max() <= 9 and tb.sent_bin.min() >= -1
    bm = syn.filter_bitmap()
    assert (bm[0] & 1) == 1 and ((bm[0] >> 1) & 1) == 0   # PAD filtered, UNK kept (dataloader.py:171)
    h = syn.make_examples(3, "multinews", seed=2, hdsg=True)
    tbh = syn.pack_token_batch(h, hdsg=True)
    assert tbh.graph_doc_ptr[-1] == len(tbh.doc_tok_ptr) - 1
    assert len(tbh.sent_doc) == tbh.tokens.shape[0]


def test_stress_edges_and_csc_pair():
    word, sup, bins, extra = syn.stress_edges(4096, 512, 16384, seed=4)
    assert len(word) == 16384 and sup.max() < 512 and word.max() < 4096 and bins.max() <= 9
    (sip, ssrc, sbin, seid), (wip, wsrc, wbin, weid) = hb.csc_pair_from_edges(word, sup, bins, 4096, 512)
    assert sip[-1] == 16384 and wip[-1] == 16384
    assert np.array_equal(np.sort(seid), 2 * np.arange(16384))
    # twins: pair t = (word, sup) appears in both CSCs
    t = seid // 2
    assert np.array_equal(ssrc, word[t]) and np.array_equal(wsrc, sup[weid // 2])


def test_encoder_plan_host_logic():
    """EncoderPlan (host side of the sentence encoder): lengths, compact rows, positions and the PackedSequence order."""
    from hetersumgraph_b200.encoder import EncoderPlan
    L = 12
    tokens = np.zeros((7, L), np.int32)
    lens = [12, 0, 3, 6, 5, 1, 9]
    for s, n in enumerate(lens):
        tokens[s, :n] = 5
    tokens[2, 1] = 0                                   # interior zero: len 2, tail 3
    ptr = np.asarray([0, 3, 6, 7])
    plan = EncoderPlan(tokens, ptr, "cpu")
    assert plan.sent_len.tolist() == [12, 0, 2, 6, 5, 1, 9]
    assert np.diff(plan.row_ptr.numpy()).tolist() == [12, 7, 10, 12, 12, 8, 12]
    assert plan.sent_pos.tolist() == [1, 2, 3, 1, 2, 3, 1]
    # same order as torch's pack_padded_sequence over the per-graph lists
    seqs = [torch.arange(ptr[i], ptr[i + 1]).float()[:, None] for i in range(3)]
    packed = torch.nn.utils.rnn.pack_padded_sequence(torch.nn.utils.rnn.pad_sequence(seqs, batch_first=True), [3, 3, 1],
                                                     batch_first=True)
    assert packed.data[:, 0].long().tolist() == plan.perm.tolist()
    assert packed.batch_sizes.tolist() == plan.batch_sizes.tolist()
    assert plan.perm[plan.inv_perm].tolist() == list(range(7))
    with pytest.raises(ValueError):
        EncoderPlan(tokens, np.asarray([0, 1, 7]), "cpu").perm   # a PackedSequence needs batch order (cuDNN option only)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) runs without a GPU and prints ONE JSON line
    with the contract's keys: impl, the metric / unit / config of our arm, cpu_baseline of this run, e2e = the line's own
    value with zero H2D / D2H bytes."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--graphs-per-gpu", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "graphs/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("HSG train graphs/sec") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert set(d["config"]) == {"workload", "graphs_per_gpu", "n_iter", "dropout"} and d["config"]["graphs_per_gpu"] == 2
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
