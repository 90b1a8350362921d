"""ctypes binding of libhsg_b200.so (the C ABI declared in include/hsg_b200.h).

The product path has NO fallback: if the library is missing it is built in-tree
with nvcc (hetersumgraph_b200/build.py); if that fails, or the device is not a
B200, every op raises.
"""
import ctypes as C
import os

from . import build as _build

_LIB = None

c_i32p = C.c_void_p   # device pointers are passed as integers (tensor.data_ptr())
c_f32p = C.c_void_p


class TokenBatchC(C.Structure):
    _fields_ = [("n_graphs", C.c_int32), ("n_sent", C.c_int32), ("sent_len", C.c_int32), ("hdsg", C.c_int32),
                ("vocab_size", C.c_int32), ("n_doc", C.c_int32), ("n_doc_tok", C.c_int32),
                ("max_sent_per_graph", C.c_int32),
                ("tokens", C.c_void_p), ("sent_bin", C.c_void_p), ("graph_sent_ptr", C.c_void_p),
                ("filter_bitmap", C.c_void_p), ("graph_doc_ptr", C.c_void_p), ("sent_doc", C.c_void_p),
                ("doc_tok_ptr", C.c_void_p), ("doc_tokens", C.c_void_p), ("doc_bin", C.c_void_p)]


class GraphOffsetsC(C.Structure):
    _fields_ = [("word_ptr", C.c_void_p), ("super_ptr", C.c_void_p), ("node_ptr", C.c_void_p),
                ("edge_ptr", C.c_void_p), ("pair_ptr", C.c_void_p)]


class CscC(C.Structure):
    _fields_ = [("n_dst", C.c_int32), ("n_src", C.c_int32), ("n_edges", C.c_int32), ("reserved", C.c_int32),
                ("indptr", C.c_void_p), ("nbr", C.c_void_p), ("bin", C.c_void_p), ("extra", C.c_void_p)]


class GraphOutC(C.Structure):
    _fields_ = [("cap_word", C.c_int32), ("cap_super", C.c_int32), ("cap_pair", C.c_int32), ("reserved", C.c_int32),
                ("off", GraphOffsetsC),
                ("word_wid", C.c_void_p), ("word_nid", C.c_void_p), ("super_nid", C.c_void_p),
                ("super_type", C.c_void_p), ("super_graph", C.c_void_p), ("super_indptr", C.c_void_p),
                ("super_src", C.c_void_p), ("super_bin", C.c_void_p), ("super_eid", C.c_void_p),
                ("super_extra", C.c_void_p), ("word_indptr", C.c_void_p), ("word_src", C.c_void_p),
                ("word_bin", C.c_void_p), ("word_eid", C.c_void_p), ("status", C.c_void_p)]


class WswgatFwdArgsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("H", "d", "in_dim", "d_hid", "n_src", "n_dst", "ldz", "reserved")] + \
               [("csc", C.POINTER(CscC))] + \
               [(n, C.c_void_p) for n in ("neighbor", "origin", "W_aug", "q", "w1", "b1", "w2", "b2", "gamma", "beta",
                                          "zp", "sh", "x", "stat", "hdn", "r", "ln_stats", "out")]


class WswgatBwdArgsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("H", "d", "in_dim", "d_hid", "n_src", "n_dst", "ldz", "reserved")] + \
               [("csc_t", C.POINTER(CscC))] + \
               [(n, C.c_void_p) for n in ("dout", "neighbor", "W_aug", "q", "w1", "w2", "gamma", "zp", "sh", "x", "hdn",
                                          "r", "ln_stats", "stat", "dr", "dhp", "g", "dzp", "dx", "d_neighbor",
                                          "dW_aug", "dq", "dw1", "db1", "dw2", "db2", "dgamma", "dbeta", "ws")] + \
               [("ws_bytes", C.c_size_t)]


class LayerParamsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("H", "d", "in_dim", "feat_dim", "d_hid", "reserved")] + \
               [(n, C.c_void_p) for n in ("W", "Wf", "bf", "a", "w1", "b1", "w2", "b2", "gamma", "beta")]


class LayerGradsC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dW", "dWf", "dbf", "da", "dw1", "db1", "dw2", "db2", "dgamma", "dbeta")]


class LoopArgsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_apps", "start_kind", "n_word", "n_super")] + \
               [("csc_super", C.POINTER(CscC)), ("csc_word", C.POINTER(CscC)),
                ("w2s", LayerParamsC), ("s2w", LayerParamsC),
                ("T", C.c_void_p), ("word_feature", C.c_void_p), ("super_feature", C.c_void_p),
                ("state", C.c_void_p), ("state_floats", C.c_size_t),
                ("attn_p", C.c_float), ("ffn_p", C.c_float), ("seed", C.c_ulonglong), ("seed_dev", C.c_void_p),
                ("input_ready", C.c_void_p)]


class LoopPlanC(C.Structure):
    _fields_ = [("state_floats", C.c_size_t), ("scratch_floats", C.c_size_t), ("ws_bytes", C.c_size_t),
                ("word_state_off", C.c_size_t), ("super_state_off", C.c_size_t), ("hdn_off", C.c_size_t * 2),
                ("pair_stride", C.c_size_t)]


class LoopBwdArgsC(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_word_state", "d_super_state", "d_word_feature", "d_super_feature")] + \
               [("w2s", LayerGradsC), ("s2w", LayerGradsC), ("dT", C.c_void_p),
                ("accumulate", C.c_int32), ("reserved", C.c_int32),
                ("scratch", C.c_void_p), ("scratch_floats", C.c_size_t), ("ws", C.c_void_p), ("ws_bytes", C.c_size_t)]


class S2SGraphC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_graphs", "n_super", "H", "d", "mult", "reserved")] + \
               [(n, C.c_void_p) for n in ("super_ptr", "deg_indptr", "extra", "xgrp", "xmember")]


class DocMapC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_sent", "n_doc", "hidden", "reserved")] + \
               [(n, C.c_void_p) for n in ("sent_row", "doc_row", "sent_doc", "doc_graph", "graph_sent_ptr")]


class HeadArgsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_sent", "n_super", "hidden", "two_part", "n_graphs", "reserved")] + \
               [(n, C.c_void_p) for n in ("state", "sent_row", "doc_row", "graph_sent_ptr", "wh_w", "wh_b", "labels")] + \
               [("inv_graphs", C.c_float), ("reserved2", C.c_float)]


_I, _P, _Z = C.c_int, C.c_void_p, C.c_size_t

_PROTOS = {
    "hsg_version": (C.c_int, []),
    "hsg_strerror": (C.c_char_p, [_I]),
    "hsg_device_check": (C.c_int, []),
    "hsg_num_sms": (C.c_int, []),
    "hsg_set_pdl": (C.c_int, [_I]),
    "hsg_profile_enable": (C.c_int, [_I]),
    "hsg_profile_reset": (C.c_int, []),
    "hsg_profile_num_slots": (C.c_int, []),
    "hsg_profile_slot_name": (C.c_char_p, [_I]),
    "hsg_profile_read": (C.c_int, [_I, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "hsg_launch_count": (C.c_longlong, []),
    "hsg_memset": (C.c_int, [_P, _I, _Z, _P]),
    "hsg_build_workspace_bytes": (_Z, [C.POINTER(TokenBatchC)]),
    "hsg_build_count": (C.c_int, [C.POINTER(TokenBatchC), GraphOffsetsC, _P, _P, _Z, _P]),
    "hsg_build_fill": (C.c_int, [C.POINTER(TokenBatchC), C.POINTER(GraphOutC), _P, _Z, _P]),
    "hsg_attn_prep_fwd": (C.c_int, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hsg_attn_prep_bwd": (C.c_int, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hsg_set_gemm_mode": (C.c_int, [_I]),
    "hsg_get_gemm_mode": (C.c_int, []),
    "hsg_set_gemm_small_flops": (C.c_int, [C.c_double]),
    "hsg_gemm_trace": (C.c_int, [_I, C.POINTER(C.c_ulonglong), _I]),
    "hsg_gemm_pair_trace": (C.c_int, [_I, C.POINTER(C.c_ulonglong), _I]),
    "hsg_gemm_nt": (C.c_int, [_I, _I, _I, _P, _I, _P, _I, _P, _I, _P, _P, _I, _I, _P]),
    "hsg_gemm_nn": (C.c_int, [_I, _I, _I, _P, _I, _P, _I, _P, _I, _P, _I, _I, _P]),
    "hsg_gemm_tn_workspace_bytes": (_Z, [_I, _I, _I]),
    "hsg_gemm_tn": (C.c_int, [_I, _I, _I, _P, _I, _P, _I, _P, _I, _P, _P, _Z, _P]),
    "hsg_gemm_tn_acc": (C.c_int, [_I, _I, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, _Z, _P]),
    "hsg_edge_layout": (C.c_int, [_I, _I, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "hsg_edge_perm": (C.c_int, [_I, _I, _I]),
    "hsg_edge_fwd": (C.c_int, [C.POINTER(CscC), _I, _I, _P, _I, _P, _P, _P, _P, _P, _P]),
    "hsg_edge_bwd_prep": (C.c_int, [_I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "hsg_edge_bwd_workspace_bytes": (_Z, [_I]),
    "hsg_set_edge_rowpar": (C.c_int, [_I]),
    "hsg_set_edge_fwd_lowdeg": (C.c_int, [_I]),
    "hsg_set_edge_recompute": (C.c_int, [_I]),
    "hsg_edge_bwd_prep_rc_ok": (C.c_int, [_I, _I, _I]),
    "hsg_edge_bwd_prep_rc": (C.c_int, [C.POINTER(CscC), _I, _I, _P, _I, _P, _P, _P, _P, _P]),
    "hsg_set_edge_fwd_rowpar": (C.c_int, [_I]),
    "hsg_set_edge_blockrow": (C.c_int, [_I]),
    "hsg_set_edge_bwd_async": (C.c_int, [_I]),
    "hsg_enc_plan_host": (C.c_int, [_I, _I, _P, _I, _P, _P, _P, _P]),
    "hsg_enc_gather": (C.c_int, [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "hsg_enc_pack_weights": (C.c_int, [_I, _P, _P, _P]),
    "hsg_enc_pool_fwd": (C.c_int, [_I, _P, _P, _I, _P, _P, _I, _P, _P]),
    "hsg_enc_conv_wgrad_workspace_bytes": (_Z, [_I, _I]),
    "hsg_enc_conv_wgrad": (C.c_int, [_I, _I, _P, _P, _I, _P, _P, _P, _I, _P, _Z, _P]),
    "hsg_add_rows": (C.c_int, [_I, _I, _P, _I, _P, _P, _P, _I, _P]),
    "hsg_lstm_fwd": (C.c_int, [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "hsg_lstm_bwd": (C.c_int, [_I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "hsg_edge_bwd": (C.c_int, [C.POINTER(CscC), _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "hsg_wswgat_fwd": (C.c_int, [C.POINTER(WswgatFwdArgsC), _P]),
    "hsg_wswgat_bwd_workspace_bytes": (_Z, [_I, _I, _I, _I, _I, _I]),
    "hsg_wswgat_bwd": (C.c_int, [C.POINTER(WswgatBwdArgsC), _P]),
    "hsg_set_bwd_overlap": (C.c_int, [_I]),
    "hsg_set_side_ctas": (C.c_int, [_I]),
    "hsg_set_tn_min_rows": (C.c_int, [_I]),
    "hsg_set_tn_item_rows": (C.c_int, [_I]),
    "hsg_set_gemm_pair": (C.c_int, [_I]),
    "hsg_update_loop_plan": (C.c_int, [C.POINTER(LoopArgsC), C.POINTER(LoopPlanC)]),
    "hsg_update_loop_fwd": (C.c_int, [C.POINTER(LoopArgsC), _P]),
    "hsg_update_loop_bwd": (C.c_int, [C.POINTER(LoopArgsC), C.POINTER(LoopBwdArgsC), _P]),
    "hsg_head_workspace_bytes": (_Z, [_I, _I]),
    "hsg_abi_sizeof": (_Z, [_I]),
    "hsg_head_fwd": (C.c_int, [C.POINTER(HeadArgsC), _P, _P, _P, _P, _Z, _P]),
    "hsg_head_bwd": (C.c_int, [C.POINTER(HeadArgsC), _P, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "hsg_head_fwd_bwd": (C.c_int, [C.POINTER(HeadArgsC), _P, _P, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "hsg_doc_mean": (C.c_int, [C.POINTER(DocMapC), _P, _P, _P]),
    "hsg_super_assemble": (C.c_int, [C.POINTER(DocMapC), _P, _P, _P, _P]),
    "hsg_doc_init_bwd": (C.c_int, [C.POINTER(DocMapC), _P, _P, _P, _P, _P]),
    "hsg_topm": (C.c_int, [_P, _P, _I, _I, _P, _P]),
    "hsg_adam_workspace_bytes": (_Z, []),
    "hsg_adam_step": (C.c_int, [_Z, _P, _P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_float, _I, C.c_float, _P, _Z,
                                _P]),
    "hsg_adam_step_dev": (C.c_int, [_Z, _P, _P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_float, _P, _I, C.c_float, _P,
                                    _Z, _P]),
    "hsg_allreduce_adam_buffer_floats": (_Z, [_Z, _I]),
    "hsg_allreduce_adam_step": (C.c_int, [_Z, _P, _P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_float, _P, _P, _I, _I,
                                          _P]),
    "hsg_embed_gather": (C.c_int, [_I, _I, _P, _P, _P, _P]),
    "hsg_s2s_fwd": (C.c_int, [C.POINTER(S2SGraphC), _P, _P, _P, _P, _P, _P, _P]),
    "hsg_s2s_bwd_workspace_bytes": (_Z, [_I, _I, _I]),
    "hsg_s2s_bwd": (C.c_int, [C.POINTER(S2SGraphC), _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "hsg_dropout_mask": (C.c_int, [_Z, C.c_float, C.c_ulonglong, C.c_uint, _P, _P]),
    "hsg_layernorm_fwd": (C.c_int, [_I, _I, _P, _P, _P, _P, _P, _P]),
    "hsg_layernorm_bwd_workspace_bytes": (_Z, [_I, _I]),
    "hsg_ffn_rows_ok": (C.c_int, [_I, _I, _I]),
    "hsg_ffn_rows_fwd": (C.c_int, [_I, _I, _I] + [_P] * 12),
    "hsg_ffn_rows_bwd": (C.c_int, [_I, _I, _I] + [_P] * 12 + [_I, _P, _Z, _P]),
    "hsg_layernorm_bwd": (C.c_int, [_I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS.keys())

EPI_BIAS, EPI_RELU, EPI_ADD, EPI_RELU_MASK = 1, 2, 4, 8


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load (building first if needed) the shared library; raises if impossible."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if build_if_missing and _build.needs_build():
        _build.build()
    if not os.path.exists(_build.LIB):
        raise RuntimeError("libhsg_b200.so is missing and could not be built; the WSWGAT path has no fallback")
    lib = C.CDLL(_build.LIB)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().hsg_strerror(int(rc)).decode()
        raise RuntimeError("hsg_b200: %s (status %d)" % (msg, rc))


_DEVICE_OK = False


def require_device():
    """Fail loudly unless a sm_100 device is current."""
    global _DEVICE_OK
    if _DEVICE_OK:
        return
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("hsg_b200: no CUDA device - the WSWGAT path runs only on B200 (sm_100a), no CPU fallback")
    check(load().hsg_device_check())
    _DEVICE_OK = True


GEMM_MODES = {"fp32": 0, "tf32x3": 1, "tf32": 2, "bf16": 3}


def set_gemm_mode(mode):
    """'fp32' (FFMA), 'tf32x3' (tcgen05, fp32-parity, default), 'tf32' (tcgen05, single pass) or 'bf16' (tcgen05
    kind::f16 on bf16-rounded operands, fp32 accumulation: the 2e-2 tolerance class)."""
    check(load().hsg_set_gemm_mode(GEMM_MODES[mode] if isinstance(mode, str) else int(mode)))


def get_gemm_mode():
    inv = {v: k for k, v in GEMM_MODES.items()}
    return inv[load().hsg_get_gemm_mode()]


_LAYOUTS = {}


def edge_layout(H, d):
    """(fp, ldz) of the lane-interleaved gathered-row layout for (heads, head_dim)."""
    key = (H, d)
    if key not in _LAYOUTS:
        fp, ldz = C.c_int(0), C.c_int(0)
        check(load().hsg_edge_layout(H, d, C.byref(fp), C.byref(ldz)))
        _LAYOUTS[key] = (fp.value, ldz.value)
    return _LAYOUTS[key]


def profile_snapshot():
    """{kernel name: (launches, total ms)} since the last reset."""
    lib = load()
    out = {}
    cnt, ms = C.c_int(0), C.c_float(0.0)
    for i in range(lib.hsg_profile_num_slots()):
        check(lib.hsg_profile_read(i, C.byref(cnt), C.byref(ms)))
        if cnt.value:
            out[lib.hsg_profile_slot_name(i).decode()] = (cnt.value, float(ms.value))
    return out
