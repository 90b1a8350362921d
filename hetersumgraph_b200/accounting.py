"""Algorithmic bytes / flops of every kernel of the path (the roofline numerators).

Compulsory traffic only (every operand read once, every result written once, gathers that hit
L2 not counted twice), fp32 storage - SURVEY.md §8-d, restated in DESIGN.md §"Roofline accounting".
Keys are the kernel slot names of hsg_profile_slot_name().
"""
from collections import defaultdict


def round_up(x, m):
    return (x + m - 1) // m * m


def edge_fwd_bytes(E, n_src, n_dst, H, d, with_x=True):
    F = H * d
    b = E * (4 + 1) + n_dst * (4 + 4)          # nbr + bin per edge, indptr + extra per destination
    b += n_src * (F + H) * 4                   # z | p rows, each source row once
    b += n_dst * F * 4                         # sh
    if with_x:
        b += 2 * n_dst * F * 4                 # origin read + x written
    b += n_dst * H * 8                         # m, den
    return b


def edge_bwd_prep_bytes(n_dst, H, d):
    F = H * d
    return 3 * n_dst * F * 4 + n_dst * H * 4   # dx, sh read; g written; s written


def edge_bwd_prep_rc_bytes(E, n_src, n_dst, H, d):
    """recomputing prep (csrc/hsg_edge_rc.cu): dx read, g written, the forward CSC, (m, den) read, s written, the source
    rows [z | p] once."""
    F = H * d
    return 2 * n_dst * F * 4 + E * 5 + n_dst * 4 + n_dst * H * 12 + n_src * (F + H) * 4


def edge_bwd_bytes(E, n_src, n_dst, H, d):
    """source-centric pass: rows = forward sources (n_src), gathered rows = forward destinations (n_dst)."""
    F = H * d
    ldz = round_up(F + H, 8)
    b = E * (4 + 1) + n_src * 4                # nbr + bin per edge, indptr per row
    b += n_dst * (F + 3 * H) * 4               # g rows + (m, den, s), each destination once
    b += n_src * (F + H) * 4                   # z | p of the row itself
    b += n_src * ldz * 4                       # dzp written
    return b


def edge_fwd_bytes_survey(E, n_src, n_dst, H, d):
    """SURVEY.md 8(d) B_fwd: E(4+1) + N_dst(4+4) + N_src(F+H)4 + N_dst F 4 (origin) + N_dst F 4 (out) + N_dst H 8.
    (edge_fwd_bytes above additionally counts the saved `sh` write, which the kernel really performs.)"""
    F = H * d
    return E * 5 + n_dst * 8 + n_src * (F + H) * 4 + 2 * n_dst * F * 4 + n_dst * H * 8


def edge_bwd_bytes_survey(E, n_src, n_dst, H, d):
    """SURVEY.md 8(d) B_bwd of the WHOLE edge backward (prep + source-centric pass) of one application whose forward
    had n_src sources and n_dst destinations: 2E(4+1) + E 4 (eid map) + (N_src+N_dst) 4 + N_dst F 4 (g) + N_dst H 8 +
    N_src(F+H)4 (z,p) + N_src(F+H)4 (dz,dp) + 10 H 4 (dq)."""
    F = H * d
    return (2 * E * 5 + E * 4 + (n_src + n_dst) * 4 + n_dst * F * 4 + n_dst * H * 8 + 2 * n_src * (F + H) * 4 +
            10 * H * 4)


def builder_bytes(n_sent, sent_len, n_pair, n_word, n_super):
    """K0: S L 5 token + bin bytes read by each of the two passes, ~26 B per pair + the node maps written."""
    return 2 * n_sent * sent_len * 5 + n_pair * 26 + n_word * 12 + n_super * 21


def wswgat_application(E, n_src, n_dst, H, d, in_dim, d_hid):
    """{slot: [flops, bytes, launches]} of ONE WSWGAT application, forward + backward."""
    F = H * d
    ldz = round_up(F + H, 8)
    acc = defaultdict(lambda: [0, 0, 0])

    def add(slot, flops=0, nbytes=0, n=1):
        acc[slot][0] += flops
        acc[slot][1] += nbytes
        acc[slot][2] += n

    def gemm(slot, M, N, K):
        add(slot, 2 * M * N * K, (M * K + N * K + M * N) * 4)

    add("attn_prep_fwd", 0, (ldz * in_dim + F * in_dim) * 4)
    add("attn_prep_bwd", 0, (ldz * in_dim + 2 * F * in_dim) * 4)
    # forward
    gemm("gemm_nt", n_src, ldz, in_dim)        # zp = h W_aug^T
    add("edge_fwd", 0, edge_fwd_bytes(E, n_src, n_dst, H, d))
    gemm("gemm_nt", n_dst, d_hid, F)           # hdn = relu(x W1^T + b1)
    gemm("gemm_nt", n_dst, F, d_hid)           # r = hdn W2^T + b2 + x
    add("layernorm_fwd", 0, 2 * n_dst * F * 4)
    # backward
    add("layernorm_bwd", 0, 3 * n_dst * F * 4)
    add("layernorm_bwd_reduce", 0, 2 * F * 4 * 64)          # per-block partials of dgamma / dbeta (order of 64 blocks)
    gemm("gemm_nn", n_dst, d_hid, F)           # dhp = (dr W2) * relu'
    gemm("gemm_tn", n_dst, F, d_hid)           # dW2
    gemm("gemm_tn", n_dst, d_hid, F)           # dW1
    gemm("gemm_nn", n_dst, F, d_hid)           # dx = dhp W1 + dr
    # fixed-order second stage of the three weight-gradient products: partials read once, result written once
    add("gemm_tn_reduce", 0, (F * d_hid * 2 + ldz * in_dim) * 4 * 9, 3)
    add("edge_bwd_prep", 0, edge_bwd_prep_bytes(n_dst, H, d))
    add("edge_bwd", 0, edge_bwd_bytes(E, n_src, n_dst, H, d))
    add("edge_bwd_dq", 0, 10 * H * 4 * 296)
    gemm("gemm_nn", n_src, in_dim, ldz)        # dh = dzp W_aug
    gemm("gemm_tn", n_src, ldz, in_dim)        # dW_aug
    return acc


def step_accounting(n_word, n_super, n_pair, n_iter=1, emb=300, hid=64, n_head=8, d_hid=512):
    """{slot: (flops, bytes, launches)} of one fwd+bwd pass of the update loop (W2S, n_iter x (S2W, W2S))."""
    total = defaultdict(lambda: [0, 0, 0])
    w2s = wswgat_application(n_pair, n_word, n_super, n_head, hid // n_head, emb, d_hid)
    s2w = wswgat_application(n_pair, n_super, n_word, 6, emb // 6, hid, d_hid)
    for app, times in ((w2s, 1 + n_iter), (s2w, n_iter)):
        for k, (f, b, n) in app.items():
            total[k][0] += f * times
            total[k][1] += b * times
            total[k][2] += n * times
    return {k: tuple(v) for k, v in total.items()}
