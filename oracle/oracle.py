"""ctypes binding of oracle/bpr_oracle.c (the CPU checker) + small numpy helpers.

TEST INFRASTRUCTURE ONLY -- see the header of bpr_oracle.c.  Parity is pinned by
tests/golden/ fixtures generated from the real reference (oracle/gen_golden.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbpr_oracle.so")


def build(force=False):
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("bpr_oracle.c", "eval_oracle.c", "ngcf_oracle.c", "nfm_oracle.c", "Makefile"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class Hyper(C.Structure):
    _fields_ = [("lr", C.c_float), ("reg_1", C.c_float), ("reg_2", C.c_float), ("opt", C.c_int32),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("loss", C.c_int32)]


OPT_KIND = {"sgd": 0, "adam": 1, "adagrad": 2, "rmsprop": 3}
LOSS_KIND = {"BPR": 0, "HL": 1, "TL": 2, "CL": 3, "SL": 4}


def hyper(lr=0.01, reg_1=0.001, reg_2=0.001, opt="sgd", beta1=0.9, beta2=0.999, eps=1e-8, loss="BPR"):
    return Hyper(lr, reg_1, reg_2, OPT_KIND[opt], beta1, beta2, eps, LOSS_KIND[loss.upper()])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_mt_next.restype = C.c_uint32
        _lib.orc_dot.restype = C.c_float
        _lib.orc_mf_bpr_step.restype = C.c_double
        _lib.orc_mf_bpr_epoch.restype = C.c_double
        _lib.orc_fm_step.restype = C.c_double
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return _p(a, C.c_float)


def _i32(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return _p(a, C.c_int32)


def _i64(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return _p(a, C.c_int64)


# ---------------------------------------------------------------- RNG
def mt_seed(seed):
    """np.random.seed(int) -> uint32[625] state (key + pos)."""
    st = np.zeros(625, np.uint32)
    lib().orc_mt_seed(_p(st, C.c_uint32), C.c_uint32(seed & 0xFFFFFFFF))
    return st


def mt_from_numpy(rs=None):
    """Snapshot numpy's (global or given) legacy RandomState into the oracle layout."""
    s = (np.random if rs is None else rs).get_state()
    st = np.zeros(625, np.uint32)
    st[:624] = s[1]
    st[624] = s[2]
    return st


def mt_to_numpy(st, rs=None):
    (np.random if rs is None else rs).set_state(("MT19937", st[:624].copy(), int(st[624]), 0, 0.0))


def mt_next(st):
    return int(lib().orc_mt_next(_p(st, C.c_uint32)))


# ---------------------------------------------------------------- sampler
def csr_from_ur(ur, user_num):
    """dict[int -> set[int]] (config['train_ur']) -> (row_ptr int64[U+1], col int32 sorted)."""
    deg = np.zeros(user_num + 1, np.int64)
    for u in range(user_num):
        deg[u + 1] = len(ur[u]) if u in ur else 0
    row_ptr = np.cumsum(deg)
    col = np.empty(int(row_ptr[-1]), np.int32)
    for u in range(user_num):
        if u in ur and len(ur[u]):
            col[row_ptr[u]:row_ptr[u + 1]] = sorted(ur[u])
    return row_ptr, col


def sample_negatives(mt_state, row_ptr, col, user_num, item_num, num_ng):
    js = np.zeros((user_num, num_ng), np.int32)
    rc = lib().orc_sample_negatives(_p(mt_state, C.c_uint32), _i64(row_ptr), _i32(col), user_num, item_num, num_ng,
                                    _i32(js))
    if rc != 0:
        raise ValueError("a cannot be empty (user %d has interacted with every item)" % (-rc - 1))
    return js


def explode_triples(coo_u, coo_i, js):
    n, g = len(coo_u), js.shape[1]
    out = np.empty((n * g, 3), np.int32)
    lib().orc_explode_triples(_i32(coo_u), _i32(coo_i), C.c_int64(n), _i32(js), g, _i32(out))
    return out


def build_candidates_user(mt_state, gt, train, item_num, cand_num):
    gt = np.ascontiguousarray(gt, np.int32)
    train = np.ascontiguousarray(train, np.int32)
    out = np.empty(cand_num, np.int64)
    rc = lib().orc_build_candidates_user(_p(mt_state, C.c_uint32), _i32(gt), len(gt), _i32(train), len(train),
                                         item_num, cand_num, _i64(out))
    if rc != 0:
        raise ValueError("a cannot be empty")
    return out


KPI_NAMES = ("recall", "mrr", "ndcg", "hit", "precision", "map", "coverage", "popularity")


def gt_csr(test_ur, test_u):
    """test_ur (dict user -> set) in test_u order -> (gt_ptr int64[n+1], gt_idx int32 ascending per row)."""
    lens = np.array([len(test_ur[u]) for u in test_u], np.int64)
    ptr = np.zeros(len(test_u) + 1, np.int64)
    np.cumsum(lens, out=ptr[1:])
    idx = np.concatenate([np.sort(np.fromiter(test_ur[u], np.int32, len(test_ur[u]))) for u in test_u]) \
        if len(test_u) else np.zeros(0, np.int32)
    return ptr, np.ascontiguousarray(idx, np.int32)


def rank_metrics(preds, gt_ptr, gt_idx, ks, item_num, item_pop=None):
    """-> float64 [len(ks), 8] in KPI_NAMES order (daisy/utils/metrics.py)."""
    preds = np.ascontiguousarray(preds, np.float32)
    ks = np.ascontiguousarray(ks, np.int32)
    out = np.zeros((len(ks), len(KPI_NAMES)), np.float64)
    pop = None if item_pop is None else np.ascontiguousarray(item_pop, np.float64)
    lib().orc_rank_metrics(_f32(preds), C.c_int64(preds.shape[0]), C.c_int32(preds.shape[1]), _i64(gt_ptr), _i32(gt_idx),
                           _i32(ks), C.c_int32(len(ks)), C.c_int32(item_num), _p(pop, C.c_double), _p(out, C.c_double))
    return out


def sample_negatives_pop(mt_state, row_ptr, col, user_num, item_num, uniform_num, other_num, cdf):
    js = np.zeros((user_num, uniform_num + other_num), np.int32)
    cdf = np.ascontiguousarray(cdf, np.float64)
    rc = lib().orc_sample_negatives_pop(_p(mt_state, C.c_uint32), _i64(row_ptr), _i32(col), user_num, item_num,
                                        uniform_num, other_num, _p(cdf, C.c_double), _i32(js))
    if rc != 0:
        raise ValueError("a cannot be empty (user %d has interacted with every item)" % (-rc - 1))
    return js


def explode_pointwise(coo_u, coo_i, label, js):
    n, g = len(coo_u), js.shape[1]
    out = np.empty((n * (1 + g), 3), np.int32)
    lib().orc_explode_pointwise(_i32(coo_u), _i32(coo_i), _i32(label), C.c_int64(n), _i32(js), g, _i32(out))
    return out


# ---------------------------------------------------------------- model
def dot(a, b):
    return float(lib().orc_dot(_f32(a), _f32(b), len(a)))


def mf_predict(P, Q, u, i):
    out = np.empty(len(u), np.float32)
    lib().orc_mf_predict(_f32(P), _f32(Q), P.shape[1], _i32(u), _i32(i), C.c_int64(len(u)), _f32(out))
    return out


def mf_bpr_step(P, Q, bu, bi, bj, hp, apply=True, adam_state=None, step_count=1):
    """In-place step on P,Q (float32 [U,F],[I,F]).  Returns (loss, parts[8])."""
    parts = np.zeros(8, np.float64)
    mP = vP = mQ = vQ = None
    if adam_state is not None:
        mP, vP, mQ, vQ = adam_state
    loss = lib().orc_mf_bpr_step(_f32(P), _f32(Q), P.shape[0], Q.shape[0], P.shape[1], _i32(bu), _i32(bi), _i32(bj),
                                 C.c_int64(len(bu)), C.byref(hp), int(apply),
                                 None if mP is None else _f32(mP), None if vP is None else _f32(vP),
                                 None if mQ is None else _f32(mQ), None if vQ is None else _f32(vQ),
                                 C.c_int64(step_count), _p(parts, C.c_double))
    return loss, parts


def fm_step(P, Q, bias, bu, bi, bj, hp, apply=True, state=None, bias_state=None, step_count=1):
    """daisy/model/FMRecommender.py:61-97.  bias = [u_bias (U), i_bias (I), bias_ (1)] float32, updated in place."""
    parts = np.zeros(8, np.float64)
    m = [None] * 4 if state is None else list(state)
    loss = lib().orc_fm_step(_f32(P), _f32(Q), _f32(bias), P.shape[0], Q.shape[0], P.shape[1], _i32(bu), _i32(bi), _i32(bj),
                             C.c_int64(len(bu)), C.byref(hp), 1 if apply else 0, *[None if a is None else _f32(a) for a in m],
                             None if bias_state is None else _f32(bias_state), C.c_int64(step_count), _p(parts, C.c_double))
    return loss, parts


def fm_scores(P, Q, bias, user, items=None):
    n = Q.shape[0] if items is None else len(items)
    out = np.empty(n, np.float32)
    it = None if items is None else np.ascontiguousarray(items, np.int64)
    lib().orc_fm_scores(_f32(P), _f32(Q), _f32(bias), P.shape[0], Q.shape[0], P.shape[1], C.c_int64(int(user)),
                        None if it is None else _i64(it), C.c_int64(n), _f32(out))
    return out


def mf_bpr_epoch(P, Q, triples, perm, batch, hp, adam_state=None, first_step_count=1):
    T = len(triples)
    nsteps = (T + batch - 1) // batch
    step_loss = np.zeros(nsteps, np.float64)
    mP = vP = mQ = vQ = None
    if adam_state is not None:
        mP, vP, mQ, vQ = adam_state
    total = lib().orc_mf_bpr_epoch(_f32(P), _f32(Q), P.shape[0], Q.shape[0], P.shape[1], _i32(triples), C.c_int64(T),
                                   None if perm is None else _i64(perm), C.c_int64(batch), C.byref(hp),
                                   None if mP is None else _f32(mP), None if vP is None else _f32(vP),
                                   None if mQ is None else _f32(mQ), None if vQ is None else _f32(vQ),
                                   C.c_int64(first_step_count), _p(step_loss, C.c_double))
    return total, step_loss


def mf_rank(P, Q, users, cands, topk):
    users = np.ascontiguousarray(users, np.int64)
    cands = np.ascontiguousarray(cands, np.int64)
    out = np.empty((len(users), topk), np.float32)
    lib().orc_mf_rank(_f32(P), _f32(Q), P.shape[1], _i64(users), C.c_int64(len(users)), _i64(cands), cands.shape[1],
                      topk, _f32(out))
    return out


def mf_full_rank(P, Q, users, topk):
    users = np.ascontiguousarray(users, np.int64)
    out = np.empty((len(users), topk), np.int64)
    lib().orc_mf_full_rank(_f32(P), _f32(Q), P.shape[1], Q.shape[0], _i64(users), C.c_int64(len(users)), topk,
                           _i64(out))
    return out


# ---------------------------------------------------------------- LightGCN
def lgcn_norm_adj(coo_u, coo_i, user_num, item_num):
    """daisy/model/LightGCNRecommender.py:73-107 (get_norm_adj_mat) as CSR over the U+I nodes.

    A = binary bipartite adjacency (duplicate interactions collapse, as the dok dict does, :87-89);
    diag = (A>0).sum(1) + 1e-7 in float64, ** -0.5 (:92-96); L = D A D in float64 ((D*A)*D, :98), cast to
    float32 (:104).  Returns (row_ptr int64[n+1], col int32 ascending per row, val float32)."""
    n = user_num + item_num
    u = np.asarray(coo_u, np.int64)
    i = np.asarray(coo_i, np.int64) + user_num
    key = np.unique(np.concatenate([u * n + i, i * n + u]))
    row, col = key // n, key % n
    deg = np.bincount(row, minlength=n).astype(np.float64)
    dinv = np.power(deg + 1e-7, -0.5)
    val = ((dinv[row] * 1.0) * dinv[col]).astype(np.float32)
    row_ptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(row, minlength=n), out=row_ptr[1:])
    return row_ptr, col.astype(np.int32), val


def lgcn_propagate(row_ptr, col, val, E0, L):
    n, F = E0.shape
    out = np.empty_like(E0)
    lib().orc_lgcn_propagate(_i64(row_ptr), _i32(col), _f32(val), C.c_int64(n), F, L, _f32(E0), _f32(out))
    return out


def lgcn_bpr_step(E0, U, I, L, row_ptr, col, val, bu, bi, bj, hp, apply=True, adam_state=None, step_count=1):
    """In-place step on E0 = cat(P, Q) float32 [(U+I), F].  Returns the loss."""
    m = v = None
    if adam_state is not None:
        m, v = adam_state
    lib().orc_lgcn_bpr_step.restype = C.c_double
    return lib().orc_lgcn_bpr_step(_f32(E0), U, I, E0.shape[1], L, _i64(row_ptr), _i32(col), _f32(val), _i32(bu), _i32(bi),
                                   _i32(bj), C.c_int64(len(bu)), C.byref(hp), int(apply),
                                   None if m is None else _f32(m), None if v is None else _f32(v), C.c_int64(step_count))


# ---------------------------------------------------------------- NeuMF
# ---------------------------------------------------------------- NGCF (ngcf_oracle.c)
def ngcf_param_count(dims):
    dims = np.ascontiguousarray(dims, np.int32)
    lib().orc_ngcf_param_count.restype = C.c_int64
    return int(lib().orc_ngcf_param_count(_i32(dims), len(dims) - 1))


def ngcf_dropout_keep(n, dims, p):
    """The factors of NGCF's nn.Dropout(mess_dropout) (NGCFRecommender.py:164) for ONE forward(): one bernoulli_ per layer on
    torch's global CPU generator over the [n, dims[l + 1]] layer output, scaled by 1 / (1 - p) -> float32, layers concatenated."""
    import torch
    scale = np.float32(1.0) / np.float32(1.0 - p)
    return np.ascontiguousarray(np.concatenate([torch.empty(n, int(d), dtype=torch.float32).bernoulli_(1.0 - p).numpy().reshape(-1)
                                                for d in list(dims)[1:]]) * scale, np.float32)


def ngcf_forward(E0, W, U, I, dims, row_ptr, col, val, keep=None):
    """NGCF.forward: [U + I, sum(dims)] concatenated layer outputs.  keep: ngcf_dropout_keep(...) when mess_dropout > 0."""
    dims = np.ascontiguousarray(dims, np.int32)
    out = np.empty((U + I, int(dims.sum())), np.float32)
    lib().orc_ngcf_forward_ex(_f32(E0), _f32(W), U, I, _i32(dims), len(dims) - 1, _i64(row_ptr), _i32(col), _f32(val), _f32(out),
                              None if keep is None else _f32(keep))
    return out


def ngcf_bpr_step(E0, W, U, I, dims, row_ptr, col, val, bu, bi, bj, hp, apply=True, state=None, step_count=1, keep=None):
    dims = np.ascontiguousarray(dims, np.int32)
    lib().orc_ngcf_bpr_step_ex.restype = C.c_double
    return lib().orc_ngcf_bpr_step_ex(_f32(E0), _f32(W), U, I, _i32(dims), len(dims) - 1, _i64(row_ptr), _i32(col), _f32(val),
                                      _i32(bu), _i32(bi), _i32(bj), C.c_int64(len(bu)), C.byref(hp), 1 if apply else 0,
                                      None if state is None else _f32(state), C.c_int64(step_count),
                                      None if keep is None else _f32(keep))


# ---------------------------------------------------------------- NFM (nfm_oracle.c)
def nfm_param_count(F, L, bn):
    lib().orc_nfm_param_count.restype = C.c_int64
    return int(lib().orc_nfm_param_count(F, L, 1 if bn else 0))


def nfm_dropout_keep(B, F, L, p):
    """The factors NFM's Dropout modules multiply with during ONE training step of B triples (NFMRecommender.py:67,:88):
    forward(user, pos) draws site 0 (FM_layers) and sites 1..L (behind each activation), then forward(user, neg) the same; one
    bernoulli_ per Dropout call on torch's global CPU generator, scaled by 1 / (1 - p) -> (keep_pos, keep_neg), float32
    [1 + L, B, F] each."""
    import torch
    scale = np.float32(1.0) / np.float32(1.0 - p)
    sides = []
    for _side in range(2):
        sides.append(np.stack([torch.empty(B, F, dtype=torch.float32).bernoulli_(1.0 - p).numpy() * scale for _ in range(1 + L)]))
    return np.ascontiguousarray(sides[0], np.float32), np.ascontiguousarray(sides[1], np.float32)


def nfm_bpr_step(P, Q, bias, N, R, L, bn, act, bu, bi, bj, hp, apply=True, state=None, step_count=1, keep=None):
    """In place on P, Q, bias, N (net block) and R (BatchNorm running statistics).  Returns the loss.
    keep: nfm_dropout_keep(...) in train mode with dropout > 0."""
    lib().orc_nfm_bpr_step_ex.restype = C.c_double
    kp, kn = (None, None) if keep is None else (_f32(keep[0]), _f32(keep[1]))
    return lib().orc_nfm_bpr_step_ex(_f32(P), _f32(Q), _f32(bias), _f32(N), _f32(R) if R.size else None, P.shape[0], Q.shape[0],
                                     P.shape[1], L, 1 if bn else 0, act, _i32(bu), _i32(bi), _i32(bj), C.c_int64(len(bu)),
                                     C.byref(hp), 1 if apply else 0, None if state is None else _f32(state),
                                     C.c_int64(step_count), kp, kn)


def nfm_scores(P, Q, bias, N, R, L, bn, act, users, items):
    users = np.ascontiguousarray(users, np.int32)
    items = np.ascontiguousarray(items, np.int32)
    out = np.empty(len(users), np.float32)
    lib().orc_nfm_scores(_f32(P), _f32(Q), _f32(bias), _f32(N), _f32(R) if R.size else None, P.shape[0], Q.shape[0], P.shape[1],
                         L, 1 if bn else 0, act, _i32(users), _i32(items), C.c_int64(len(users)), _f32(out))
    return out


NEUMF_MODE = {"NeuMF": 0, "NeuMF-pre": 0, "GMF": 1, "MLP": 2}


def neumf_param_count(F, L, mode=0):
    lib().orc_neumf_param_count_ex.restype = C.c_int64
    return int(lib().orc_neumf_param_count_ex(F, L, mode))


def torch_dropout_keep(B, F, L, p):
    """The keep factors nn.Dropout draws for ONE training step of B triples (NeuMFRecommender.py:61, forward(user, pos) then
    forward(user, neg), one bernoulli_ per Dropout call on torch's global CPU generator) -> float32 [2B, sum n_l]."""
    import torch
    widths = [F * (2 ** (L - i)) for i in range(L)]
    sides = []
    for _side in range(2):
        sides.append(np.concatenate([torch.empty(B, n, dtype=torch.float32).bernoulli_(1.0 - p).numpy() for n in widths], 1))
    return np.ascontiguousarray(np.concatenate(sides, 0) / np.float32(1.0 - p), np.float32)


def neumf_predict(tabs, W, F, L, u, it):
    UG, IG, UM, IM = tabs
    out = np.empty(len(u), np.float32)
    lib().orc_neumf_predict(_f32(UG), _f32(IG), _f32(UM), _f32(IM), _f32(W), F, L, _i32(u), _i32(it), C.c_int64(len(u)),
                            _f32(out))
    return out


def neumf_bpr_step(tabs, W, F, L, bu, bi, bj, hp, apply=True, adam_state=None, step_count=1, mode=0, keep=None):
    """In-place step on the 4 tables (UG, IG, UM, IM) and the flat tower block W.  adam_state = (m[5], v[5]).
    mode: NEUMF_MODE; keep: torch_dropout_keep(...) in train mode with dropout > 0."""
    UG, IG, UM, IM = tabs
    PF = C.POINTER(C.c_float)
    m = v = None
    if adam_state is not None:
        m = (PF * 5)(*[_f32(a) for a in adam_state[0]])
        v = (PF * 5)(*[_f32(a) for a in adam_state[1]])
    lib().orc_neumf_bpr_step_ex.restype = C.c_double
    return lib().orc_neumf_bpr_step_ex(_f32(UG), _f32(IG), _f32(UM), _f32(IM), _f32(W), UG.shape[0], IG.shape[0], F, L,
                                       _i32(bu), _i32(bi), _i32(bj), C.c_int64(len(bu)), C.byref(hp), int(apply), m, v,
                                       C.c_int64(step_count), mode, None if keep is None else _f32(keep))


def neumf_rank(tabs, W, F, L, users, cands, topk, mode=0):
    UG, IG, UM, IM = tabs
    users = np.ascontiguousarray(users, np.int64)
    cands = np.ascontiguousarray(cands, np.int64)
    out = np.empty((len(users), topk), np.float32)
    lib().orc_neumf_rank_ex(_f32(UG), _f32(IG), _f32(UM), _f32(IM), _f32(W), F, L, _i64(users), C.c_int64(len(users)),
                            _i64(cands), cands.shape[1], IG.shape[0], topk, _f32(out), None, mode)
    return out


def neumf_full_rank(tabs, W, F, L, users, topk, mode=0):
    UG, IG, UM, IM = tabs
    users = np.ascontiguousarray(users, np.int64)
    out = np.empty((len(users), topk), np.int64)
    lib().orc_neumf_rank_ex(_f32(UG), _f32(IG), _f32(UM), _f32(IM), _f32(W), F, L, _i64(users), C.c_int64(len(users)),
                            None, 0, IG.shape[0], topk, None, _i64(out), mode)
    return out
