"""CPU dry run of GPU tests: the test bodies of tests/test_gpu_nfm.py and tests/test_gpu_zzz_late.py run with every CUDA op
replaced by an oracle-backed stand-in and every 'cuda' tensor kept on the CPU.

What this checks: the test harnesses (fixture indexing, mask layouts, tolerances, RNG-state assertions) and the host logic of the
NFM / NGCF classes (mask draw order, chunking, train_step / calc_loss / rank flows) -- everything except the kernels themselves.
A GPU failure of one of these tests then points at a kernel, not at the test.  TEST INFRASTRUCTURE ONLY (it drives oracle/);
run in its own process (it patches torch), e.g. from tests/test_dryrun_cpu.py:   python tests/dryrun_on_oracle.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc  # noqa: E402
from daisyrec_b200 import ops  # noqa: E402
import daisyrec_b200.model.AbstractRecommender as AR  # noqa: E402

orc.build()

# ------------------------------------------------------------------ keep every tensor on the CPU
ops.require_cuda = lambda: None
AR.ops.require_cuda = ops.require_cuda
torch.cuda.current_device = lambda: 0
torch.cuda.device_count = lambda: 1
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a, **k: None
_to = torch.Tensor.to


def _is_cuda_dev(x):
    return (isinstance(x, torch.device) and x.type == "cuda") or (isinstance(x, str) and x.startswith("cuda"))


def _to_cpu(self, *a, **k):
    a = tuple(x for x in a if not _is_cuda_dev(x))
    k = {kk: v for kk, v in k.items() if not (kk == "device" and _is_cuda_dev(v))}
    return _to(self, *a, **k) if (a or k) else self


torch.Tensor.to = _to_cpu
torch.Tensor.cuda = lambda self, *a, **k: self


def _strip(fn):
    def g(*a, **k):
        if "device" in k and _is_cuda_dev(k["device"]):
            k.pop("device")
        return fn(*a, **k)
    return g


torch.zeros, torch.empty, torch.full, torch.arange, torch.tensor = map(_strip, (torch.zeros, torch.empty, torch.full, torch.arange,
                                                                               torch.tensor))
ops.mf_step_variant = lambda *a, **k: (False, 16, 1)
ops.hyper = lambda lr, r1, r2, opt="sgd": orc.hyper(lr=lr, reg_1=r1, reg_2=r2, opt=opt)
ops.check_index_range = lambda *a, **k: None


def _gather_triples(d_triples, d_perm=None):
    t = d_triples if d_perm is None else d_triples[d_perm]
    return t[:, 0].contiguous(), t[:, 1].contiguous(), t[:, 2].contiguous()


ops.gather_triples = _gather_triples


def _topk_from_scores(scores, cands, topk):
    sc = scores.numpy()
    idx = np.argsort(-sc, axis=1, kind="stable")[:, :topk]
    c = cands.numpy() if cands is not None else np.tile(np.arange(sc.shape[1]), (sc.shape[0], 1))
    return torch.from_numpy(np.take_along_axis(c, idx, 1).astype(np.float32 if cands is not None else np.int64))


ops.topk_from_scores = _topk_from_scores
ops.mf_rank = lambda P, Q, users, cands, topk: torch.from_numpy(
    orc.mf_rank(np.ascontiguousarray(P.numpy()), np.ascontiguousarray(Q.numpy()), users.numpy(), cands.numpy(), topk))
ops.mf_full_rank = lambda P, Q, users, topk: torch.from_numpy(
    orc.mf_full_rank(np.ascontiguousarray(P.numpy()), np.ascontiguousarray(Q.numpy()), users.numpy(), topk))
ops.mf_predict = lambda P, Q, u, i: torch.from_numpy(
    orc.mf_predict(np.ascontiguousarray(P.numpy()), np.ascontiguousarray(Q.numpy()), u.numpy(), i.numpy()))


# ------------------------------------------------------------------ NFM stand-ins (oracle/nfm_oracle.c)
class _NfmWs:
    def __init__(self, U, I, F, L, bn, opt, rows, dev):
        self.U, self.I, self.F, self.Ln, self.bn, self.opt, self.max_rows, self.state = U, I, F, L, 1 if bn else 0, opt, rows, None


def _nfm_steps(P, Q, bias, N, R, ws, act, bu, bi, bj, batch, first, n_steps, hp, adam_step0=0, apply=True, check=True, tower_dtype=0,
               dropout=0.0, keep=None):
    F, L, out = ws.F, ws.Ln, []
    Rn = R.numpy() if R is not None and R.numel() else np.zeros(0, np.float32)
    if ws.opt == "adam" and ws.state is None:
        ws.state = np.zeros(2 * (P.numel() + Q.numel() + bias.numel() + N.numel()), np.float32)
    for s in range(n_steps):
        lo = (first + s) * batch
        hi = min(lo + batch, bu.numel())
        B, kk = hi - lo, None
        if keep is not None:                       # documented layout: per step [forward call][site][rows][F] bytes, stride `batch` rows
            per = 2 * (1 + L) * batch * F
            k = keep.numpy()[s * per:s * per + 2 * (1 + L) * B * F].reshape(2, 1 + L, B, F).astype(np.float32)
            k = k * (np.float32(1) / np.float32(1 - dropout))
            kk = (np.ascontiguousarray(k[0]), np.ascontiguousarray(k[1]))
        out.append(orc.nfm_bpr_step(P.numpy(), Q.numpy(), bias.numpy(), N.numpy(), Rn, L, bool(ws.bn), act,
                                    *(np.ascontiguousarray(t.numpy()[lo:hi]) for t in (bu, bi, bj)), hp, apply, ws.state,
                                    adam_step0 + s + 1, keep=kk))
    return torch.tensor(out, dtype=torch.float64)


def _nfm_scores(P, Q, bias, N, R, ws, act, u, i, tower_dtype=0):
    Rn = R.numpy() if R is not None and R.numel() else np.zeros(0, np.float32)
    return torch.from_numpy(orc.nfm_scores(P.numpy(), Q.numpy(), bias.numpy(), N.numpy(), Rn, ws.Ln, bool(ws.bn), act, u.numpy(), i.numpy()))


ops.NfmWorkspace, ops.nfm_bpr_train_steps, ops.nfm_scores = _NfmWs, _nfm_steps, _nfm_scores


# ------------------------------------------------------------------ NGCF stand-ins (oracle/ngcf_oracle.c)
class _Graph:
    def __init__(self, row_ptr, col, val, dev):
        self.a = (row_ptr, col, val)


class _NgcfWs:
    def __init__(self, U, I, dims, opt, dev):
        self.U, self.I, self.dims, self.opt, self.state = U, I, [int(d) for d in dims], opt, None


def _factors(keep, p):
    return np.ascontiguousarray(keep.numpy().astype(np.float32) * (np.float32(1) / np.float32(1 - p)))


def _ngcf_forward(E0, W, ws, graph, tower_dtype=0, dropout=0.0, keep=None):
    return torch.from_numpy(orc.ngcf_forward(E0.numpy(), W.numpy(), ws.U, ws.I, np.asarray(ws.dims, np.int32), *graph.a,
                                             keep=None if keep is None else _factors(keep, dropout)))


def _ngcf_steps(E0, W, ws, graph, bu, bi, bj, batch, first, n_steps, hp, adam_step0=0, apply=True, check=True, tower_dtype=0,
                dropout=0.0, keep=None):
    if ws.opt == "adam" and ws.state is None:
        ws.state = np.zeros(2 * (E0.numel() + W.numel()), np.float32)
    per, out = ops.ngcf_keep_bytes(ws), []
    for s in range(n_steps):
        lo = (first + s) * batch
        hi = min(lo + batch, bu.numel())
        out.append(orc.ngcf_bpr_step(E0.numpy(), W.numpy(), ws.U, ws.I, np.asarray(ws.dims, np.int32), *graph.a,
                                     *(np.ascontiguousarray(t.numpy()[lo:hi]) for t in (bu, bi, bj)), hp, apply, ws.state,
                                     adam_step0 + s + 1, keep=None if keep is None else _factors(keep[s * per:(s + 1) * per], dropout)))
    return torch.tensor(out, dtype=torch.float64)


ops.LgcnGraph, ops.NgcfWorkspace, ops.ngcf_forward, ops.ngcf_bpr_train_steps = _Graph, _NgcfWs, _ngcf_forward, _ngcf_steps


def main():
    import test_gpu_nfm as TN
    import test_gpu_ngcf as TG
    import test_gpu_zzz_late as TL
    to_cpu = lambda a: torch.from_numpy(np.ascontiguousarray(a).copy())  # noqa: E731
    for mod in (TN, TG, TL):
        mod.dev = to_cpu
    runs = [("test_gpu_nfm steps", lambda: TN.test_nfm_steps_match_reference_fixture(orc)),
            ("test_gpu_nfm class", TN.test_nfm_class_drop_in),
            ("test_gpu_ngcf steps", lambda: TG.test_ngcf_forward_and_steps_match_reference_fixture(orc)),
            ("test_gpu_ngcf class", TG.test_ngcf_class_drop_in),
            ("late nfm dropout steps", lambda: TL.test_nfm_dropout_steps_match_reference_fixture(orc)),
            ("late nfm dropout class", TL.test_nfm_class_runs_the_reference_default_config),
            ("late ngcf dropout steps", lambda: TL.test_ngcf_message_dropout_matches_reference_fixture(orc)),
            ("late ngcf dropout class", TL.test_ngcf_class_runs_the_reference_default_config)]
    for name, fn in runs:
        fn()
        print(f"DRYRUN OK: {name}", flush=True)


if __name__ == "__main__":
    main()
