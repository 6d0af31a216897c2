"""Secondary measurements of the rows either side of the hot path (not the driver's bench): KPIs, device CSR /
adjacency builders, the point-wise (CL) step and the Adagrad / RMSprop sweeps, at BASELINE config-2/4 shapes.

    python scripts/bench_eval.py [out.json]
One JSON object; every timing is CUDA events around `reps` launches after warm-up, inputs resident in HBM.
"""
import json
import sys
import traceback

import torch

sys.path.insert(0, ".")
from daisyrec_b200 import ops  # noqa: E402


def timed(fn, warm=2, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def section(out, name, fn):
    try:
        out[name] = fn()
    except Exception as e:  # noqa: BLE001  (keep the other sections)
        out[name] = {"error": repr(e), "trace": traceback.format_exc()[-600:]}


def main():
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    out = {}
    U, I, F = 138493, 26744, 64                       # ml-20m shape (BASELINE config 2)

    def kpis():
        n, K = U, 50
        lens = torch.randint(1, 21, (n,), device=dev, generator=g)
        ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(lens, 0)
        nnz = int(ptr[-1].item())
        # ascending ids inside a row: sorted random offsets
        rows = torch.repeat_interleave(torch.arange(n, device=dev), lens)
        idx = torch.randint(0, I, (nnz,), device=dev, generator=g)
        key, _ = torch.sort(rows * I + idx)
        idx = (key % I).to(torch.int32).contiguous()
        preds = torch.randint(0, I, (n, K), device=dev, generator=g).to(torch.float32).contiguous()
        ks = [1, 5, 10, 20, 30, 50]
        pop = torch.rand(I, dtype=torch.float64, device=dev)
        ms = timed(lambda: ops.rank_metrics(preds, ptr, idx, ks, I), reps=20)
        ms_pop = timed(lambda: ops.rank_metrics(preds, ptr, idx, ks, I, pop), reps=20)
        return {"users": n, "topk": K, "cutoffs": ks, "ms": ms, "ms_with_popularity": ms_pop,
                "users_per_s": n / ms * 1e3, "bytes_read": n * K * 4 + nnz * 4 + (n + 1) * 8}

    def csr():
        nnz = 20_000_000
        cu = torch.randint(0, U, (nnz,), device=dev, dtype=torch.int32, generator=g)
        ci = torch.randint(0, I, (nnz,), device=dev, dtype=torch.int32, generator=g)
        ms = timed(lambda: ops.csr_build(cu, ci, U, I), warm=1, reps=3)
        ptr, col = ops.csr_build(cu, ci, U, I)
        return {"rows": U, "cols": I, "nnz_in": nnz, "nnz_unique": int(col.numel()), "ms": ms, "pairs_per_s": nnz / ms * 1e3}

    def adj():
        U4, I4, nnz = 52643, 91599, 3_000_000         # amazon-book shape (BASELINE config 4)
        cu = torch.randint(0, U4, (nnz,), device=dev, dtype=torch.int32, generator=g)
        ci = torch.randint(0, I4, (nnz,), device=dev, dtype=torch.int32, generator=g)
        ms = timed(lambda: ops.lgcn_build_adj(cu, ci, U4, I4), warm=1, reps=3)
        return {"users": U4, "items": I4, "nnz_in": nnz, "ms": ms}

    def steps(loss, opt, B=1 << 20, k=20):
        P = (torch.randn(U, F, device=dev) * 0.01).contiguous()
        Q = (torch.randn(I, F, device=dev) * 0.01).contiguous()
        ws = ops.MFWorkspace(U, I, F, opt, dev)
        hp = ops.hyper(0.01, 0.001, 0.001, opt, loss=loss)
        n = B * k
        bu = torch.randint(0, U, (n,), device=dev, dtype=torch.int32, generator=g)
        bi = torch.randint(0, I, (n,), device=dev, dtype=torch.int32, generator=g)
        third = torch.randint(0, 2 if loss == "CL" else I, (n,), device=dev, dtype=torch.int32, generator=g)
        ms = timed(lambda: ops.mf_bpr_train_steps(P, Q, ws, bu, bi, third, B, 0, k, hp, check=False), warm=1, reps=3) / k
        return {"loss": loss, "opt": opt, "batch": B, "ms_per_step": ms, "rows_per_s": B / ms * 1e3}

    section(out, "kpis_ml20m", kpis)
    section(out, "csr_build_ml20m", csr)
    section(out, "lgcn_build_adj_amazon_book", adj)
    section(out, "mf_step_bpr_sgd", lambda: steps("BPR", "sgd"))
    section(out, "mf_step_cl_sgd", lambda: steps("CL", "sgd"))
    section(out, "mf_step_bpr_adagrad", lambda: steps("BPR", "adagrad"))
    section(out, "mf_step_bpr_rmsprop", lambda: steps("BPR", "rmsprop"))
    section(out, "mf_step_bpr_adam", lambda: steps("BPR", "adam"))
    txt = json.dumps(out)
    print(txt)
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
