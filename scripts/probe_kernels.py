"""Developer probe for ncu: one launch of every kernel either side of the step (rank, full_rank, sampler, KPIs, permutation)
plus the MF step kernel at the config-2 and config-5 shapes.  usage: python scripts/probe_kernels.py [c2|c5|all]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from daisyrec_b200 import ops  # noqa: E402
from daisyrec_b200.utils.synthetic import SHAPES, init_tables, make_interactions  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(3)


def step_kernel(shape, F, B=1 << 20, steps=8):
    U, I, _ = SHAPES[shape]
    P, Q = init_tables(U, I, F, 1, dev)
    bu = torch.randint(0, U, (B * steps,), device=dev, dtype=torch.int32, generator=g)
    bi = (I * torch.rand(B * steps, device=dev, generator=g).pow(2.0)).to(torch.int32).clamp_(0, I - 1)
    bj = torch.randint(0, I, (B * steps,), device=dev, dtype=torch.int32, generator=g)
    ws = ops.MFWorkspace(U, I, F, "sgd", dev)
    hp = ops.hyper(0.01, 0.001, 0.001)
    ops.mf_bpr_train_steps(P, Q, ws, bu, bi, bj, B, 0, steps, hp, check=False)
    torch.cuda.synchronize()
    ops.mf_bpr_train_steps(P, Q, ws, bu, bi, bj, B, 0, steps, hp, check=False)
    torch.cuda.synchronize()
    return P, Q


if what in ("c2", "all"):
    P, Q = step_kernel("ml-20m", 64)
    U, I, nnz = SHAPES["ml-20m"]
    users = torch.randint(0, U, (4096,), device=dev, generator=g)
    cands = torch.randint(0, I, (4096, 1000), device=dev, generator=g)
    for _ in range(2):
        preds = ops.mf_rank(P, Q, users, cands, 50)
        ops.mf_full_rank(P, Q, users[:512], 50)
    lens = torch.randint(1, 21, (4096,), device=dev, generator=g)
    ptr = torch.zeros(4097, dtype=torch.int64, device=dev); ptr[1:] = torch.cumsum(lens, 0)
    rows = torch.repeat_interleave(torch.arange(4096, device=dev), lens)
    idx = torch.randint(0, I, (int(ptr[-1].item()),), device=dev, generator=g)
    key, _ = torch.sort(rows * I + idx)
    idx = (key % I).to(torch.int32).contiguous()
    for _ in range(2):
        ops.rank_metrics(preds, ptr, idx, [1, 5, 10, 20, 30, 50], I)
    d = make_interactions(U, I, nnz, seed=2022, device=dev)
    draws = ops.sampler_draw_mt19937(ops.mt19937_seed(1), d["row_ptr"].cpu().numpy(), U, I, 4)
    for _ in range(2):
        js = ops.sampler_kth_complement(d["row_ptr"], d["col"], torch.from_numpy(draws).to(dev), I)
        tr = ops.sampler_explode(d["coo_u"], d["coo_i"], js)
    for _ in range(2):
        ops.randperm_torch(7, tr.shape[0], dev)
    torch.cuda.synchronize()
if what in ("c5", "all"):
    step_kernel("netflix", 128)
print("done")
