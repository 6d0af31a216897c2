"""Quick throughput probe of the step kernel at the ML-20M shape (dev tool, not the bench)."""
import sys, json
import torch
sys.path.insert(0, ".")
from daisyrec_b200 import ops

U, I, F = 138493, 26744, int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
P = (torch.randn(U, F, device="cuda") * 0.01)
Q = (torch.randn(I, F, device="cuda") * 0.01)
g = torch.Generator(device="cuda"); g.manual_seed(1)
for B in [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else '256,8192,65536,1048576,4194304'.split(','))]:
    K = max(4, min(200, (1 << 25) // B))
    n = B * K
    bu = torch.randint(0, U, (n,), device="cuda", dtype=torch.int32, generator=g)
    # zipf-ish item popularity
    r = torch.rand(n, device="cuda", generator=g)
    bi = (I * r.pow(3.0)).to(torch.int32).clamp_(0, I - 1)
    bj = torch.randint(0, I, (n,), device="cuda", dtype=torch.int32, generator=g)
    for reg in (0.001,):
        hp = ops.hyper(0.01, reg, reg)
        ws = ops.MFWorkspace(U, I, F, "sgd", "cuda")
        ops.mf_bpr_train_steps(P, Q, ws, bu, bi, bj, B, 0, min(K, 3), hp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.mf_bpr_train_steps(P, Q, ws, bu, bi, bj, B, 0, K, hp, check=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        tps = n / ms * 1e3
        print(json.dumps(dict(F=F, B=B, steps=K, reg=reg, ms_per_step=ms / K, triples_per_s=tps,
                              alg_GBps=tps * (24 * F + 12) / 1e9)), flush=True)
