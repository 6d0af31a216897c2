"""Developer A/B of the MF step kernel: ms/step at the config-2 (ML-20M, F=64) and config-5 (Netflix, F=128) shapes.
usage: [DRB_NO_LEAN=1] python scripts/ab_step.py [c2] [c5] [adam]"""
import sys

import torch

sys.path.insert(0, ".")
from daisyrec_b200 import ops  # noqa: E402
from daisyrec_b200.utils.synthetic import SHAPES, init_tables  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(3)


def run(shape, F, opt, B=1 << 20, steps=40):
    U, I, _ = SHAPES[shape]
    P, Q = init_tables(U, I, F, 1, dev)
    n = B * steps
    bu = torch.randint(0, U, (n,), device=dev, dtype=torch.int32, generator=g)
    bi = (I * torch.rand(n, device=dev, generator=g).pow(2.0)).to(torch.int32).clamp_(0, I - 1)
    bj = torch.randint(0, I, (n,), device=dev, dtype=torch.int32, generator=g)
    ws = ops.MFWorkspace(U, I, F, opt, dev)
    hp = ops.hyper(0.01, 0.001, 0.001, opt)
    ops.mf_bpr_train_steps(P, Q, ws, bu, bi, bj, B, 0, steps, hp, check=False)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        losses = ops.mf_bpr_train_steps(P, Q, ws, bu, bi, bj, B, 0, steps, hp, check=False)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    print(f"{shape} F={F} {opt}: {best:.4f} ms/step  {B / best / 1e6:.3f} G triples/s  loss[-1]={float(losses[-1]):.6f}", flush=True)


what = sys.argv[1:] or ["c2", "c5"]
opt = "adam" if "adam" in what else "sgd"
if "c2" in what:
    run("ml-20m", 64, opt)
if "c5" in what:
    run("netflix", 128, opt)
