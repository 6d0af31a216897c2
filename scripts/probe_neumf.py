"""Developer probe: a few NeuMF steps at BASELINE config 3's shape (ML-20M user/item counts, F=32, L=2, Adam, B=1 M) for ncu /
timing.  usage: python scripts/probe_neumf.py [fused|bf16|fp32] [steps]"""
import sys

import torch

sys.path.insert(0, ".")
from daisyrec_b200 import ops  # noqa: E402

td = {"fp32": 0, "bf16": 1, "fused": 2}[sys.argv[1] if len(sys.argv) > 1 else "fused"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda")
U, I, F, L, B = 138493, 26744, 32, 2, 1 << 20
D = F * 2 ** (L - 1)
g = torch.Generator(device=dev); g.manual_seed(1)
tabs = [(torch.randn(s, device=dev, generator=g) * 0.05).contiguous() for s in ((U, F), (I, F), (U, D), (I, D))]
W = (torch.randn(ops.neumf_param_count(F, L), device=dev, generator=g) * 0.1).contiguous()
bu = torch.randint(0, U, (B * 2,), device=dev, dtype=torch.int32, generator=g)
bi = (I * torch.rand(B * 2, device=dev, generator=g).pow(2.0)).to(torch.int32).clamp_(0, I - 1)
bj = torch.randint(0, I, (B * 2,), device=dev, dtype=torch.int32, generator=g)
ws = ops.NeumfWorkspace(U, I, F, L, "adam", 2 * B, dev)
hp = ops.hyper(0.001, 0.001, 0.001, "adam")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for s in range(steps + 2):
    if s == 2:
        torch.cuda.synchronize(); e0.record()
    ops.neumf_bpr_train_steps(tabs, W, ws, bu, bi, bj, B, s % 2, 1, hp, adam_step0=s, check=False, tower_dtype=td)
e1.record(); torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1) / steps, "triples/s", B * steps / e0.elapsed_time(e1) * 1e3)
