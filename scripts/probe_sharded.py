"""dev tool: per-phase timing of the sharded (multi-GPU) step.  torchrun --nproc-per-node N scripts/probe_sharded.py"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
import bench as B
from daisyrec_b200 import ops
from daisyrec_b200.parallel import ShardedTrainer, partition_users, allreduce_step_buffers
from daisyrec_b200.utils.synthetic import init_tables

local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
class A: pass
a = A(); a.shape = "ml-20m"; a.num_ng = 4; a.seed = 2022; a.factors = 64; a.batch = 1 << 20
d, triples = B.build_workload(a.shape, dev, a.num_ng, a.seed, "cuda")
U, I, F = d["user_num"], d["item_num"], a.factors
T = triples.shape[0]; Bg = a.batch * world
bounds = partition_users((d["row_ptr"][1:] - d["row_ptr"][:-1]).cpu().numpy(), world)
g = torch.Generator(device=dev); g.manual_seed(a.seed)
perm = torch.randperm(T, generator=g, device=dev)
P0, Q0 = init_tables(U, I, F, a.seed, dev)
lo, hi = int(bounds[rank]), int(bounds[rank + 1])
tr = ShardedTrainer(P0[lo:hi].contiguous(), Q0.contiguous(), bounds, rank, world, ops.hyper(0.01, 0.001, 0.001))
spe = tr.prepare_epoch(triples, perm, Bg)
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = np.zeros(3)
for s in range(60):
    b, e = int(tr.offsets_host[s % spe]), int(tr.offsets_host[s % spe + 1])
    e0, e1, e2, e3 = ev(), ev(), ev(), ev()
    e0.record(); tr._phase(1, tr.bu, tr.bi, tr.bj, b, e - b); e1.record()
    allreduce_step_buffers(tr.gq, tr.cnt_i, tr.acc, tr.group); e2.record()
    tr._phase(2, tr.bu, tr.bi, tr.bj, b, e - b); e3.record()
    tr.opt_steps += 1
    torch.cuda.synchronize()
    if s >= 10:
        acc += [e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)]
if rank == 0:
    print(json.dumps(dict(world=world, phase1_ms=acc[0] / 50, allreduce_ms=acc[1] / 50, phase2_ms=acc[2] / 50)))
dist.destroy_process_group()
