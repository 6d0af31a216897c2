"""CPU suite for the N>1 host logic (world_size 2, gloo): user partitioning, the per-step exchange and the
rank-time all-gather.  The CUDA kernels themselves are covered by tests/test_gpu_multi.py on >=2 GPUs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from daisyrec_b200.parallel import (partition_users, owner_of, allgather_rows, allreduce_step_buffers, broadcast_cpu_,
                                    broadcast_int)


def test_partition_users_balances_weight():
    rng = np.random.default_rng(0)
    w = rng.zipf(1.5, size=5000).clip(max=2000)
    for world in (1, 2, 3, 8):
        b = partition_users(w, world)
        assert b[0] == 0 and b[-1] == len(w) and np.all(np.diff(b) >= 0)
        loads = [w[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(loads) - min(loads) <= 2 * w.max()              # contiguous split: off by at most one user each side
        own = owner_of(np.arange(len(w)), b)
        for r in range(world):
            assert np.array_equal(np.flatnonzero(own == r), np.arange(b[r], b[r + 1]))
    assert np.array_equal(partition_users(np.zeros(10, np.int64), 2), [0, 0, 10])   # degenerate weights


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # per-step exchange: fp32 gradient accumulator, packed u64 counters (as int64), fp64 scalars
        gq = torch.full((6, 4), float(rank + 1))
        cnt = torch.tensor([(3 << 32) | 5, 7], dtype=torch.int64) * (rank + 1)
        acc = torch.arange(8, dtype=torch.float64) + rank
        allreduce_step_buffers(gq, cnt, acc)
        ok = bool((gq == 3.0).all()) and cnt.tolist() == [((3 << 32) | 5) * 3, 21] and \
            torch.equal(acc, torch.arange(8, dtype=torch.float64) * 2 + 1)
        # rank-time gather: rank 0 owns rows {0,3,4}, rank 1 owns {1,2}; K=3
        mine = [torch.tensor([0, 3, 4]), torch.tensor([1, 2])][rank]
        rows = (mine[:, None] * 10 + torch.arange(3)[None, :]).to(torch.float32)
        full = allgather_rows(rows, mine, 5)
        want = (torch.arange(5)[:, None] * 10 + torch.arange(3)[None, :]).to(torch.float32)
        ok = ok and torch.equal(full, want)
        # a rank with no rows still participates
        empty = allgather_rows(rows[:0] if rank == 1 else rows, mine[:0] if rank == 1 else mine, 5)
        ok = ok and torch.equal(empty[[0, 3, 4]], want[[0, 3, 4]])
        # RNG-dependent state comes from rank 0 whatever the ranks drew themselves (model init, epoch seed)
        torch.manual_seed(100 + rank)                                      # deliberately different RNG histories
        w = torch.empty(7, 3).normal_()
        mine_before = w.clone()
        broadcast_cpu_(w, torch.device("cpu"))
        g = torch.Generator(); g.manual_seed(100)
        ok = ok and torch.equal(w, torch.empty(7, 3).normal_(generator=g)) and (rank == 0 or not torch.equal(w, mine_before))
        from daisyrec_b200.model.AbstractRecommender import epoch_seed, epoch_permutation
        seed = broadcast_int(epoch_seed(True), torch.device("cpu"))
        perm = epoch_permutation(1000, True, seed=seed)
        both = [torch.empty_like(perm) for _ in range(world)]
        dist.all_gather(both, perm)
        ok = ok and torch.equal(both[0], both[1]) and sorted(perm.tolist()) == list(range(1000))
        # the host-only group the peer-buffer mapping takes turns on (turn-by-turn barriers, no device work)
        from daisyrec_b200.parallel import _host_group
        hg = _host_group(None)
        order = []
        for turn in range(world):
            if turn == rank:
                order.append(turn)
            dist.barrier(group=hg)
        ok = ok and order == [rank] and _host_group(None) is hg and dist.get_backend(hg) == "gloo"
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_exchange_and_gather_world2_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
