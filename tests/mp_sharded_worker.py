"""torchrun worker for tests/test_gpu_multi.py: user-sharded training + ranking on N GPUs, checked on rank 0
against the CPU oracle run on the same global batches (the single-GPU semantics)."""
import logging
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from daisyrec_b200.model.MFRecommender import MF
    from daisyrec_b200.utils.dataset import BasicDataset, CandidatesDataset, get_dataloader
    from oracle import oracle as orc
    from daisyrec_b200.model.AbstractRecommender import epoch_permutation

    rng = np.random.default_rng(5)
    U, I, F, T, B = 1500, 900, 64, 60_000, 4096
    users = np.minimum(U - 1, rng.zipf(1.3, size=T) - 1)
    data = np.stack([users, rng.integers(I, size=T), rng.integers(I, size=T)], 1).astype(np.int32)
    cfg = dict(gpu='', logger=logging.getLogger('w'), lr=0.01, reg_1=0.001, reg_2=0.001, epochs=2, topk=20, user_num=U,
               item_num=I, factors=F, loss_type='BPR', optimizer='default', init_method='default', early_stop=False,
               progress=False, sharded_comm=os.environ.get("DRB_SHARDED_COMM", "p2p"))
    torch.manual_seed(11 + 1000 * rank)          # different RNG histories per rank: rank 0's draws must win (broadcast)
    model = MF(cfg)
    P0 = model._P_full_cpu.numpy().copy()
    Q0 = model.embed_item.weight.cpu().numpy().copy()
    state = torch.get_rng_state()
    model.fit(get_dataloader(BasicDataset(data), batch_size=B, shuffle=True))
    P = model.gather_user_table().cpu().numpy()
    Q = model.embed_item.weight.cpu().numpy()
    # every rank must hold the same item table bit for bit
    q_all = [torch.empty_like(model.embed_item.weight) for _ in range(world)]
    dist.all_gather(q_all, model.embed_item.weight)
    same_q = all(torch.equal(q_all[0], q) for q in q_all)
    # ranking: all-gather of per-user top-K
    tu = rng.permutation(U)[:97].astype(np.int64)
    cands = rng.integers(I, size=(97, 300)).astype(np.int64)
    loader = get_dataloader(CandidatesDataset([[int(u), c] for u, c in zip(tu, cands)]), batch_size=128, shuffle=False)
    preds = model.rank(loader)
    ok = True
    if rank == 0:
        torch.set_rng_state(state)
        Po, Qo = P0.copy(), Q0.copy()
        hp = orc.hyper(0.01, 0.001, 0.001)
        for _ in range(2):
            perm = epoch_permutation(T, True).numpy().astype(np.int64)
            orc.mf_bpr_epoch(Po, Qo, np.ascontiguousarray(data), perm, B, hp)
        errP, errQ = np.abs(P - Po).max(), np.abs(Q - Qo).max()
        want = orc.mf_rank(P, Q, tu, cands, 20)
        ok = same_q and errP < 1e-5 and errQ < 1e-5 and np.array_equal(preds, want)
        print(f"[mp_sharded_worker] world={world} comm={model._trainer.comm if model._trainer else None} same_q={same_q} errP={errP:.2e} errQ={errQ:.2e} "
              f"rank_equal={np.array_equal(preds, want)} -> {'OK' if ok else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
