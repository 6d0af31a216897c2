"""Generate tests/golden/*.npz by RUNNING THE REAL REFERENCE (build container only).

TEST INFRASTRUCTURE.  Usage:  python -m oracle.gen_golden [--only name,...]
The reference (AmazingDD/daisyRec) has no tests or golden vectors of its own, so the
parity pins of this repo are outputs of the reference's own code, imported from
/root/reference through oracle/ref_harness.py (3 library-compat shims, documented there).
Each fixture stores the exact inputs and the reference's outputs; the files are small and
committed, because /root/reference does not exist on the GPU box.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _save(name, **arrs):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def _ur_from(users, items):
    import pandas as pd
    from daisy.utils.utils import get_ur
    return get_ur(pd.DataFrame({"user": users, "item": items}))


def _synthetic_inter(rng, U, I, nnz, dense_user=None):
    """Random implicit-feedback COO (unique (u,i) pairs), every user >= 1 item."""
    pairs = set()
    for u in range(U):
        pairs.add((u, int(rng.integers(I))))
    while len(pairs) < nnz:
        pairs.add((int(rng.integers(U)), int(min(I - 1, rng.zipf(1.3) - 1 if rng.random() < .5 else rng.integers(I)))))
    if dense_user is not None:               # a user who has seen all but 1 / all but 3 items
        for it in range(I):
            if it != 7:
                pairs.add((dense_user, it))
        for it in range(I):
            if it not in (2, 3, I - 1):
                pairs.add((dense_user + 1, it))
        pairs.discard((dense_user, 7))
        for it in (2, 3, I - 1):
            pairs.discard((dense_user + 1, it))
    pairs = np.array(sorted(pairs), dtype=np.int64)
    order = rng.permutation(len(pairs))       # "time" order
    return pairs[order, 0].astype(np.int32), pairs[order, 1].astype(np.int32)


# --------------------------------------------------------------------------- sampler
def gen_sampler_small():
    """daisy/utils/sampler.py:55-103 on a small synthetic set (incl. near-full users, G=1..5)."""
    import pandas as pd
    from daisy.utils.sampler import BasicNegtiveSampler
    rng = np.random.default_rng(7)
    out = {}
    for case, (U, I, nnz, G, seed, dense) in enumerate([(50, 80, 600, 4, 2022, 10), (31, 1025, 900, 1, 5, None),
                                                         (17, 33, 200, 5, 123456789, 3), (64, 4096, 3000, 3, 0, None)]):
        cu, ci = _synthetic_inter(rng, U, I, nnz, dense)
        df = pd.DataFrame({"user": cu, "item": ci, "rating": 1.0, "timestamp": np.arange(len(cu))})
        cfg = rh.make_config("mf", user_num=U, item_num=I, num_ng=G, train_ur=_ur_from(cu, ci))
        np.random.seed(seed)
        triples = BasicNegtiveSampler(df, cfg).sampling()
        nxt = np.random.randint(0, 2 ** 31 - 1, size=3)            # pins how many MT words were consumed
        out.update({f"c{case}_coo_u": cu, f"c{case}_coo_i": ci, f"c{case}_triples": triples,
                    f"c{case}_meta": np.array([U, I, G, seed], np.int64), f"c{case}_next": nxt})
    out["ncases"] = np.array(4)
    _save("sampler_small", **out)


def _ml100k(factors=32, epochs=1, **kw):
    cfg = rh.make_config("mf", factors=factors, epochs=epochs, **kw)
    rh.seed_everything(cfg["seed"])
    art = rh.load_ml100k(cfg)
    return cfg, art


def gen_ml100k_pipeline():
    """Config 1 of BASELINE.json: test.py:41-120 with MF+BPR, factors=32, 1 epoch, CPU, seed 2022.

    Follows the driver's order of RNG consumption: seed -> (data, split: no RNG) -> MF(config)
    [torch RNG: init] -> sampling() [numpy RNG] -> fit [torch RNG: DataLoader permutation] ->
    build_candidates_set [numpy RNG] -> rank.
    """
    import torch
    from daisy.model.MFRecommender import MF
    from daisy.utils.sampler import BasicNegtiveSampler
    from daisy.utils.dataset import BasicDataset, CandidatesDataset, get_dataloader
    from daisy.utils.utils import build_candidates_set
    from daisy.utils.metrics import calc_ranking_results
    cfg, art = _ml100k()
    train_set, test_ur, train_ur = art["train_set"], art["test_ur"], art["train_ur"]
    model = MF(cfg)
    P0 = model.embed_user.weight.detach().numpy().copy()
    Q0 = model.embed_item.weight.detach().numpy().copy()
    coo_u = train_set["user"].values.astype(np.int32)
    coo_i = train_set["item"].values.astype(np.int32)
    triples = BasicNegtiveSampler(train_set, cfg).sampling()

    class Rec(BasicDataset):                                   # records the DataLoader's index order
        def __init__(self, s):
            super().__init__(s)
            self.order = []

        def __getitem__(self, idx):
            self.order.append(idx)
            return super().__getitem__(idx)

    ds = Rec(triples)
    loader = get_dataloader(ds, batch_size=cfg["batch_size"], shuffle=True, num_workers=0)
    torch_state = torch.get_rng_state().numpy().copy()          # state right before fit()
    step_losses = []
    orig = model.calc_loss

    def rec_loss(batch):
        l = orig(batch)
        step_losses.append(float(l.item()))
        return l

    model.calc_loss = rec_loss
    model.fit(loader)
    perm = np.array(ds.order, np.int32)
    P1 = model.embed_user.weight.detach().numpy().copy()
    Q1 = model.embed_item.weight.detach().numpy().copy()

    test_u, test_ucands = build_candidates_set(test_ur, train_ur, cfg)
    cands = np.stack([c[1] for c in test_ucands]).astype(np.int16)
    gt_flat = np.concatenate([np.array(list(test_ur[u]), np.int32) for u in test_u])
    gt_len = np.array([len(test_ur[u]) for u in test_u], np.int32)
    nxt = np.random.randint(0, 2 ** 31 - 1, size=3)
    loader_t = get_dataloader(CandidatesDataset(test_ucands), batch_size=128, shuffle=False, num_workers=0)
    preds = model.rank(loader_t)
    full = np.stack([model.full_rank(int(u)) for u in test_u[:16]])
    pred_pairs = np.array([model.predict(int(test_u[k]), int(cands[k][-1])) for k in range(8)], np.float32)
    import tempfile
    cfg['res_path'] = tempfile.mkdtemp() + '/'
    res = calc_ranking_results(test_ur, preds, test_u, cfg)

    _save("ml100k_sampler", coo_u=coo_u.astype(np.int16), coo_i=coo_i.astype(np.int16), triples_j=triples[:, 2].astype(np.int16),
          meta=np.array([cfg["user_num"], cfg["item_num"], cfg["num_ng"], cfg["seed"]], np.int64))
    _save("ml100k_fit", P0=P0, Q0=Q0, P1=P1, Q1=Q1, perm=perm, step_losses=np.array(step_losses, np.float64),
          torch_state=torch_state,
          hyper=np.array([cfg["lr"], cfg["reg_1"], cfg["reg_2"], cfg["batch_size"], cfg["factors"]], np.float64))
    _save("ml100k_rank", test_u=np.array(test_u, np.int32), cands=cands, gt_flat=gt_flat, gt_len=gt_len,
          next=nxt, preds=preds, full=full, pred_pairs=pred_pairs, topk=np.array(cfg["topk"]),
          kpi=res.values[:, 1:].astype(np.float64) if res.shape[1] > 1 else res.values.astype(np.float64))
    print(res)


# --------------------------------------------------------------------------- steps
def gen_mf_steps():
    """MF.calc_loss + backward + optimizer.step (MFRecommender.py:70-97, AbstractRecommender.py:119-126)
    on small random tables: several F, reg on/off, duplicate-heavy batches, SGD and Adam, 3 steps each."""
    import torch
    from daisy.model.MFRecommender import MF
    out = {}
    cases = [  # U, I, F, B, lr, reg1, reg2, opt, seed
        (40, 60, 8, 64, 0.01, 0.001, 0.001, "sgd", 1),
        (40, 60, 32, 256, 0.05, 0.0, 0.0, "sgd", 2),
        (25, 30, 100, 200, 0.01, 0.01, 0.02, "sgd", 3),     # reference default F=100 (not a power of two)
        (13, 9, 6, 50, 0.01, 0.001, 0.001, "sgd", 4),       # F%4 != 0, heavy duplicates
        (40, 60, 64, 128, 0.001, 0.001, 0.001, "adam", 5),
        (20, 20, 7, 33, 0.01, 0.0, 0.001, "adam", 6),       # odd F
        (40, 60, 32, 128, 0.01, 0.001, 0.001, "sgd", 7, "HL"),   # HingeLoss (loss.py:16-23)
        (40, 60, 64, 128, 0.01, 0.001, 0.001, "sgd", 8, "TL"),   # TOP1Loss (loss.py:26-33)
        (30, 30, 16, 96, 0.001, 0.0, 0.0, "adam", 9, "TL"),
    ]
    for k, case in enumerate(cases):
        U, I, F, B, lr, r1, r2, opt, seed = case[:9]
        loss_type = case[9] if len(case) > 9 else "BPR"
        cfg = rh.make_config("mf", user_num=U, item_num=I, factors=F, lr=lr, reg_1=r1, reg_2=r2, optimizer=opt,
                             epochs=1, loss_type=loss_type)
        torch.manual_seed(seed)
        model = MF(cfg)
        with torch.no_grad():                                 # larger weights so the loss is not ~log 2 everywhere
            model.embed_user.weight.mul_(30.0)
            model.embed_item.weight.mul_(30.0)
        model.criterion = model._build_criterion(model.loss_type)
        optim = model._build_optimizer(optimizer=model.optimizer, lr=model.lr)
        rng = np.random.default_rng(seed)
        P = [model.embed_user.weight.detach().numpy().copy()]
        Q = [model.embed_item.weight.detach().numpy().copy()]
        batches, losses = [], []
        for step in range(3):
            b = np.stack([rng.integers(U, size=B), rng.integers(I, size=B), rng.integers(I, size=B)]).astype(np.int32)
            batches.append(b)
            model.zero_grad()
            loss = model.calc_loss([torch.from_numpy(b[0]), torch.from_numpy(b[1]), torch.from_numpy(b[2])])
            loss.backward()
            optim.step()
            losses.append(float(loss.item()))
            P.append(model.embed_user.weight.detach().numpy().copy())
            Q.append(model.embed_item.weight.detach().numpy().copy())
        out.update({f"c{k}_P": np.stack(P), f"c{k}_Q": np.stack(Q), f"c{k}_batches": np.stack(batches),
                    f"c{k}_loss": np.array(losses, np.float64),
                    f"c{k}_hyper": np.array([lr, r1, r2, 0 if opt == "sgd" else 1,
                                             {"BPR": 0, "HL": 1, "TL": 2}[loss_type]], np.float64)})
    out["ncases"] = np.array(len(cases))
    _save("mf_steps", **out)


def gen_mf_pointwise():
    """Point-wise branch of MF.calc_loss (MFRecommender.py:75-81: CL = BCEWithLogitsLoss(sum), SL = MSELoss(sum),
    AbstractRecommender.py:79-82) + backward + optimizer.step; batch[2] holds the label (sampler.py:93-98)."""
    import torch
    from daisy.model.MFRecommender import MF
    out = {}
    cases = [  # U, I, F, B, lr, reg1, reg2, opt, seed, loss
        (40, 60, 8, 64, 0.01, 0.001, 0.001, "sgd", 11, "CL"),
        (40, 60, 32, 256, 0.05, 0.0, 0.0, "sgd", 12, "CL"),
        (25, 30, 100, 200, 0.001, 0.01, 0.02, "sgd", 13, "SL"),
        (13, 9, 6, 50, 0.001, 0.001, 0.001, "sgd", 14, "SL"),    # F%4 != 0, heavy duplicates
        (40, 60, 64, 128, 0.001, 0.001, 0.001, "adam", 15, "CL"),
        (20, 20, 7, 33, 0.01, 0.0, 0.001, "adam", 16, "SL"),
    ]
    for k, (U, I, F, B, lr, r1, r2, opt, seed, loss_type) in enumerate(cases):
        cfg = rh.make_config("mf", user_num=U, item_num=I, factors=F, lr=lr, reg_1=r1, reg_2=r2, optimizer=opt,
                             epochs=1, loss_type=loss_type)
        torch.manual_seed(seed)
        model = MF(cfg)
        with torch.no_grad():
            model.embed_user.weight.mul_(30.0)
            model.embed_item.weight.mul_(30.0)
        model.criterion = model._build_criterion(model.loss_type)        # what fit() does first (:108)
        optim = model._build_optimizer(optimizer=model.optimizer, lr=model.lr)
        rng = np.random.default_rng(seed)
        P = [model.embed_user.weight.detach().numpy().copy()]
        Q = [model.embed_item.weight.detach().numpy().copy()]
        batches, losses = [], []
        for step in range(3):
            label = rng.integers(0, 2, size=B) if loss_type == "CL" else rng.integers(0, 6, size=B)
            b = np.stack([rng.integers(U, size=B), rng.integers(I, size=B), label]).astype(np.int32)
            batches.append(b)
            model.zero_grad()
            loss = model.calc_loss([torch.from_numpy(b[0]), torch.from_numpy(b[1]), torch.from_numpy(b[2])])
            loss.backward()
            optim.step()
            losses.append(float(loss.item()))
            P.append(model.embed_user.weight.detach().numpy().copy())
            Q.append(model.embed_item.weight.detach().numpy().copy())
        out.update({f"c{k}_P": np.stack(P), f"c{k}_Q": np.stack(Q), f"c{k}_batches": np.stack(batches),
                    f"c{k}_loss": np.array(losses, np.float64),
                    f"c{k}_hyper": np.array([lr, r1, r2, 0 if opt == "sgd" else 1, {"CL": 3, "SL": 4}[loss_type]],
                                            np.float64)})
        print(f"mf_pointwise case {k} ({loss_type}/{opt}): losses {losses}")
    out["ncases"] = np.array(len(cases))
    _save("mf_pointwise", **out)


def gen_mf_optim():
    """optim.Adagrad / optim.RMSprop (AbstractRecommender.py:57-60, torch defaults) behind MF.calc_loss, 3 steps each."""
    import torch
    from daisy.model.MFRecommender import MF
    out = {}
    cases = [  # U, I, F, B, lr, reg1, reg2, opt, seed, loss
        (40, 60, 32, 128, 0.01, 0.001, 0.001, "adagrad", 21, "BPR"),
        (40, 60, 64, 256, 0.001, 0.001, 0.001, "rmsprop", 22, "BPR"),
        (13, 9, 6, 50, 0.01, 0.0, 0.0, "adagrad", 23, "HL"),
        (25, 30, 100, 200, 0.001, 0.01, 0.02, "rmsprop", 24, "CL"),
    ]
    for k, (U, I, F, B, lr, r1, r2, opt, seed, loss_type) in enumerate(cases):
        cfg = rh.make_config("mf", user_num=U, item_num=I, factors=F, lr=lr, reg_1=r1, reg_2=r2, optimizer=opt,
                             epochs=1, loss_type=loss_type)
        torch.manual_seed(seed)
        model = MF(cfg)
        with torch.no_grad():
            model.embed_user.weight.mul_(30.0)
            model.embed_item.weight.mul_(30.0)
        model.criterion = model._build_criterion(model.loss_type)
        optim = model._build_optimizer(optimizer=model.optimizer, lr=model.lr)
        assert type(optim).__name__.lower() == opt
        rng = np.random.default_rng(seed)
        P = [model.embed_user.weight.detach().numpy().copy()]
        Q = [model.embed_item.weight.detach().numpy().copy()]
        batches, losses = [], []
        for step in range(3):
            third = rng.integers(0, 2, size=B) if loss_type == "CL" else rng.integers(I, size=B)
            b = np.stack([rng.integers(U, size=B), rng.integers(I, size=B), third]).astype(np.int32)
            batches.append(b)
            model.zero_grad()
            loss = model.calc_loss([torch.from_numpy(b[0]), torch.from_numpy(b[1]), torch.from_numpy(b[2])])
            loss.backward()
            optim.step()
            losses.append(float(loss.item()))
            P.append(model.embed_user.weight.detach().numpy().copy())
            Q.append(model.embed_item.weight.detach().numpy().copy())
        out.update({f"c{k}_P": np.stack(P), f"c{k}_Q": np.stack(Q), f"c{k}_batches": np.stack(batches),
                    f"c{k}_loss": np.array(losses, np.float64), f"c{k}_opt": np.array(opt), f"c{k}_losskind": np.array(loss_type),
                    f"c{k}_hyper": np.array([lr, r1, r2], np.float64)})
        print(f"mf_optim case {k} ({opt}/{loss_type}): losses {losses}")
    out["ncases"] = np.array(len(cases))
    _save("mf_optim", **out)


# --------------------------------------------------------------------------- rank
def gen_mf_rank():
    """MF.rank / full_rank / predict (MFRecommender.py:99-133) on fixed random tables; candidate
    lists contain duplicates (sampled with replacement, utils.py:79); 130 users = batches 128 + 2."""
    import torch
    from daisy.model.MFRecommender import MF
    from daisy.utils.dataset import CandidatesDataset, get_dataloader
    out = {}
    cases = [(300, 2000, 32, 130, 1000, 50, 11), (90, 700, 100, 40, 1000, 50, 12), (50, 120, 64, 7, 100, 10, 13),
             (64, 1500, 128, 9, 1000, 50, 14)]
    for k, (U, I, F, n, C, K, seed) in enumerate(cases):
        cfg = rh.make_config("mf", user_num=U, item_num=I, factors=F, topk=K, cand_num=C)
        torch.manual_seed(seed)
        model = MF(cfg)
        rng = np.random.default_rng(seed)
        users = rng.permutation(U)[:n].astype(np.int64)
        cands = rng.integers(I, size=(n, C)).astype(np.int64)
        cands[:, -3:] = cands[:, :3]                           # forced duplicates
        ucands = [[int(users[r]), cands[r]] for r in range(n)]
        loader = get_dataloader(CandidatesDataset(ucands), batch_size=128, shuffle=False, num_workers=0)
        preds = model.rank(loader)
        P = model.embed_user.weight.detach().numpy().copy()
        Q = model.embed_item.weight.detach().numpy().copy()
        # ambiguity screen: reference fp32 scores (float64 recomputation) gaps inside the top-(K+1)
        sc = np.einsum("nf,ncf->nc", P[users].astype(np.float64), Q[cands].astype(np.float64))
        srt = -np.sort(-sc, axis=1)[:, :K + 1]
        gaps = srt[:, :-1] - srt[:, 1:]
        gaps[gaps == 0] = np.inf                               # exact duplicates
        min_gap_rel = float((gaps / np.abs(srt[:, :-1]).clip(1e-30)).min())
        full = np.stack([model.full_rank(int(u)) for u in users[:5]])
        out.update({f"c{k}_P": P, f"c{k}_Q": Q, f"c{k}_users": users, f"c{k}_cands": cands.astype(np.int32),
                    f"c{k}_preds": preds, f"c{k}_full": full, f"c{k}_K": np.array(K),
                    f"c{k}_min_gap_rel": np.array(min_gap_rel)})
        print(f"rank case {k}: preds {preds.shape} {preds.dtype}, min relative score gap in top-K+1 = {min_gap_rel:.3e}")
    out["ncases"] = np.array(len(cases))
    _save("mf_rank", **out)


# --------------------------------------------------------------------------- FM
def gen_fm():
    """FM (daisy/model/FMRecommender.py:16-131): MF's factor product + u_bias + i_bias + bias_; 3 training steps per case
    (biases are zero-initialised (:58-59), so they are set to random values first), then rank / full_rank / predict."""
    import torch
    from daisy.model.FMRecommender import FM
    from daisy.utils.dataset import CandidatesDataset, get_dataloader
    out = {}
    cases = [  # U, I, F, B, lr, reg1, reg2, opt, seed, loss
        (40, 60, 8, 64, 0.01, 0.001, 0.001, "sgd", 31, "BPR"),
        (40, 60, 32, 256, 0.01, 0.0, 0.0, "sgd", 32, "CL"),
        (25, 30, 100, 200, 0.001, 0.01, 0.02, "adam", 33, "TL"),
        (13, 9, 6, 50, 0.001, 0.001, 0.001, "adam", 34, "BPR"),
        (30, 500, 64, 128, 0.01, 0.001, 0.001, "sgd", 35, "HL"),
    ]
    for k, (U, I, F, B, lr, r1, r2, opt, seed, loss_type) in enumerate(cases):
        cfg = rh.make_config("fm", user_num=U, item_num=I, factors=F, lr=lr, reg_1=r1, reg_2=r2, optimizer=opt,
                             epochs=1, loss_type=loss_type, topk=10, cand_num=100)
        torch.manual_seed(seed)
        model = FM(cfg)
        init_tabs = (model.embed_user.weight.detach().numpy().copy(), model.embed_item.weight.detach().numpy().copy())
        with torch.no_grad():
            model.embed_user.weight.mul_(30.0)
            model.embed_item.weight.mul_(30.0)
            model.u_bias.weight.normal_(0, 0.3)
            model.i_bias.weight.normal_(0, 0.3)
            model.bias_.fill_(0.25)
        model.criterion = model._build_criterion(model.loss_type)
        optim = model._build_optimizer(optimizer=model.optimizer, lr=model.lr)
        rng = np.random.default_rng(seed)

        def snap():
            return (model.embed_user.weight.detach().numpy().copy(), model.embed_item.weight.detach().numpy().copy(),
                    np.concatenate([model.u_bias.weight.detach().numpy().ravel(), model.i_bias.weight.detach().numpy().ravel(),
                                    model.bias_.detach().numpy().ravel()]).astype(np.float32))
        snaps, batches, losses = [snap()], [], []
        for step in range(3):
            third = rng.integers(0, 2, size=B) if loss_type == "CL" else rng.integers(I, size=B)
            b = np.stack([rng.integers(U, size=B), rng.integers(I, size=B), third]).astype(np.int32)
            batches.append(b)
            model.zero_grad()
            loss = model.calc_loss([torch.from_numpy(b[0]), torch.from_numpy(b[1]), torch.from_numpy(b[2])])
            loss.backward()
            optim.step()
            losses.append(float(loss.item()))
            snaps.append(snap())
        n, C, K = min(U, 20), 100, 10
        users = rng.permutation(U)[:n].astype(np.int64)
        cands = rng.integers(I, size=(n, C)).astype(np.int64)
        loader = get_dataloader(CandidatesDataset([[int(users[r]), cands[r]] for r in range(n)]), batch_size=128,
                                shuffle=False, num_workers=0)
        with torch.no_grad():
            preds = model.rank(loader)
            full = np.stack([model.full_rank(int(u)) for u in users[:4]])
            pp = np.array([model.predict(int(users[q]), int(cands[q][0])) for q in range(4)], np.float32)
        out.update({f"c{k}_P": np.stack([s_[0] for s_ in snaps]), f"c{k}_Q": np.stack([s_[1] for s_ in snaps]),
                    f"c{k}_bias": np.stack([s_[2] for s_ in snaps]), f"c{k}_batches": np.stack(batches),
                    f"c{k}_loss": np.array(losses, np.float64), f"c{k}_opt": np.array(opt), f"c{k}_losskind": np.array(loss_type),
                    f"c{k}_hyper": np.array([lr, r1, r2], np.float64), f"c{k}_users": users, f"c{k}_cands": cands.astype(np.int32),
                    f"c{k}_preds": preds, f"c{k}_full": full, f"c{k}_pred_pairs": pp,
                    f"c{k}_P_init": init_tabs[0], f"c{k}_Q_init": init_tabs[1]})
        print(f"fm case {k} ({loss_type}/{opt}): losses {losses}")
    out["ncases"] = np.array(len(cases))
    _save("fm", **out)


# --------------------------------------------------------------------------- LightGCN
def gen_lightgcn():
    """LightGCN (LightGCNRecommender.py:73-211): normalised adjacency, forward propagation, 3 training steps
    (Adam default and SGD, reg on/off, L=2/3), rank / full_rank / predict on the propagated tables."""
    import pandas as pd
    import torch
    from daisy.model.LightGCNRecommender import LightGCN
    from daisy.utils.utils import get_inter_matrix
    from daisy.utils.dataset import CandidatesDataset, get_dataloader
    out = {}
    cases = [  # U, I, nnz, F, L, B, lr, reg1, reg2, opt, seed
        (60, 90, 700, 16, 2, 128, 0.01, 0.0, 0.0, "default", 21),
        (40, 50, 400, 64, 3, 256, 0.01, 0.001, 0.002, "default", 22),
        (30, 45, 300, 8, 3, 64, 0.05, 0.001, 0.001, "sgd", 23),
        (25, 30, 200, 6, 1, 50, 0.01, 0.0, 0.0, "sgd", 24),
    ]
    rng0 = np.random.default_rng(99)
    for k, (U, I, nnz, F, L, B, lr, r1, r2, opt, seed) in enumerate(cases):
        cu, ci = _synthetic_inter(rng0, U, I, nnz)
        df = pd.DataFrame({"user": cu, "item": ci, "rating": 1.0, "timestamp": np.arange(len(cu))})
        cfg = rh.make_config("lightgcn", user_num=U, item_num=I, factors=F, num_layers=L, lr=lr, reg_1=r1, reg_2=r2,
                             optimizer=opt, epochs=1, topk=10, cand_num=40)
        cfg["inter_matrix"] = get_inter_matrix(df, cfg)
        torch.manual_seed(seed)
        model = LightGCN(cfg)
        adj = model.norm_adj_matrix.coalesce()
        model.criterion = model._build_criterion(model.loss_type)
        optim = model._build_optimizer(optimizer=model.optimizer, lr=model.lr)
        rng = np.random.default_rng(seed)
        E = [torch.cat([model.embed_user.weight, model.embed_item.weight]).detach().numpy().copy()]
        with torch.no_grad():
            eu, ei = model.forward()
        Em0 = torch.cat([eu, ei]).numpy().copy()
        batches, losses = [], []
        for step in range(3):
            b = np.stack([rng.integers(U, size=B), rng.integers(I, size=B), rng.integers(I, size=B)]).astype(np.int32)
            batches.append(b)
            model.zero_grad()
            loss = model.calc_loss([torch.from_numpy(b[0]), torch.from_numpy(b[1]), torch.from_numpy(b[2])])
            loss.backward()
            optim.step()
            losses.append(float(loss.item()))
            E.append(torch.cat([model.embed_user.weight, model.embed_item.weight]).detach().numpy().copy())
        users = rng.permutation(U)[:9].astype(np.int64)
        cands = rng.integers(I, size=(9, 40)).astype(np.int64)
        loader = get_dataloader(CandidatesDataset([[int(u), c] for u, c in zip(users, cands)]), batch_size=128,
                                shuffle=False, num_workers=0)
        with torch.no_grad():
            preds = model.rank(loader)
            full = np.stack([model.full_rank(int(u)) for u in users[:4]])
            pp = np.array([model.predict(int(users[0]), int(cands[0][0]))], np.float32)
            Em_final = torch.cat([model.restore_user_e, model.restore_item_e]).numpy().copy()
        out.update({f"c{k}_coo_u": cu, f"c{k}_coo_i": ci, f"c{k}_adj_idx": adj.indices().numpy().astype(np.int32),
                    f"c{k}_adj_val": adj.values().numpy(), f"c{k}_E": np.stack(E), f"c{k}_Em0": Em0,
                    f"c{k}_batches": np.stack(batches), f"c{k}_loss": np.array(losses, np.float64),
                    f"c{k}_hyper": np.array([U, I, F, L, lr, r1, r2, 0 if opt == "sgd" else 1], np.float64),
                    f"c{k}_users": users, f"c{k}_cands": cands.astype(np.int32), f"c{k}_preds": preds, f"c{k}_full": full,
                    f"c{k}_pred_pair": pp, f"c{k}_Em_final": Em_final})
        print(f"lightgcn case {k}: losses {losses}")
    out["ncases"] = np.array(len(cases))
    _save("lightgcn", **out)


# --------------------------------------------------------------------------- NGCF
def gen_ngcf():
    """NGCF (NGCFRecommender.py:38-59,157-252) with node_dropout = mess_dropout = 0: forward (concatenated layer outputs),
    3 training steps (Adam default / SGD, reg on / off, 1-3 layers of unequal width), rank / full_rank / predict."""
    import pandas as pd
    import torch
    from daisy.model.NGCFRecommender import NGCF
    from daisy.utils.utils import get_inter_matrix
    from daisy.utils.dataset import CandidatesDataset, get_dataloader
    out = {}
    cases = [  # U, I, nnz, F, hidden, B, lr, reg1, reg2, opt, seed
        (60, 90, 700, 16, [16, 16], 128, 0.01, 0.0, 0.0, "default", 41),
        (40, 50, 400, 36, [64, 64, 64], 256, 0.01, 0.001, 0.002, "default", 42),       # assets/ngcf.yaml defaults
        (30, 45, 300, 8, [12, 6], 64, 0.05, 0.001, 0.001, "sgd", 43),
        (25, 30, 200, 6, [10], 50, 0.01, 0.0, 0.0, "sgd", 44),
    ]
    rng0 = np.random.default_rng(77)

    def flat(model):
        ws = []
        for g in model.gnn_layers:
            ws += [g.linear.weight, g.linear.bias, g.interact_transform.weight, g.interact_transform.bias]
        return np.concatenate([w.detach().numpy().ravel() for w in ws]).astype(np.float32)
    for k, (U, I, nnz, F, hidden, B, lr, r1, r2, opt, seed) in enumerate(cases):
        cu, ci = _synthetic_inter(rng0, U, I, nnz)
        df = pd.DataFrame({"user": cu, "item": ci, "rating": 1.0, "timestamp": np.arange(len(cu))})
        cfg = rh.make_config("ngcf", user_num=U, item_num=I, factors=F, hidden_size_list=hidden, node_dropout=0.0,
                             mess_dropout=0.0, lr=lr, reg_1=r1, reg_2=r2, optimizer=opt, epochs=1, topk=10, cand_num=40)
        cfg["inter_matrix"] = get_inter_matrix(df, cfg)
        torch.manual_seed(seed)
        model = NGCF(cfg)
        model.criterion = model._build_criterion(model.loss_type)
        optim = model._build_optimizer(optimizer=model.optimizer, lr=model.lr)
        rng = np.random.default_rng(seed)
        E = [torch.cat([model.embed_user.weight, model.embed_item.weight]).detach().numpy().copy()]
        Ws = [flat(model)]
        with torch.no_grad():
            eu, ei = model.forward()
        all0 = torch.cat([eu, ei]).numpy().copy()
        batches, losses = [], []
        for step in range(3):
            b = np.stack([rng.integers(U, size=B), rng.integers(I, size=B), rng.integers(I, size=B)]).astype(np.int32)
            batches.append(b)
            model.zero_grad()
            loss = model.calc_loss([torch.from_numpy(b[0]), torch.from_numpy(b[1]), torch.from_numpy(b[2])])
            loss.backward()
            optim.step()
            losses.append(float(loss.item()))
            E.append(torch.cat([model.embed_user.weight, model.embed_item.weight]).detach().numpy().copy())
            Ws.append(flat(model))
        users = rng.permutation(U)[:9].astype(np.int64)
        cands = rng.integers(I, size=(9, 40)).astype(np.int64)
        loader = get_dataloader(CandidatesDataset([[int(u), c] for u, c in zip(users, cands)]), batch_size=128,
                                shuffle=False, num_workers=0)
        with torch.no_grad():
            preds = model.rank(loader)
            full = np.stack([model.full_rank(int(u)) for u in users[:4]])
            pp = np.array([model.predict(int(users[0]), int(cands[0][0]))], np.float32)
        out.update({f"c{k}_coo_u": cu, f"c{k}_coo_i": ci, f"c{k}_E": np.stack(E), f"c{k}_W": np.stack(Ws), f"c{k}_all0": all0,
                    f"c{k}_batches": np.stack(batches), f"c{k}_loss": np.array(losses, np.float64),
                    f"c{k}_dims": np.array([F] + hidden, np.int32),
                    f"c{k}_hyper": np.array([U, I, lr, r1, r2, 0 if opt == "sgd" else 1], np.float64),
                    f"c{k}_users": users, f"c{k}_cands": cands.astype(np.int32), f"c{k}_preds": preds, f"c{k}_full": full,
                    f"c{k}_pred_pair": pp})
        print(f"ngcf case {k}: losses {losses}")
    out["ncases"] = np.array(len(cases))
    _save("ngcf", **out)


def gen_ngcf_dropout():
    """NGCF at the reference's DEFAULT message dropout (assets/ngcf.yaml: mess_dropout 0.1, node_dropout 0): forward() builds a
    fresh nn.Dropout per layer (:164) -- a module in training mode, so the masks are drawn on EVERY forward(), the one behind
    rank() / full_rank() / predict() included (one draw per layer over the [n, width] layer output, torch's global CPU generator).
    Per case: the forward right after seeding, 3 training steps from a seeded generator, then rank() from another seed."""
    import pandas as pd
    import torch
    from daisy.model.NGCFRecommender import NGCF
    from daisy.utils.utils import get_inter_matrix
    from daisy.utils.dataset import CandidatesDataset, get_dataloader
    out = {}
    cases = [  # U, I, nnz, F, hidden, B, lr, reg1, reg2, opt, mess_dropout, seed
        (40, 50, 400, 36, [64, 64, 64], 256, 0.01, 0.0, 0.0, "default", 0.1, 71),        # assets/ngcf.yaml defaults
        (30, 45, 300, 8, [12, 6], 64, 0.05, 0.001, 0.001, "sgd", 0.3, 72),
        (25, 30, 200, 6, [10], 50, 0.01, 0.0, 0.001, "sgd", 0.5, 73),
    ]
    rng0 = np.random.default_rng(78)

    def flat(model):
        ws = []
        for g in model.gnn_layers:
            ws += [g.linear.weight, g.linear.bias, g.interact_transform.weight, g.interact_transform.bias]
        return np.concatenate([w.detach().numpy().ravel() for w in ws]).astype(np.float32)
    for k, (U, I, nnz, F, hidden, B, lr, r1, r2, opt, drop, seed) in enumerate(cases):
        cu, ci = _synthetic_inter(rng0, U, I, nnz)
        df = pd.DataFrame({"user": cu, "item": ci, "rating": 1.0, "timestamp": np.arange(len(cu))})
        cfg = rh.make_config("ngcf", user_num=U, item_num=I, factors=F, hidden_size_list=hidden, node_dropout=0.0,
                             mess_dropout=drop, lr=lr, reg_1=r1, reg_2=r2, optimizer=opt, epochs=1, topk=10, cand_num=40)
        cfg["inter_matrix"] = get_inter_matrix(df, cfg)
        torch.manual_seed(seed)
        model = NGCF(cfg)
        model.criterion = model._build_criterion(model.loss_type)
        optim = model._build_optimizer(optimizer=model.optimizer, lr=model.lr)
        rng = np.random.default_rng(seed)
        E = [torch.cat([model.embed_user.weight, model.embed_item.weight]).detach().numpy().copy()]
        Ws = [flat(model)]
        torch.manual_seed(seed + 50)
        with torch.no_grad():
            eu, ei = model.forward()
        all0 = torch.cat([eu, ei]).numpy().copy()
        batches, losses = [], []
        torch.manual_seed(seed + 100)                                   # the masks of the 3 steps come from here
        for step in range(3):
            b = np.stack([rng.integers(U, size=B), rng.integers(I, size=B), rng.integers(I, size=B)]).astype(np.int32)
            batches.append(b)
            model.zero_grad()
            loss = model.calc_loss([torch.from_numpy(b[0]), torch.from_numpy(b[1]), torch.from_numpy(b[2])])
            loss.backward()
            optim.step()
            losses.append(float(loss.item()))
            E.append(torch.cat([model.embed_user.weight, model.embed_item.weight]).detach().numpy().copy())
            Ws.append(flat(model))
        rng_after = torch.get_rng_state().numpy().copy()
        users = rng.permutation(U)[:9].astype(np.int64)
        cands = rng.integers(I, size=(9, 40)).astype(np.int64)
        loader = get_dataloader(CandidatesDataset([[int(u), c] for u, c in zip(users, cands)]), batch_size=128,
                                shuffle=False, num_workers=0)
        torch.manual_seed(seed + 200)                                   # rank() runs forward() once: masks from here
        with torch.no_grad():
            preds = model.rank(loader)
            all_rank = torch.cat([model.restore_user_e, model.restore_item_e]).numpy().copy()
        out.update({f"c{k}_coo_u": cu, f"c{k}_coo_i": ci, f"c{k}_E": np.stack(E), f"c{k}_W": np.stack(Ws), f"c{k}_all0": all0,
                    f"c{k}_batches": np.stack(batches), f"c{k}_loss": np.array(losses, np.float64),
                    f"c{k}_dims": np.array([F] + hidden, np.int32),
                    f"c{k}_hyper": np.array([U, I, lr, r1, r2, 0 if opt == "sgd" else 1, drop, seed], np.float64),
                    f"c{k}_users": users, f"c{k}_cands": cands.astype(np.int32), f"c{k}_preds": preds, f"c{k}_all_rank": all_rank,
                    f"c{k}_rng_after": rng_after})
        print(f"ngcf_dropout case {k} (p={drop}): losses {losses}")
    out["ncases"] = np.array(len(cases))
    _save("ngcf_dropout", **out)


# --------------------------------------------------------------------------- NFM
def gen_nfm():
    """NFM (NFMRecommender.py:14-209) with dropout = 0: bi-interaction + [BatchNorm] + MLP + broadcast biases + prediction;
    3 training steps (separate positive / negative forward calls => separate batch statistics), eval-mode ranking."""
    import torch
    from daisy.model.NFMRecommender import NFM
    from daisy.utils.dataset import CandidatesDataset, get_dataloader
    # compat shim (torch >= 2: F.dropout(x, p=0) hands back x itself, so the in-place `fm += ...` of :120 invalidates the
    # activation's saved output and backward() raises; the torch the reference was written for returned a new tensor).
    # Same arithmetic, only the aliasing differs.  Restored after the run.
    _drop_fwd = torch.nn.Dropout.forward
    torch.nn.Dropout.forward = lambda self, x: x.clone() if self.p == 0 else _drop_fwd(self, x)
    out = {}
    cases = [  # U, I, F, L, bn, act, B, lr, reg1, reg2, opt, seed
        (40, 60, 8, 2, True, "relu", 64, 0.01, 0.0, 0.0, "sgd", 51),
        (40, 60, 30, 2, True, "relu", 128, 0.001, 0.001, 0.002, "adam", 52),         # assets/nfm.yaml shape
        (25, 30, 12, 1, False, "tanh", 50, 0.01, 0.001, 0.001, "sgd", 53),
        (30, 45, 16, 3, True, "sigmoid", 96, 0.01, 0.0, 0.0, "sgd", 54),
        (20, 20, 6, 2, False, "relu", 33, 0.001, 0.0, 0.001, "adam", 55),
    ]
    ACT = {"relu": 0, "sigmoid": 1, "tanh": 2}
    for k, (U, I, F, L, bn, act, B, lr, r1, r2, opt, seed) in enumerate(cases):
        cfg = rh.make_config("nfm", user_num=U, item_num=I, factors=F, num_layers=L, batch_norm=bn, act_function=act, dropout=0.0,
                             lr=lr, reg_1=r1, reg_2=r2, optimizer=opt, epochs=1, topk=10, cand_num=40)
        torch.manual_seed(seed)
        model = NFM(cfg)
        with torch.no_grad():
            model.embed_user.weight.mul_(3.0)
            model.embed_item.weight.mul_(3.0)
            model.u_bias.weight.normal_(0, 0.3)
            model.i_bias.weight.normal_(0, 0.3)
            model.bias_.fill_(0.25)
        model.criterion = model._build_criterion(model.loss_type)
        optim = model._build_optimizer(optimizer=model.optimizer, lr=model.lr)
        rng = np.random.default_rng(seed)
        bns = [m for m in list(model.FM_layers) + list(model.deep_layers) if isinstance(m, torch.nn.BatchNorm1d)]

        def snap():
            net = []
            for m in list(model.FM_layers) + list(model.deep_layers):
                if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.Linear)):
                    net += [m.weight, m.bias]
            net.append(model.prediction.weight)
            return (model.embed_user.weight.detach().numpy().copy(), model.embed_item.weight.detach().numpy().copy(),
                    np.concatenate([model.u_bias.weight.detach().numpy().ravel(), model.i_bias.weight.detach().numpy().ravel(),
                                    model.bias_.detach().numpy().ravel()]).astype(np.float32),
                    np.concatenate([w.detach().numpy().ravel() for w in net]).astype(np.float32),
                    np.concatenate([np.concatenate([m.running_mean.numpy(), m.running_var.numpy()]) for m in bns]).astype(np.float32)
                    if bns else np.zeros(0, np.float32))
        snaps, batches, losses = [snap()], [], []
        model.train()
        for step in range(3):
            b = np.stack([rng.integers(U, size=B), rng.integers(I, size=B), rng.integers(I, size=B)]).astype(np.int32)
            batches.append(b)
            model.zero_grad()
            loss = model.calc_loss([torch.from_numpy(b[0]), torch.from_numpy(b[1]), torch.from_numpy(b[2])])
            loss.backward()
            optim.step()
            losses.append(float(loss.item()))
            snaps.append(snap())
        model.eval()
        users = rng.permutation(U)[:9].astype(np.int64)
        cands = rng.integers(I, size=(9, 40)).astype(np.int64)
        loader = get_dataloader(CandidatesDataset([[int(u), c] for u, c in zip(users, cands)]), batch_size=128,
                                shuffle=False, num_workers=0)
        with torch.no_grad():
            preds = model.rank(loader)
            full = np.stack([model.full_rank(int(u)) for u in users[:4]])
            # predict() feeds a 1-D row to BatchNorm1d, which rejects it (:155-157): only callable without batch_norm
            pp = np.zeros(0, np.float32) if bn else \
                np.array([model.predict(int(users[q]), int(cands[q][0])) for q in range(4)], np.float32)
        out.update({f"c{k}_P": np.stack([s_[0] for s_ in snaps]), f"c{k}_Q": np.stack([s_[1] for s_ in snaps]),
                    f"c{k}_bias": np.stack([s_[2] for s_ in snaps]), f"c{k}_N": np.stack([s_[3] for s_ in snaps]),
                    f"c{k}_R": np.stack([s_[4] for s_ in snaps]), f"c{k}_batches": np.stack(batches),
                    f"c{k}_loss": np.array(losses, np.float64),
                    f"c{k}_hyper": np.array([L, 1 if bn else 0, ACT[act], lr, r1, r2, 0 if opt == "sgd" else 1], np.float64),
                    f"c{k}_users": users, f"c{k}_cands": cands.astype(np.int32), f"c{k}_preds": preds, f"c{k}_full": full,
                    f"c{k}_pred_pairs": pp})
        print(f"nfm case {k} (L={L} bn={bn} {act}/{opt}): losses {losses}")
    torch.nn.Dropout.forward = _drop_fwd
    out["ncases"] = np.array(len(cases))
    _save("nfm", **out)


def gen_nfm_dropout():
    """NFM at the reference's DEFAULT dropout (assets/nfm.yaml: 0.5; NFMRecommender.py:67,:88): the masks come from torch's
    global CPU generator, one bernoulli_ per Dropout call -- forward(user, pos): FM_layers' Dropout, then the Dropout behind
    each activation; forward(user, neg): the same.  With dropout > 0 the unmodified reference trains (no aliasing shim needed).
    Per case: snapshots around 3 training steps, the RNG seed set right before them, losses, eval-mode ranking."""
    import torch
    from daisy.model.NFMRecommender import NFM
    from daisy.utils.dataset import CandidatesDataset, get_dataloader
    out = {}
    cases = [  # U, I, F, L, bn, act, B, lr, reg1, reg2, opt, dropout, seed
        (40, 60, 30, 2, True, "relu", 128, 0.001, 0.0, 0.0, "sgd", 0.5, 61),          # assets/nfm.yaml
        (30, 40, 12, 1, False, "tanh", 50, 0.01, 0.001, 0.001, "sgd", 0.3, 62),
        (30, 45, 16, 3, True, "sigmoid", 96, 0.01, 0.0, 0.001, "sgd", 0.5, 63),
        (25, 30, 8, 2, False, "relu", 40, 0.001, 0.001, 0.0, "adam", 0.2, 64),
        (25, 30, 8, 0, True, "relu", 40, 0.01, 0.0, 0.0, "sgd", 0.5, 65),              # no hidden layer: only FM_layers' Dropout
    ]
    ACT = {"relu": 0, "sigmoid": 1, "tanh": 2}
    for k, (U, I, F, L, bn, act, B, lr, r1, r2, opt, drop, seed) in enumerate(cases):
        cfg = rh.make_config("nfm", user_num=U, item_num=I, factors=F, num_layers=L, batch_norm=bn, act_function=act, dropout=drop,
                             lr=lr, reg_1=r1, reg_2=r2, optimizer=opt, epochs=1, topk=10, cand_num=40)
        torch.manual_seed(seed)
        model = NFM(cfg)
        with torch.no_grad():
            model.embed_user.weight.mul_(3.0)
            model.embed_item.weight.mul_(3.0)
            model.u_bias.weight.normal_(0, 0.3)
            model.i_bias.weight.normal_(0, 0.3)
            model.bias_.fill_(0.25)
        model.criterion = model._build_criterion(model.loss_type)
        optim = model._build_optimizer(optimizer=model.optimizer, lr=model.lr)
        rng = np.random.default_rng(seed)
        bns = [m for m in list(model.FM_layers) + list(model.deep_layers) if isinstance(m, torch.nn.BatchNorm1d)]

        def snap():
            net = []
            for m in list(model.FM_layers) + list(model.deep_layers):
                if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.Linear)):
                    net += [m.weight, m.bias]
            net.append(model.prediction.weight)
            return (model.embed_user.weight.detach().numpy().copy(), model.embed_item.weight.detach().numpy().copy(),
                    np.concatenate([model.u_bias.weight.detach().numpy().ravel(), model.i_bias.weight.detach().numpy().ravel(),
                                    model.bias_.detach().numpy().ravel()]).astype(np.float32),
                    np.concatenate([w.detach().numpy().ravel() for w in net]).astype(np.float32),
                    np.concatenate([np.concatenate([m.running_mean.numpy(), m.running_var.numpy()]) for m in bns]).astype(np.float32)
                    if bns else np.zeros(0, np.float32))
        snaps, batches, losses = [snap()], [], []
        model.train()
        torch.manual_seed(seed + 100)                                   # the dropout masks of the 3 steps come from here
        for step in range(3):
            b = np.stack([rng.integers(U, size=B), rng.integers(I, size=B), rng.integers(I, size=B)]).astype(np.int32)
            batches.append(b)
            model.zero_grad()
            loss = model.calc_loss([torch.from_numpy(b[0]), torch.from_numpy(b[1]), torch.from_numpy(b[2])])
            loss.backward()
            optim.step()
            losses.append(float(loss.item()))
            snaps.append(snap())
        rng_after = torch.get_rng_state().numpy().copy()
        model.eval()
        users = rng.permutation(U)[:9].astype(np.int64)
        cands = rng.integers(I, size=(9, 40)).astype(np.int64)
        loader = get_dataloader(CandidatesDataset([[int(u), c] for u, c in zip(users, cands)]), batch_size=128,
                                shuffle=False, num_workers=0)
        with torch.no_grad():
            preds = model.rank(loader)
        out.update({f"c{k}_P": np.stack([s_[0] for s_ in snaps]), f"c{k}_Q": np.stack([s_[1] for s_ in snaps]),
                    f"c{k}_bias": np.stack([s_[2] for s_ in snaps]), f"c{k}_N": np.stack([s_[3] for s_ in snaps]),
                    f"c{k}_R": np.stack([s_[4] for s_ in snaps]), f"c{k}_batches": np.stack(batches),
                    f"c{k}_loss": np.array(losses, np.float64),
                    f"c{k}_hyper": np.array([L, 1 if bn else 0, ACT[act], lr, r1, r2, 0 if opt == "sgd" else 1, drop, seed], np.float64),
                    f"c{k}_users": users, f"c{k}_cands": cands.astype(np.int32), f"c{k}_preds": preds,
                    f"c{k}_rng_after": rng_after})
        print(f"nfm_dropout case {k} (L={L} bn={bn} {act}/{opt} p={drop}): losses {losses}")
    out["ncases"] = np.array(len(cases))
    _save("nfm_dropout", **out)


# --------------------------------------------------------------------------- NeuMF
def _neumf_flat(model):
    """4 tables + the flat tower block in module-registration order (layer weight, bias, ..., predict weight, bias)."""
    import torch.nn as nn
    tabs = [t.weight.detach().numpy().copy() for t in (model.embed_user_GMF, model.embed_item_GMF,
                                                        model.embed_user_MLP, model.embed_item_MLP)]
    parts = []
    for mod in model.MLP_layers:
        if isinstance(mod, nn.Linear):
            parts += [mod.weight.detach().numpy().ravel(), mod.bias.detach().numpy().ravel()]
    parts += [model.predict_layer.weight.detach().numpy().ravel(), model.predict_layer.bias.detach().numpy().ravel()]
    return tabs, np.concatenate(parts).astype(np.float32)


def gen_neumf():
    """NeuMF (NeuMFRecommender.py:16-232), model_name='NeuMF', dropout=0 (dropout>0 draws torch-RNG masks that no
    other implementation can reproduce): init stream, 3 training steps (Adam default / SGD, reg with the :158/:160
    quirk), rank / full_rank / predict."""
    import torch
    from daisy.model.NeuMFRecommender import NeuMF
    from daisy.utils.dataset import CandidatesDataset, get_dataloader
    out = {}
    cases = [  # U, I, F, L, B, lr, reg1, reg2, opt, seed
        (40, 60, 8, 2, 64, 0.001, 0.001, 0.001, "default", 31),
        (30, 50, 32, 2, 128, 0.001, 0.0, 0.0, "default", 32),          # BASELINE config 3 tower 128->64->32
        (25, 35, 24, 2, 100, 0.01, 0.002, 0.003, "sgd", 33),           # reference default factors=24
        (20, 30, 16, 3, 96, 0.001, 0.001, 0.001, "default", 34),
        (15, 20, 6, 1, 40, 0.01, 0.001, 0.0, "sgd", 35),
    ]
    for k, (U, I, F, L, B, lr, r1, r2, opt, seed) in enumerate(cases):
        cfg = rh.make_config("neumf", user_num=U, item_num=I, factors=F, num_layers=L, lr=lr, reg_1=r1, reg_2=r2,
                             optimizer=opt, dropout=0.0, epochs=1, topk=10, cand_num=40)
        torch.manual_seed(seed)
        model = NeuMF(cfg)
        with torch.no_grad():                                        # spread the scores so ranking is well separated
            for t in (model.embed_user_GMF, model.embed_item_GMF, model.embed_user_MLP, model.embed_item_MLP):
                t.weight.mul_(3.0)
        tabs0_init, _ = _neumf_flat(model)
        model.train()
        model.criterion = model._build_criterion(model.loss_type)
        optim = model._build_optimizer(optimizer=model.optimizer, lr=model.lr)
        rng = np.random.default_rng(seed)
        snaps = [_neumf_flat(model)]
        batches, losses = [], []
        for step in range(3):
            b = np.stack([rng.integers(U, size=B), rng.integers(I, size=B), rng.integers(I, size=B)]).astype(np.int32)
            batches.append(b)
            model.zero_grad()
            loss = model.calc_loss([torch.from_numpy(b[0]).long(), torch.from_numpy(b[1]).long(), torch.from_numpy(b[2]).long()])
            loss.backward()
            optim.step()
            losses.append(float(loss.item()))
            snaps.append(_neumf_flat(model))
        model.eval()
        users = rng.permutation(U)[:7].astype(np.int64)
        cands = rng.integers(I, size=(7, 40)).astype(np.int64)
        loader = get_dataloader(CandidatesDataset([[int(u), c] for u, c in zip(users, cands)]), batch_size=128,
                                shuffle=False, num_workers=0)
        with torch.no_grad():
            preds = model.rank(loader)
            full = np.stack([model.full_rank(int(u)) for u in users[:3]])
            pp = np.array([model.predict(int(users[q]), int(cands[q][0])) for q in range(4)], np.float32)
        for q, name in enumerate(("UG", "IG", "UM", "IM")):
            out[f"c{k}_{name}"] = np.stack([s_[0][q] for s_ in snaps])
        out.update({f"c{k}_W": np.stack([s_[1] for s_ in snaps]), f"c{k}_batches": np.stack(batches),
                    f"c{k}_loss": np.array(losses, np.float64),
                    f"c{k}_hyper": np.array([U, I, F, L, lr, r1, r2, 0 if opt == "sgd" else 1, seed], np.float64),
                    f"c{k}_users": users, f"c{k}_cands": cands.astype(np.int32), f"c{k}_preds": preds, f"c{k}_full": full,
                    f"c{k}_pred_pairs": pp})
        print(f"neumf case {k}: losses {losses}")
    out["ncases"] = np.array(len(cases))
    _save("neumf", **out)


def gen_neumf_modes():
    """The rest of NeuMF's surface (NeuMFRecommender.py:48-50,61,97-137): the reference's DEFAULT dropout (0.5, masks from
    torch's global CPU generator: one bernoulli_ per Dropout call, pos forward then neg forward) and model_name 'GMF' / 'MLP'
    / 'NeuMF-pre' (pre-trained GMF + MLP models, the bias-into-weight line :116 included).  Per case: init snapshot, the
    RNG seed set right before the 3 training steps, batches, losses, snapshots, eval-mode rank / full_rank / predict."""
    import torch
    from daisy.model.NeuMFRecommender import NeuMF
    from daisy.utils.dataset import CandidatesDataset, get_dataloader
    out = {}
    cases = [  # name, U, I, F, L, B, lr, reg1, reg2, opt, dropout, seed
        ("NeuMF", 40, 60, 24, 2, 64, 0.001, 0.001, 0.001, "default", 0.5, 41),     # assets/neumf.yaml defaults
        ("NeuMF", 30, 45, 8, 3, 48, 0.01, 0.002, 0.001, "sgd", 0.3, 42),
        ("GMF", 35, 50, 16, 2, 64, 0.001, 0.001, 0.001, "default", 0.5, 43),
        ("MLP", 35, 50, 16, 2, 64, 0.001, 0.001, 0.001, "default", 0.5, 44),
        ("NeuMF-pre", 35, 50, 16, 2, 64, 0.001, 0.001, 0.001, "default", 0.0, 45),   # from the two trained models above
    ]
    trained = {}
    for k, (name, U, I, F, L, B, lr, r1, r2, opt, drop, seed) in enumerate(cases):
        cfg = rh.make_config("neumf", user_num=U, item_num=I, factors=F, num_layers=L, lr=lr, reg_1=r1, reg_2=r2,
                             optimizer=opt, dropout=drop, epochs=1, topk=10, cand_num=40, model_name=name,
                             GMF_model=trained.get("GMF"), MLP_model=trained.get("MLP"))
        torch.manual_seed(seed)
        model = NeuMF(cfg)
        if name != "NeuMF-pre":
            with torch.no_grad():
                for t in (model.embed_user_GMF, model.embed_item_GMF, model.embed_user_MLP, model.embed_item_MLP):
                    t.weight.mul_(3.0)
        model.train()
        model.criterion = model._build_criterion(model.loss_type)
        optim = model._build_optimizer(optimizer=model.optimizer, lr=model.lr)
        rng = np.random.default_rng(seed)
        snaps = [_neumf_flat(model)]
        batches, losses = [], []
        torch.manual_seed(seed + 100)                                   # the dropout masks of the 3 steps come from here
        for step in range(3):
            b = np.stack([rng.integers(U, size=B), rng.integers(I, size=B), rng.integers(I, size=B)]).astype(np.int32)
            batches.append(b)
            model.zero_grad()
            loss = model.calc_loss([torch.from_numpy(b[0]).long(), torch.from_numpy(b[1]).long(), torch.from_numpy(b[2]).long()])
            loss.backward()
            optim.step()
            losses.append(float(loss.item()))
            snaps.append(_neumf_flat(model))
        rng_after = torch.get_rng_state().numpy().copy()
        model.eval()
        trained[name] = model
        users = rng.permutation(U)[:7].astype(np.int64)
        cands = rng.integers(I, size=(7, 40)).astype(np.int64)
        loader = get_dataloader(CandidatesDataset([[int(u), c] for u, c in zip(users, cands)]), batch_size=128,
                                shuffle=False, num_workers=0)
        with torch.no_grad():
            preds = model.rank(loader)
            full = np.stack([model.full_rank(int(u)) for u in users[:3]])
            pp = np.array([model.predict(int(users[q]), int(cands[q][0])) for q in range(4)], np.float32)
        for q, tn in enumerate(("UG", "IG", "UM", "IM")):
            out[f"c{k}_{tn}"] = np.stack([s_[0][q] for s_ in snaps])
        out.update({f"c{k}_W": np.stack([s_[1] for s_ in snaps]), f"c{k}_batches": np.stack(batches),
                    f"c{k}_loss": np.array(losses, np.float64), f"c{k}_name": np.array(name),
                    f"c{k}_hyper": np.array([U, I, F, L, lr, r1, r2, 0 if opt == "sgd" else 1, seed, drop], np.float64),
                    f"c{k}_users": users, f"c{k}_cands": cands.astype(np.int32), f"c{k}_preds": preds, f"c{k}_full": full,
                    f"c{k}_pred_pairs": pp, f"c{k}_rng_after": rng_after})
        print(f"neumf_modes case {k} ({name}, dropout {drop}): losses {losses}")
    out["ncases"] = np.array(len(cases))
    _save("neumf_modes", **out)


# --------------------------------------------------------------------------- evaluation KPIs
def gen_metrics():
    """daisy/utils/metrics.py:18-57,59-96,98-251 (calc_ranking_results / Metric.run) on synthetic rank lists:
    duplicate ids, users without hits, users whose whole list hits, topk inside and outside common_ks."""
    import logging
    import tempfile
    from daisy.utils.metrics import calc_ranking_results, Metric
    rng = np.random.default_rng(11)
    out = {}
    cases = [(200, 500, 50), (64, 90, 7), (33, 40, 20), (1, 30, 10)]
    for c, (n, I, topk) in enumerate(cases):
        test_u = rng.permutation(n * 3)[:n].astype(np.int64)             # user ids, arbitrary order
        test_ur, preds = {}, np.zeros((n, topk), np.float32)
        for r, u in enumerate(test_u):
            g = int(rng.integers(1, min(I, 40)))
            gt = rng.choice(I, size=g, replace=False)
            test_ur[int(u)] = set(int(x) for x in gt)
            row = rng.integers(0, I, size=topk)                            # duplicates happen (with replacement)
            mode = r % 5
            if mode == 0:                                                  # no hit at all
                pool = np.setdiff1d(np.arange(I), gt)
                row = rng.choice(pool, size=topk) if len(pool) else row
            elif mode == 1:                                                # every position hits (duplicates of gt)
                row = rng.choice(gt, size=topk)
            elif mode == 2:                                                # one planted hit, repeated twice
                pool = np.setdiff1d(np.arange(I), gt)
                if len(pool):
                    row = rng.choice(pool, size=topk)
                pos = int(rng.integers(topk))
                row[pos] = gt[0]
                row[min(topk - 1, pos + 2)] = gt[0]
            preds[r] = row
        item_pop = rng.integers(0, 1000, size=I).astype(np.float64)
        names = ["recall", "mrr", "ndcg", "hit", "precision", "coverage"]
        cfg = {"logger": logging.getLogger("gold"), "res_path": tempfile.mkdtemp() + "/", "metrics": names,
               "item_num": I, "topk": topk, "item_pop": item_pop}
        res = calc_ranking_results(test_ur, preds, list(test_u), cfg)
        ks = np.array([int(k) for k in res.columns[1:]], np.int32)
        cfg2 = dict(cfg, metrics=names + ["popularity"])                  # Popularity indexes item_pop: integer preds only
        res2 = calc_ranking_results(test_ur, preds.astype(np.int64), list(test_u), cfg2)
        # 'map' is accepted by Metric.run (:84) but has no display name (:5-16), so calc_ranking_results raises
        # KeyError on it; pin it through Metric.run with the same per-cutoff slicing (:49-50)
        mrun = Metric(dict(cfg, metrics=["map"]))
        kpi_map = np.array([mrun.run(test_ur, preds[:, :int(k)], list(test_u))[0] for k in ks], np.float64)
        gt_len = np.array([len(test_ur[int(u)]) for u in test_u], np.int32)
        gt_flat = np.concatenate([np.array(sorted(test_ur[int(u)]), np.int32) for u in test_u])
        out.update({f"c{c}_preds": preds, f"c{c}_test_u": test_u, f"c{c}_gt_flat": gt_flat, f"c{c}_gt_len": gt_len,
                    f"c{c}_ks": ks, f"c{c}_item_pop": item_pop, f"c{c}_meta": np.array([n, I, topk], np.int64),
                    f"c{c}_kpi": res.values[:, 1:].astype(np.float64),     # [6 metrics, len(ks)], rows = kpi_names
                    f"c{c}_kpi_pop": res2.values[:, 1:].astype(np.float64),   # [7 metrics, len(ks)] (+ popularity)
                    f"c{c}_kpi_map": kpi_map})                                 # [len(ks)]
        print(res2)
    out["ncases"] = np.array(len(cases))
    out["kpi_names"] = np.array(["recall", "mrr", "ndcg", "hit", "precision", "coverage", "popularity"])
    _save("metrics", **out)


# --------------------------------------------------------------------------- popularity-mixed / point-wise sampler
def gen_sampler_pop():
    """daisy/utils/sampler.py:43-53,64-81 ('low-pop' / 'high-pop' mix) and :93-98 (point-wise explode, CL / SL)."""
    import pandas as pd
    from daisy.utils.sampler import BasicNegtiveSampler
    rng = np.random.default_rng(19)
    out = {}
    cases = [(40, 70, 500, 4, 2022, "high-pop", 0.5, "BPR"), (25, 300, 400, 5, 7, "low-pop", 0.3, "BPR"),
             (30, 64, 350, 3, 99, "high-pop", 1.0, "BPR"), (30, 64, 350, 4, 5, "low-pop", 0.1, "BPR"),
             (20, 50, 200, 2, 3, "uniform", 0.0, "CL"), (20, 50, 200, 3, 4, "high-pop", 0.4, "SL")]
    for c, (U, I, nnz, G, seed, method, ratio, loss) in enumerate(cases):
        cu, ci = _synthetic_inter(rng, U, I, nnz, None)
        rating = rng.integers(1, 6, size=len(cu)).astype(np.float64) if loss == "SL" else np.ones(len(cu))
        df = pd.DataFrame({"user": cu, "item": ci, "rating": rating, "timestamp": np.arange(len(cu))})
        cfg = rh.make_config("mf", user_num=U, item_num=I, num_ng=G, train_ur=_ur_from(cu, ci), sample_method=method,
                             sample_ratio=ratio, loss_type=loss)
        np.random.seed(seed)
        smp = BasicNegtiveSampler(df, cfg)
        rows = smp.sampling()
        nxt = np.random.randint(0, 2 ** 31 - 1, size=3)
        out.update({f"c{c}_coo_u": cu, f"c{c}_coo_i": ci, f"c{c}_rating": rating, f"c{c}_rows": rows, f"c{c}_next": nxt,
                    f"c{c}_meta": np.array([U, I, G, seed], np.int64), f"c{c}_ratio": np.array(ratio),
                    f"c{c}_method": np.array(method), f"c{c}_loss": np.array(loss),
                    f"c{c}_pop_prob": np.zeros(0) if smp.pop_prob is None else np.asarray(smp.pop_prob, np.float64)})
        print(f"sampler_pop case {c}: rows {rows.shape}, head {rows[:2].tolist()}")
    out["ncases"] = np.array(len(cases))
    _save("sampler_pop", **out)


ALL = {"ngcf_dropout": gen_ngcf_dropout, "nfm_dropout": gen_nfm_dropout, "neumf_modes": gen_neumf_modes, "nfm": gen_nfm, "ngcf": gen_ngcf, "fm": gen_fm, "mf_optim": gen_mf_optim, "mf_pointwise": gen_mf_pointwise, "metrics": gen_metrics, "sampler_pop": gen_sampler_pop, "neumf": gen_neumf, "lightgcn": gen_lightgcn, "sampler_small": gen_sampler_small, "ml100k": gen_ml100k_pipeline, "mf_steps": gen_mf_steps,
       "mf_rank": gen_mf_rank}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    rh.import_reference()
    for name, fn in ALL.items():
        if a.only and name not in a.only.split(","):
            continue
        print(f"== {name}")
        fn()
