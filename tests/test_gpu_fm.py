"""FM on the device (daisy/model/FMRecommender.py:16-131) against the reference-generated
fixture tests/golden/fm.npz (5 cases: BPR / CL / TL / HL, SGD / Adam) and the pinned oracle, through the C ABI."""
import logging

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


def _cases():
    g = golden("fm")
    return g, range(int(g["ncases"]))


def test_fm_steps_match_reference_fixture(orc):
    from daisyrec_b200 import ops
    g, cases = _cases()
    for c in cases:
        lr, r1, r2 = (float(x) for x in g[f"c{c}_hyper"])
        opt, loss = str(g[f"c{c}_opt"]), str(g[f"c{c}_losskind"])
        Ps, Qs, Bs, bs, losses = g[f"c{c}_P"], g[f"c{c}_Q"], g[f"c{c}_bias"], g[f"c{c}_batches"], g[f"c{c}_loss"]
        U, F = Ps.shape[1:]
        I = Qs.shape[1]
        hp = ops.hyper(lr, r1, r2, opt, loss=loss)
        ws = ops.FMWorkspace(U, I, F, opt, "cuda")                         # optimiser state carried across the 3 steps
        ohp = orc.hyper(lr=lr, reg_1=r1, reg_2=r2, opt=opt, loss=loss)
        ostate = None if opt == "sgd" else tuple(np.zeros_like(a) for a in (Ps[0], Ps[0], Qs[0], Qs[0]))
        obstate = None if opt == "sgd" else np.zeros(2 * Bs.shape[1], np.float32)
        for s in range(bs.shape[0]):
            P, Q, bias = (torch.from_numpy(a[s].copy()).cuda() for a in (Ps, Qs, Bs))
            b = [torch.from_numpy(np.ascontiguousarray(bs[s][k])).cuda() for k in range(3)]
            # loss only (no update), then the step
            l0 = ops.fm_train_steps(P, Q, bias, ws, *b, b[0].numel(), 0, 1, hp, adam_step0=s, apply=False).item()
            assert np.array_equal(P.cpu().numpy(), Ps[s]) and np.array_equal(bias.cpu().numpy(), Bs[s])
            l1 = ops.fm_train_steps(P, Q, bias, ws, *b, b[0].numel(), 0, 1, hp, adam_step0=s).item()
            assert l0 == l1
            assert abs(l1 - losses[s]) <= 1e-5 * abs(losses[s]), (c, s, l1, losses[s])
            Po, Qo, bo = Ps[s].copy(), Qs[s].copy(), Bs[s].copy()
            lo, _ = orc.fm_step(Po, Qo, bo, *(np.ascontiguousarray(bs[s][k]) for k in range(3)), ohp, True, ostate, obstate,
                                s + 1)
            assert abs(l1 - lo) <= 2e-6 * abs(lo), (c, s, l1, lo)
            tol = 3e-6 if opt == "sgd" else 3e-5
            for got, want, orc_ in ((P, Ps[s + 1], Po), (Q, Qs[s + 1], Qo), (bias, Bs[s + 1], bo)):
                got = got.cpu().numpy()
                scale = max(1.0, np.abs(want).max())
                if opt == "sgd":
                    np.testing.assert_allclose(got, want, rtol=0, atol=tol * scale)
                    np.testing.assert_allclose(got, orc_, rtol=0, atol=tol * scale)
                else:
                    # Adam turns a gradient that is pure fp32 cancellation noise into a +-lr step of implementation-
                    # defined sign: bounded by 2*lr, on a small fraction of the elements
                    bad = np.abs(got - want) > tol * scale
                    assert bad.mean() <= 0.01 and np.abs(got - want).max() <= 2.1 * lr, (c, s, bad.mean())


def test_fm_rank_full_rank_predict(orc):
    from daisyrec_b200 import ops
    g, cases = _cases()
    for c in cases:
        P, Q, bias = g[f"c{c}_P"][-1], g[f"c{c}_Q"][-1], g[f"c{c}_bias"][-1]
        users, cands = g[f"c{c}_users"], g[f"c{c}_cands"].astype(np.int64)
        K = g[f"c{c}_preds"].shape[1]
        dP, dQ, db = (torch.from_numpy(a.copy()).cuda() for a in (P, Q, bias))
        got = ops.fm_rank(dP, dQ, db, torch.from_numpy(users).cuda(), torch.from_numpy(cands).cuda(), K).cpu().numpy()
        want = np.stack([cands[r][np.argsort(-orc.fm_scores(P, Q, bias, u, cands[r]), kind="stable")[:K]]
                         for r, u in enumerate(users)]).astype(np.float32)
        assert np.array_equal(got, want), c                                  # bit-exact vs the oracle's canonical scores
        assert (got == g[f"c{c}_preds"]).mean() >= 0.98, c                   # vs the reference's bmm summation order
        nf, Kf = g[f"c{c}_full"].shape                                       # Kf = min(topk, item_num)
        full = ops.fm_full_rank(dP, dQ, db, torch.from_numpy(users[:nf]).cuda(), Kf).cpu().numpy()
        wfull = np.stack([np.argsort(-orc.fm_scores(P, Q, bias, u), kind="stable")[:Kf] for u in users[:nf]])
        assert np.array_equal(full, wfull), c
        assert (full == g[f"c{c}_full"]).mean() >= 0.9, c
        pp = ops.fm_predict(dP, dQ, db, torch.from_numpy(users[:4].astype(np.int32)).cuda(),
                            torch.from_numpy(cands[:4, 0].astype(np.int32)).cuda()).cpu().numpy()
        wpp = np.array([orc.fm_scores(P, Q, bias, users[q], cands[q][:1])[0] for q in range(4)], np.float32)
        assert np.array_equal(pp, wpp)
        np.testing.assert_allclose(pp, g[f"c{c}_pred_pairs"], rtol=2e-6, atol=2e-6)


def test_fm_class_drop_in():
    """FM(config): the reference's constructor RNG stream (bit-identical initial tables, zero biases), calc_loss / fit /
    rank through the class surface."""
    from daisyrec_b200.model import FM
    from daisyrec_b200.utils.dataset import BasicDataset, CandidatesDataset, get_dataloader
    g = golden("fm")
    seeds = [31, 32, 33, 34, 35]
    for c in (0, 3):
        lr, r1, r2 = (float(x) for x in g[f"c{c}_hyper"])
        U, F = g[f"c{c}_P"].shape[1:]
        I = g[f"c{c}_Q"].shape[1]
        cfg = dict(gpu="", logger=logging.getLogger("t"), epochs=1, lr=lr, reg_1=r1, reg_2=r2, user_num=U, item_num=I,
                   factors=F, loss_type=str(g[f"c{c}_losskind"]), optimizer=str(g[f"c{c}_opt"]), init_method="default",
                   early_stop=False, topk=10, progress=False)
        torch.manual_seed(seeds[c])
        m = FM(cfg)
        assert np.array_equal(m.embed_user.weight.cpu().numpy(), g[f"c{c}_P_init"])
        assert np.array_equal(m.embed_item.weight.cpu().numpy(), g[f"c{c}_Q_init"])
        assert float(m.bias.abs().sum().item()) == 0.0 and m.u_bias.weight.shape == (U, 1) and m.bias_.shape == (1,)
        Bs = g[f"c{c}_bias"]
        m.load_state_dict({"embed_user.weight": g[f"c{c}_P"][0], "embed_item.weight": g[f"c{c}_Q"][0],
                           "u_bias.weight": Bs[0][:U], "i_bias.weight": Bs[0][U:U + I], "bias_": Bs[0][U + I:]})
        b = g[f"c{c}_batches"]
        loss = m.calc_loss([torch.from_numpy(b[0][k]) for k in range(3)])
        assert loss.dtype == torch.float32 and abs(loss.item() - g[f"c{c}_loss"][0]) <= 1e-5 * abs(g[f"c{c}_loss"][0])
        # fit over the three fixture batches in order == the reference's three steps
        rows = np.ascontiguousarray(np.concatenate([b[s].T for s in range(3)]).astype(np.int32))
        m.fit(get_dataloader(BasicDataset(rows), batch_size=b.shape[2], shuffle=False))
        tol = 3e-6 if str(g[f"c{c}_opt"]) == "sgd" else 3e-5
        got = m.embed_user.weight.cpu().numpy()
        want = g[f"c{c}_P"][3]
        assert (np.abs(got - want) > tol * max(1.0, np.abs(want).max())).mean() <= 0.01
        users, cands = g[f"c{c}_users"], g[f"c{c}_cands"].astype(np.int64)
        loader = get_dataloader(CandidatesDataset([[int(u), cands[r]] for r, u in enumerate(users)]), batch_size=128,
                                shuffle=False)
        preds = m.rank(loader)
        assert preds.dtype == np.float32 and preds.shape == g[f"c{c}_preds"].shape
        assert (preds == g[f"c{c}_preds"]).mean() >= 0.97
        assert m.full_rank(int(users[0])).shape == (min(10, I),)
        assert isinstance(m.predict(int(users[0]), int(cands[0][0])), float)
