"""NGCF on the device (daisy/model/NGCFRecommender.py:38-252) against the reference-generated fixture
tests/golden/ngcf.npz (4 cases: 1-3 BiGNN layers of unequal width, Adam / SGD, regulariser on / off) and the pinned oracle."""
import logging

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
SEEDS = [41, 42, 43, 44]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _case(g, c):
    U, I, lr, r1, r2, opt = g[f"c{c}_hyper"]
    return int(U), int(I), float(lr), float(r1), float(r2), ("sgd" if opt == 0 else "adam"), [int(d) for d in g[f"c{c}_dims"]]


def test_ngcf_forward_and_steps_match_reference_fixture(orc):
    from daisyrec_b200 import ops
    g = golden("ngcf")
    for c in range(int(g["ncases"])):
        U, I, lr, r1, r2, optn, dims = _case(g, c)
        row_ptr, col, val = ops.lgcn_norm_adj(g[f"c{c}_coo_u"], g[f"c{c}_coo_i"], U, I)
        graph = ops.LgcnGraph(row_ptr, col, val, "cuda")
        Es, Ws, bs, losses = g[f"c{c}_E"], g[f"c{c}_W"], g[f"c{c}_batches"], g[f"c{c}_loss"]
        assert Ws.shape[1] == ops.ngcf_param_count(dims)
        ws = ops.NgcfWorkspace(U, I, dims, optn, "cuda")                  # optimiser state carried across the 3 steps
        rep = ops.ngcf_forward(dev(Es[0]), dev(Ws[0]), ws, graph).cpu().numpy()
        np.testing.assert_allclose(rep, g[f"c{c}_all0"], rtol=0, atol=3e-6, err_msg=f"case {c} forward")
        want0 = orc.ngcf_forward(Es[0].copy(), Ws[0].copy(), U, I, np.asarray(dims, np.int32), row_ptr, col, val)
        np.testing.assert_allclose(rep, want0, rtol=0, atol=3e-6)
        hp = ops.hyper(lr, r1, r2, optn)
        for s in range(bs.shape[0]):
            E, W = dev(Es[s]), dev(Ws[s])
            b = [dev(bs[s][k]) for k in range(3)]
            l0 = ops.ngcf_bpr_train_steps(E, W, ws, graph, *b, b[0].numel(), 0, 1, hp, adam_step0=s, apply=False).item()
            assert np.array_equal(E.cpu().numpy(), Es[s]) and np.array_equal(W.cpu().numpy(), Ws[s])
            l1 = ops.ngcf_bpr_train_steps(E, W, ws, graph, *b, b[0].numel(), 0, 1, hp, adam_step0=s).item()
            assert abs(l0 - losses[s]) <= 3e-5 * abs(losses[s]) and abs(l1 - losses[s]) <= 3e-5 * abs(losses[s]), (c, s, l0, l1)
            for got, want, nm in ((E.cpu().numpy(), Es[s + 1], "E"), (W.cpu().numpy(), Ws[s + 1], "W")):
                err = np.abs(got - want)
                tol = (5e-6 if optn == "sgd" else 5e-5) * max(1.0, np.abs(want).max())
                # Adam turns cancellation noise of a ~0 gradient into a +-lr step: a handful of elements, bounded by 2 lr
                assert (err <= tol).mean() >= 0.99 and err.max() <= 2.1 * lr + tol, (c, s, nm, float((err <= tol).mean()),
                                                                                   float(err.max()))


def test_ngcf_class_drop_in():
    """NGCF(config): the reference's constructor RNG stream (bit-identical tables and BiGNN weights), calc_loss, fit over the
    fixture batches, rank / full_rank / predict on the concatenated representation."""
    import pandas as pd
    from daisyrec_b200.model import NGCF
    from daisyrec_b200.utils.dataset import BasicDataset, CandidatesDataset, get_dataloader
    from daisyrec_b200.utils.utils import get_inter_matrix
    g = golden("ngcf")
    for c in range(int(g["ncases"])):
        U, I, lr, r1, r2, optn, dims = _case(g, c)
        cu, ci = g[f"c{c}_coo_u"], g[f"c{c}_coo_i"]
        df = pd.DataFrame({"user": cu, "item": ci, "rating": 1.0, "timestamp": np.arange(len(cu))})
        cfg = dict(gpu="", logger=logging.getLogger("t"), epochs=1, lr=lr, reg_1=r1, reg_2=r2, user_num=U, item_num=I,
                   factors=dims[0], hidden_size_list=dims[1:], node_dropout=0.0, mess_dropout=0.0, loss_type="BPR",
                   optimizer="sgd" if optn == "sgd" else "default", init_method="default", early_stop=False, topk=10,
                   progress=False, UID_NAME="user", IID_NAME="item", INTER_NAME="rating")
        cfg["inter_matrix"] = get_inter_matrix(df, cfg)
        torch.manual_seed(SEEDS[c])
        m = NGCF(cfg)
        assert np.array_equal(m.E0.cpu().numpy(), g[f"c{c}_E"][0]), c
        assert np.array_equal(m.gnn.cpu().numpy(), g[f"c{c}_W"][0]), c
        eu, ei = m.forward()
        np.testing.assert_allclose(torch.cat([eu, ei]).cpu().numpy(), g[f"c{c}_all0"], rtol=0, atol=3e-6)
        b = g[f"c{c}_batches"]
        loss = m.calc_loss([torch.from_numpy(b[0][k]) for k in range(3)])
        assert abs(loss.item() - g[f"c{c}_loss"][0]) <= 3e-5 * abs(g[f"c{c}_loss"][0])
        rows = np.ascontiguousarray(np.concatenate([b[s].T for s in range(3)]).astype(np.int32))
        m.fit(get_dataloader(BasicDataset(rows), batch_size=b.shape[2], shuffle=False))
        err = np.abs(m.E0.cpu().numpy() - g[f"c{c}_E"][3])
        tol = (2e-5 if optn == "sgd" else 2e-4) * max(1.0, np.abs(g[f"c{c}_E"][3]).max())
        assert (err <= tol).mean() >= 0.98, (c, float((err <= tol).mean()))
        m.load_state_dict({"embed_user.weight": g[f"c{c}_E"][3][:U], "embed_item.weight": g[f"c{c}_E"][3][U:],
                           "gnn": g[f"c{c}_W"][3]})
        users, cands = g[f"c{c}_users"], g[f"c{c}_cands"].astype(np.int64)
        loader = get_dataloader(CandidatesDataset([[int(u), cands[r]] for r, u in enumerate(users)]), batch_size=128,
                                shuffle=False)
        preds = m.rank(loader)
        assert preds.dtype == np.float32 and (preds == g[f"c{c}_preds"]).mean() >= 0.97, c
        full = np.stack([m.full_rank(int(u)) for u in users[:4]])
        assert (full == g[f"c{c}_full"]).mean() >= 0.9
        np.testing.assert_allclose(m.predict(int(users[0]), int(cands[0][0])), float(g[f"c{c}_pred_pair"][0]), rtol=3e-5, atol=3e-6)
        with pytest.raises(NotImplementedError):                          # sparse dropout of the adjacency (reference default 0)
            NGCF(dict(cfg, node_dropout=0.1))                              # (mess_dropout > 0 runs: tests/test_gpu_zzz_late.py)


def test_ngcf_vs_oracle_random(orc):
    """A larger random graph with the reference's default widths against the pinned oracle (Adam, regulariser on)."""
    from daisyrec_b200 import ops
    rng = np.random.default_rng(8)
    U, I, nnz, B = 300, 400, 5000, 2000
    dims = [64, 64, 64, 64]
    cu, ci = rng.integers(U, size=nnz), rng.integers(I, size=nnz)
    row_ptr, col, val = ops.lgcn_norm_adj(cu, ci, U, I)
    graph = ops.LgcnGraph(row_ptr, col, val, "cuda")
    E_h = (rng.standard_normal((U + I, dims[0])) * 0.1).astype(np.float32)
    W_h = (rng.standard_normal(ops.ngcf_param_count(dims)) * 0.1).astype(np.float32)
    b = [rng.integers(n, size=B).astype(np.int32) for n in (U, I, I)]
    E, W = dev(E_h), dev(W_h)
    ws = ops.NgcfWorkspace(U, I, dims, "adam", "cuda")
    loss = ops.ngcf_bpr_train_steps(E, W, ws, graph, *[dev(x) for x in b], B, 0, 1, ops.hyper(0.01, 0.001, 0.001, "adam")).item()
    state = np.zeros(2 * (E_h.size + W_h.size), np.float32)
    Eo, Wo = E_h.copy(), W_h.copy()
    lo = orc.ngcf_bpr_step(Eo, Wo, U, I, np.asarray(dims, np.int32), row_ptr, col, val, *b, orc.hyper(0.01, 0.001, 0.001, "adam"),
                           True, state, 1)
    assert abs(loss - lo) <= 3e-5 * abs(lo), (loss, lo)
    for got, want in ((E.cpu().numpy(), Eo), (W.cpu().numpy(), Wo)):
        err = np.abs(got - want)
        assert (err <= 2e-4).mean() >= 0.99 and err.max() <= 0.0211, (float((err <= 2e-4).mean()), float(err.max()))
