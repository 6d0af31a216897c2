"""GPU parity for NeuMF: golden fixtures from the reference + oracle on random inputs."""
import logging

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
NAMES = ("UG", "IG", "UM", "IM")


@pytest.fixture(scope="module")
def ops():
    from daisyrec_b200 import ops as o
    o.require_cuda()
    return o


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_neumf_golden(ops):
    g = golden("neumf")
    for c in range(int(g["ncases"])):
        U, I, F, L, lr, r1, r2, opt, seed = g[f"c{c}_hyper"]
        U, I, F, L = int(U), int(I), int(F), int(L)
        if F % 4:
            continue                                                  # 128-bit rows: factors multiple of 4
        optn = "sgd" if opt == 0 else "adam"
        hp = ops.hyper(lr, r1, r2, optn)
        bs, losses = g[f"c{c}_batches"], g[f"c{c}_loss"]
        tabs = [dev(g[f"c{c}_{n}"][0]) for n in NAMES]
        W = dev(g[f"c{c}_W"][0])
        ws = ops.NeumfWorkspace(U, I, F, L, optn, 4096, "cuda")
        for s in range(bs.shape[0]):
            b = [dev(bs[s][k]) for k in range(3)]
            l0 = ops.neumf_bpr_train_steps(tabs, W, ws, *b, b[0].numel(), 0, 1, hp, apply=False).item()
            loss = ops.neumf_bpr_train_steps(tabs, W, ws, *b, b[0].numel(), 0, 1, hp, adam_step0=s).item()
            assert abs(l0 - losses[s]) <= 3e-5 * abs(losses[s]) and abs(loss - losses[s]) <= 3e-5 * abs(losses[s]), (c, s)
            tol = (5e-6 if opt == 0 else 1e-4) * (s + 1)
            for q, n in enumerate(NAMES):
                want = g[f"c{c}_{n}"][s + 1]
                np.testing.assert_allclose(tabs[q].cpu().numpy(), want, rtol=0, atol=tol * max(1.0, np.abs(want).max()),
                                           err_msg=f"case {c} step {s} table {n}")
            np.testing.assert_allclose(W.cpu().numpy(), g[f"c{c}_W"][s + 1], rtol=0, atol=tol * max(1.0, np.abs(g[f"c{c}_W"][s + 1]).max()))
        # inference on the reference's final parameters: ids bit-exact, scores to fp32 rounding
        tabs = [dev(g[f"c{c}_{n}"][-1]) for n in NAMES]
        W = dev(g[f"c{c}_W"][-1])
        users, cands = g[f"c{c}_users"], g[f"c{c}_cands"].astype(np.int64)
        sc = ops.neumf_scores(tabs, W, ws, dev(users), dev(cands), cands.shape[1])
        assert np.array_equal(ops.topk_from_scores(sc, dev(cands), 10).cpu().numpy(), g[f"c{c}_preds"])
        scf = ops.neumf_scores(tabs, W, ws, dev(users[:3]), None, I)
        assert np.array_equal(ops.topk_from_scores(scf, None, 10).cpu().numpy(), g[f"c{c}_full"])
        np.testing.assert_allclose(sc[:4, 0].cpu().numpy(), g[f"c{c}_pred_pairs"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("F,L,opt,reg,B", [(32, 2, "adam", 0.001, 3000), (64, 1, "sgd", 0.0, 1000), (8, 3, "adam", 0.002, 777),
                                           (24, 2, "sgd", 0.001, 2049)])
def test_neumf_vs_oracle_random(ops, orc, F, L, opt, reg, B):
    rng = np.random.default_rng(F * 7 + L)
    U, I = 400, 300
    D = F * 2 ** (L - 1)
    shapes = [(U, F), (I, F), (U, D), (I, D)]
    tabs_h = [(rng.standard_normal(s) * 0.3).astype(np.float32) for s in shapes]
    nW = orc.neumf_param_count(F, L)
    W_h = (rng.standard_normal(nW) * 0.2).astype(np.float32)
    tabs_o, W_o = [t.copy() for t in tabs_h], W_h.copy()
    tabs, W = [dev(t) for t in tabs_h], dev(W_h)
    hp_d, hp_o = ops.hyper(0.01, reg, reg, opt), orc.hyper(0.01, reg, reg, opt)
    ws = ops.NeumfWorkspace(U, I, F, L, opt, 2 * B, "cuda")
    adam = None if opt == "sgd" else ([np.zeros_like(a) for a in tabs_o + [W_o]], [np.zeros_like(a) for a in tabs_o + [W_o]])
    for s in range(2):
        b = [rng.integers(U, size=B).astype(np.int32), np.minimum(I - 1, rng.zipf(1.3, size=B) - 1).astype(np.int32),
             rng.integers(I, size=B).astype(np.int32)]
        lo = orc.neumf_bpr_step(tabs_o, W_o, F, L, *b, hp_o, True, adam, s + 1)
        ld = ops.neumf_bpr_train_steps(tabs, W, ws, *[dev(x) for x in b], B, 0, 1, hp_d, adam_step0=s).item()
        assert abs(ld - lo) <= 1e-5 * abs(lo), (ld, lo)
        tol = 2e-5 if opt == "sgd" else 3e-4
        # Adam turns a gradient that is pure fp32 cancellation noise (dead ReLU units) into a +-lr step whose sign is
        # implementation-defined: allow <=0.2 % such elements, bounded by 2*lr per step.
        for got, want, t in [(tabs[q].cpu().numpy(), tabs_o[q], tol) for q in range(4)] + [(W.cpu().numpy(), W_o, tol * 5)]:
            err = np.abs(got - want)
            if opt == "sgd":
                assert err.max() <= t, (s, err.max())
            else:
                assert (err <= t).mean() >= 0.998 and err.max() <= 2.1 * 0.01 * (s + 1), (s, err.max(), int((err > t).sum()))
    # scores vs oracle
    users = rng.integers(U, size=20).astype(np.int64)
    cands = rng.integers(I, size=(20, 50)).astype(np.int64)
    sc = ops.neumf_scores(tabs, W, ws, dev(users), dev(cands), 50).cpu().numpy()
    # score the DEVICE-trained parameters with the oracle: isolates the inference path from the Adam sign noise above
    tabs_d, W_d = [t.cpu().numpy() for t in tabs], W.cpu().numpy()
    want = orc.neumf_predict(tabs_d, W_d, F, L, np.repeat(users, 50).astype(np.int32), cands.reshape(-1).astype(np.int32))
    np.testing.assert_allclose(sc.reshape(-1), want, rtol=0, atol=5e-6 * max(1.0, np.abs(want).max()))
    got_ids = ops.topk_from_scores(dev(sc), dev(cands), 10).cpu().numpy()
    assert (got_ids == orc.neumf_rank(tabs_d, W_d, F, L, users, cands, 10)).mean() > 0.99


def test_neumf_dropin_class(ops):
    from daisyrec_b200.model.NeuMFRecommender import NeuMF
    from daisyrec_b200.utils.dataset import BasicDataset, CandidatesDataset, get_dataloader
    g = golden("neumf")
    c = 1
    U, I, F, L, lr, r1, r2, opt, seed = g[f"c{c}_hyper"]
    U, I, F, L = int(U), int(I), int(F), int(L)
    cfg = dict(gpu='', logger=logging.getLogger('t'), lr=lr, epochs=1, reg_1=r1, reg_2=r2, dropout=0.0, model_name='NeuMF',
               GMF_model=None, MLP_model=None, user_num=U, item_num=I, factors=F, num_layers=L, loss_type='BPR',
               optimizer='default', init_method='default', early_stop=False, topk=10, progress=False)
    torch.manual_seed(int(seed))
    model = NeuMF(cfg)
    # same constructor RNG stream as the reference (the fixture multiplied the embeddings by 3 afterwards)
    np.testing.assert_array_equal(model.tower.cpu().numpy(), g[f"c{c}_W"][0])
    np.testing.assert_allclose(model.embed_user_GMF.weight.cpu().numpy() * 3.0, g[f"c{c}_UG"][0], rtol=1e-6)
    model.load_state_dict({'embed_user_GMF.weight': dev(g[f"c{c}_UG"][0]), 'embed_item_GMF.weight': dev(g[f"c{c}_IG"][0]),
                           'embed_user_MLP.weight': dev(g[f"c{c}_UM"][0]), 'embed_item_MLP.weight': dev(g[f"c{c}_IM"][0]),
                           'tower': dev(g[f"c{c}_W"][0])})
    bs = g[f"c{c}_batches"]
    data = np.ascontiguousarray(np.concatenate([bs[s].T for s in range(3)]))
    model.fit(get_dataloader(BasicDataset(data), batch_size=bs.shape[2], shuffle=False))
    np.testing.assert_allclose(model.embed_item_MLP.weight.cpu().numpy(), g[f"c{c}_IM"][3], rtol=0, atol=5e-4)
    users, cands = g[f"c{c}_users"], g[f"c{c}_cands"].astype(np.int64)
    loader = get_dataloader(CandidatesDataset([[int(u), cc] for u, cc in zip(users, cands)]), batch_size=128, shuffle=False)
    preds = model.rank(loader)
    assert preds.dtype == np.float32 and preds.shape == (7, 10)
    assert model.full_rank(int(users[0])).dtype == np.int64
    # the reference's default dropout=0.5 trains (device masks), and eval-mode inference ignores dropout
    torch.manual_seed(3)
    m2 = NeuMF(dict(cfg, dropout=0.5))
    m2.fit(get_dataloader(BasicDataset(data), batch_size=bs.shape[2], shuffle=False))
    assert np.isfinite(m2.tower.cpu().numpy()).all() and m2.rank(loader).shape == (7, 10)
    with pytest.raises(ValueError):
        NeuMF(dict(cfg, model_name='SomethingElse'))
    with pytest.raises(ValueError):
        NeuMF(dict(cfg, model_name='NeuMF-pre'))                      # needs config['GMF_model'] / config['MLP_model']


@pytest.mark.parametrize("tower_dtype", [0])
def test_neumf_dropout_gradient_consistency(ops, tower_dtype):
    """Dropout forward/backward consistency without access to torch's masks: with counter-based masks the loss at fixed
    (seed, step) is a deterministic function of the parameters, so the SGD update must equal -lr * its numerical gradient."""
    rng = np.random.default_rng(1)
    U, I, F, L, B, p, seed = 60, 50, 8, 2, 512, 0.3, 12345
    D = F * 2 ** (L - 1)
    tabs_h = [(rng.standard_normal(s) * 0.5).astype(np.float32) for s in ((U, F), (I, F), (U, D), (I, D))]
    W_h = (rng.standard_normal(ops.neumf_param_count(F, L)) * 0.4).astype(np.float32)
    b = [dev(rng.integers(m, size=B).astype(np.int32)) for m in (U, I, I)]
    lr = 1e-3
    hp = ops.hyper(lr, 0.0, 0.0, "sgd")
    ws = ops.NeumfWorkspace(U, I, F, L, "sgd", 2 * B, "cuda")

    def loss_at(W_np, tabs_np):
        return ops.neumf_bpr_train_steps([dev(t) for t in tabs_np], dev(W_np), ws, *b, B, 0, 1, hp, adam_step0=5, apply=False,
                                         tower_dtype=tower_dtype, dropout=p, dropout_seed=seed).item()
    l_drop, l_nodrop = loss_at(W_h, tabs_h), ops.neumf_bpr_train_steps([dev(t) for t in tabs_h], dev(W_h), ws, *b, B, 0, 1, hp,
                                                                       apply=False).item()
    assert l_drop != l_nodrop and loss_at(W_h, tabs_h) == l_drop          # masks active and deterministic
    tabs, W = [dev(t) for t in tabs_h], dev(W_h)
    ops.neumf_bpr_train_steps(tabs, W, ws, *b, B, 0, 1, hp, adam_step0=5, tower_dtype=tower_dtype, dropout=p, dropout_seed=seed)
    gW = (W_h - W.cpu().numpy()) / lr
    gUM = (tabs_h[2] - tabs[2].cpu().numpy()) / lr
    eps = 2e-2
    # numerical gradient on the largest-gradient coordinates of the tower block and of the user MLP table
    for k in np.argsort(-np.abs(gW))[:6]:
        Wp, Wm = W_h.copy(), W_h.copy()
        Wp[k] += eps; Wm[k] -= eps
        num = (loss_at(Wp, tabs_h) - loss_at(Wm, tabs_h)) / (2 * eps)
        assert abs(num - gW[k]) <= 0.03 * abs(gW[k]) + 0.05, (k, num, gW[k])
    for flat in np.argsort(-np.abs(gUM).ravel())[:4]:
        r, c = divmod(int(flat), D)
        tp, tm = [t.copy() for t in tabs_h], [t.copy() for t in tabs_h]
        tp[2][r, c] += eps; tm[2][r, c] -= eps
        num = (loss_at(W_h, tp) - loss_at(W_h, tm)) / (2 * eps)
        assert abs(num - gUM[r, c]) <= 0.03 * abs(gUM[r, c]) + 0.05, (r, c, num, gUM[r, c])


def _neumf_cfg(g, c, **kw):
    U, I, F, L, lr, r1, r2, opt, seed, drop = g[f"c{c}_hyper"]
    cfg = dict(gpu="", logger=logging.getLogger("t"), lr=float(lr), reg_1=float(r1), reg_2=float(r2), epochs=1, topk=10,
               user_num=int(U), item_num=int(I), factors=int(F), num_layers=int(L), dropout=float(drop),
               model_name=str(g[f"c{c}_name"]), loss_type="BPR", optimizer="sgd" if opt == 0 else "default",
               init_method="default", early_stop=False, progress=False, GMF_model=None, MLP_model=None)
    cfg.update(kw)
    return cfg, int(seed)


def test_neumf_default_dropout_and_modes_match_reference():
    """The reference's default config (dropout 0.5) and model_name GMF / MLP / NeuMF-pre through the class surface: host-drawn
    torch masks (dropout_engine 'auto' -> 'torch' at these sizes) reproduce the reference's losses, tables and RNG position;
    NeuMF-pre reproduces the reference's initial state bit for bit, quirk of :116 and nn.Linear's own bias draw included."""
    from daisyrec_b200.model import NeuMF
    from daisyrec_b200.utils.dataset import CandidatesDataset, get_dataloader
    g = golden("neumf_modes")
    trained = {}
    for c in range(int(g["ncases"])):
        name = str(g[f"c{c}_name"])
        cfg, seed = _neumf_cfg(g, c, GMF_model=trained.get("GMF"), MLP_model=trained.get("MLP"))
        torch.manual_seed(seed)
        m = NeuMF(cfg)
        init = {"embed_user_GMF.weight": g[f"c{c}_UG"][0], "embed_item_GMF.weight": g[f"c{c}_IG"][0],
                "embed_user_MLP.weight": g[f"c{c}_UM"][0], "embed_item_MLP.weight": g[f"c{c}_IM"][0], "tower": g[f"c{c}_W"][0]}
        if name == "NeuMF-pre":
            for k, want in init.items():                                  # copied tables + the reference's predict layer
                assert np.array_equal(m.state_dict()[k].cpu().numpy(), want), (name, k)
        else:
            scale = {k: (3.0 if k != "tower" else 1.0) for k in init}     # the generator spreads the tables by 3x
            for k, want in init.items():
                np.testing.assert_allclose(m.state_dict()[k].cpu().numpy() * scale[k], want, rtol=1e-6, atol=0, err_msg=k)
            m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in init.items()})
        opt_is_sgd = cfg["optimizer"] == "sgd"
        bs, losses = g[f"c{c}_batches"], g[f"c{c}_loss"]
        m.train()
        torch.manual_seed(seed + 100)
        for s in range(bs.shape[0]):
            loss = m.train_step([torch.from_numpy(np.ascontiguousarray(bs[s][k])) for k in range(3)])
            assert abs(loss - losses[s]) <= 3e-5 * abs(losses[s]), (c, name, s, loss, losses[s])
            tol = (5e-6 if opt_is_sgd else 1e-4) * (s + 1)
            for k, fx in (("embed_user_GMF.weight", "UG"), ("embed_item_GMF.weight", "IG"), ("embed_user_MLP.weight", "UM"),
                          ("embed_item_MLP.weight", "IM"), ("tower", "W")):
                want = g[f"c{c}_{fx}"][s + 1]
                got = m.state_dict()[k].cpu().numpy()
                bad = np.abs(got - want) > tol * max(1.0, np.abs(want).max())
                assert bad.mean() <= (0.0 if opt_is_sgd else 0.01), (c, name, s, fx, float(bad.mean()))
        assert np.array_equal(torch.get_rng_state().numpy(), g[f"c{c}_rng_after"]), (c, name)
        m.eval()
        trained[name] = m
        m.load_state_dict({"embed_user_GMF.weight": g[f"c{c}_UG"][-1], "embed_item_GMF.weight": g[f"c{c}_IG"][-1],
                           "embed_user_MLP.weight": g[f"c{c}_UM"][-1], "embed_item_MLP.weight": g[f"c{c}_IM"][-1],
                           "tower": g[f"c{c}_W"][-1]})
        users, cands = g[f"c{c}_users"], g[f"c{c}_cands"].astype(np.int64)
        loader = get_dataloader(CandidatesDataset([[int(u), cands[r]] for r, u in enumerate(users)]), batch_size=128,
                                shuffle=False)
        assert (m.rank(loader) == g[f"c{c}_preds"]).mean() >= 0.97, (c, name)
        assert (np.stack([m.full_rank(int(u)) for u in users[:3]]) == g[f"c{c}_full"]).mean() >= 0.9
        pp = np.array([m.predict(int(users[q]), int(cands[q][0])) for q in range(4)], np.float32)
        np.testing.assert_allclose(pp, g[f"c{c}_pred_pairs"], rtol=3e-5, atol=3e-6)


def test_neumf_modes_match_oracle(ops, orc):
    """GMF / MLP steps and host-mask dropout on random inputs against the pinned oracle (same masks on both sides)."""
    rng = np.random.default_rng(4)
    U, I, F, L, B = 200, 150, 16, 2, 500
    D = F * 2 ** (L - 1)
    for name, drop in (("GMF", 0.0), ("MLP", 0.4), ("NeuMF", 0.5)):
        mode = ops.NEUMF_MODE[name]
        tabs_h = [(rng.standard_normal(s) * 0.1).astype(np.float32) for s in ((U, F), (I, F), (U, D), (I, D))]
        W_h = (rng.standard_normal(ops.neumf_param_count(F, L, mode)) * 0.2).astype(np.float32)
        b = [rng.integers(n, size=B).astype(np.int32) for n in (U, I, I)]
        keep = None
        masks = None
        if drop > 0:
            torch.manual_seed(3)
            keep = orc.torch_dropout_keep(B, F, L, drop)
            words, col = [], 0
            for n in [F * (2 ** (L - i)) for i in range(L)]:
                bits = np.packbits((keep[:, col:col + n] != 0).reshape(-1), bitorder="little")
                words.append(np.pad(bits, (0, (-len(bits)) % 4)).view(np.int32))
                col += n
            masks = dev(np.concatenate(words))
        tabs, W = [dev(t) for t in tabs_h], dev(W_h)
        ws = ops.NeumfWorkspace(U, I, F, L, "sgd", 2 * B, "cuda")
        hp = ops.hyper(0.05, 0.001, 0.002, "sgd")
        loss = ops.neumf_bpr_train_steps(tabs, W, ws, *[dev(x) for x in b], B, 0, 1, hp, dropout=drop, drop_masks=masks,
                                         mode=mode).item()
        lo = orc.neumf_bpr_step(tabs_h, W_h, F, L, *b, orc.hyper(0.05, 0.001, 0.002, "sgd"), True, None, 1, mode=mode, keep=keep)
        assert abs(loss - lo) <= 2e-5 * abs(lo), (name, loss, lo)
        for q in range(4):
            np.testing.assert_allclose(tabs[q].cpu().numpy(), tabs_h[q], rtol=0, atol=5e-6, err_msg=f"{name} table {q}")
        np.testing.assert_allclose(W.cpu().numpy(), W_h, rtol=0, atol=2e-5, err_msg=name)
