"""NeuMF + BPR on the B200 path, with the reference's class name, config keys and methods
(daisy/model/NeuMFRecommender.py:15-232, model_name 'NeuMF').

Four raw fp32 embedding tables (``embed_user_GMF/embed_item_GMF/embed_user_MLP/embed_item_MLP`` ``.weight``)
and the tower as one flat fp32 block (``tower``; per layer weight then bias, then predict weight and bias).
No nn.Module in the compute path, no autograd, no torch.optim: fit / calc_loss / rank / full_rank / predict
forward to ``drb_neumf_*`` (include/daisyrec_b200.h).  torch.nn is touched at construction time only, to
consume the global CPU RNG exactly as the reference's constructor does (bit-identical initial weights).
"""
import numpy as np
import torch

from .. import ops
from .AbstractRecommender import GeneralRecommender, _Table, _INIT


class NeuMF(GeneralRecommender):
    def __init__(self, config):
        super().__init__(config)
        if self.world > 1:
            raise NotImplementedError('NeuMF runs as independent replicas only (DESIGN.md, multi-GPU section)')
        self.lr = config['lr']
        self.epochs = config['epochs']
        self.reg_1 = config['reg_1']
        self.reg_2 = config['reg_2']
        self.dropout = config['dropout']
        self.model = config['model_name']
        self.user_num, self.item_num = config['user_num'], config['item_num']
        self.factors, self.num_layers = config['factors'], config['num_layers']
        self.loss_type = config['loss_type']
        self.optimizer = config['optimizer'] if config['optimizer'] != 'default' else 'adam'
        self.initializer = config['init_method'] if config['init_method'] != 'default' else 'xavier_normal'
        self.early_stop = config['early_stop']
        self.topk = config['topk']
        if self.model != 'NeuMF':
            raise NotImplementedError(f"model_name={self.model!r}: the B200 hot path covers 'NeuMF' (GMF / MLP / NeuMF-pre "
                                      "are outside BASELINE.json's configs)")
        self.dropout = float(self.dropout or 0.0)
        if not (0.0 <= self.dropout < 1.0):
            raise ValueError(f'dropout must be in [0, 1), got {self.dropout}')
        if self.factors % 4 != 0:
            raise NotImplementedError('NeuMF on the B200 path needs factors to be a multiple of 4 (128-bit rows)')
        F, Ln = self.factors, self.num_layers
        D = F * (2 ** (Ln - 1))
        self.mlp_dim = D

        # ---- reference RNG stream (NeuMFRecommender.py:52-71 constructors, :81-96 _init_weight) on the CPU
        import torch.nn as nn
        embs = [nn.Embedding(self.user_num, F), nn.Embedding(self.item_num, F),
                nn.Embedding(self.user_num, D), nn.Embedding(self.item_num, D)]
        linears = [nn.Linear(F * (2 ** (Ln - i)), F * (2 ** (Ln - i)) // 2) for i in range(Ln)]
        predict = nn.Linear(2 * F, 1)
        init = _INIT[self.initializer]
        with torch.no_grad():
            for e in embs:
                init(e.weight)
            bare = {'normal': torch.nn.init.normal_, 'uniform': torch.nn.init.uniform_,
                    'xavier_normal': torch.nn.init.xavier_normal_, 'xavier_uniform': torch.nn.init.xavier_uniform_}
            for lin in linears:
                bare[self.initializer](lin.weight)      # :88-90 passes NO param config to the hidden layers: 'normal' is N(0, 1)
            init(predict.weight)
            for lin in linears + [predict]:
                lin.bias.zero_()
            parts = []
            for lin in linears:
                parts += [lin.weight.reshape(-1), lin.bias.reshape(-1)]
            parts += [predict.weight.reshape(-1), predict.bias.reshape(-1)]
            tower = torch.cat(parts).contiguous()
        assert tower.numel() == ops.neumf_param_count(F, Ln)
        self.embed_user_GMF = _Table(embs[0].weight.detach().to(self.device).contiguous())
        self.embed_item_GMF = _Table(embs[1].weight.detach().to(self.device).contiguous())
        self.embed_user_MLP = _Table(embs[2].weight.detach().to(self.device).contiguous())
        self.embed_item_MLP = _Table(embs[3].weight.detach().to(self.device).contiguous())
        self.tower = tower.to(self.device)
        self._ws = None
        self._opt_steps = 0
        self._rows = int(config.get('neumf_scratch_rows', 1 << 16))
        # optional B200 key: 'fp32' (CUDA cores, parity path, default) | 'bf16' (tcgen05 tensor cores, BASELINE config 3)
        td = str(config.get('tower_dtype', 'fp32')).lower()
        if td not in ('fp32', 'bf16'):
            raise ValueError(f"tower_dtype must be 'fp32' or 'bf16', got {td!r}")
        self._tower_dtype = 1 if td == 'bf16' else 0

    # ------------------------------------------------------------------ plumbing
    def _tabs(self):
        return (self.embed_user_GMF.weight, self.embed_item_GMF.weight, self.embed_user_MLP.weight,
                self.embed_item_MLP.weight)

    def parameters(self):
        return list(self._tabs()) + [self.tower]

    def state_dict(self):
        return {'embed_user_GMF.weight': self.embed_user_GMF.weight, 'embed_item_GMF.weight': self.embed_item_GMF.weight,
                'embed_user_MLP.weight': self.embed_user_MLP.weight, 'embed_item_MLP.weight': self.embed_item_MLP.weight,
                'tower': self.tower}

    def load_state_dict(self, sd):
        for k, t in self.state_dict().items():
            t.copy_(sd[k])

    def _hyper(self, opt=None):
        return ops.hyper(self.lr, self.reg_1, self.reg_2, opt or self._optimizer_name())

    def _workspace(self, rows, opt=None, fresh=False):
        rows = max(int(rows), self._rows)
        if fresh or self._ws is None or self._ws.max_rows < rows:
            keep = None if fresh or self._ws is None else self._ws
            if keep is not None and opt is None:
                # growing the scratch would drop the optimiser state: size it up front instead
                raise RuntimeError('NeuMF scratch too small; set config["neumf_scratch_rows"] >= 2 * batch_size')
            self._ws = ops.NeumfWorkspace(self.user_num, self.item_num, self.factors, self.num_layers,
                                          opt or self._optimizer_name(), rows, self.device)
        return self._ws

    def _begin_fit(self, opt):
        # dropout masks are counter-based (Philox) on the device; the key is drawn from torch's global RNG so that
        # torch.manual_seed makes runs reproducible (the masks themselves are NOT torch's: parity holds at dropout=0)
        self._drop_seed = int(torch.empty((), dtype=torch.int64).random_().item()) if self.dropout > 0 else 0
        self._hp = self._hyper(opt)
        self._opt_steps = 0
        self._fit_opt = opt
        self._ws = None                                              # fresh optimiser state per fit()

    def _train_steps(self, bu, bi, bj, batch, first, n_steps):
        if self._ws is None:
            self._workspace(2 * batch, self._fit_opt, fresh=True)
        losses = ops.neumf_bpr_train_steps(self._tabs(), self.tower, self._ws, bu, bi, bj, batch, first, n_steps, self._hp,
                                           adam_step0=self._opt_steps, tower_dtype=self._tower_dtype,
                                           dropout=self.dropout if self.training else 0.0, dropout_seed=self._drop_seed)
        self._opt_steps += n_steps
        return losses

    def _ensure(self, rows):
        if self._ws is None:
            self._begin_fit(self._optimizer_name())
            self._workspace(rows, self._fit_opt, fresh=True)
        elif self._ws.max_rows < rows:
            self._workspace(rows)

    # ------------------------------------------------------------------ reference surface
    def forward(self, user, item):
        u = torch.as_tensor(user).to(self.device, torch.int64).reshape(-1).contiguous()
        i = torch.as_tensor(item).to(self.device, torch.int64).reshape(-1, 1).contiguous()
        self._ensure(1)
        return ops.neumf_scores(self._tabs(), self.tower, self._ws, u, i, 1, self._tower_dtype).reshape(-1)

    __call__ = forward

    def calc_loss(self, batch):
        self._check_loss_type()
        bu, bi, bj = (torch.as_tensor(b).to(self.device, torch.int32).contiguous() for b in batch[:3])
        self._ensure(2 * bu.numel())
        loss = ops.neumf_bpr_train_steps(self._tabs(), self.tower, self._ws, bu, bi, bj, bu.numel(), 0, 1, self._hp,
                                         apply=False, tower_dtype=self._tower_dtype, adam_step0=self._opt_steps,
                                         dropout=self.dropout if self.training else 0.0, dropout_seed=self._drop_seed)
        return loss.to(torch.float32).reshape(())

    def train_step(self, batch):
        self._check_loss_type()
        bu, bi, bj = (torch.as_tensor(b).to(self.device, torch.int32).contiguous() for b in batch[:3])
        self._ensure(2 * bu.numel())
        was = self.training
        self.train()                                                 # a training step runs in train mode (dropout on)
        try:
            return float(self._train_steps(bu, bi, bj, bu.numel(), 0, 1).item())
        finally:
            self.train(was)

    def predict(self, u, i):
        return float(self.forward([int(u)], [int(i)]).item())

    def rank(self, test_loader):
        data = getattr(getattr(test_loader, 'dataset', None), 'data', None)
        if isinstance(data, (list, tuple)) and len(data) and len(data[0]) == 2:
            users = np.fromiter((int(r[0]) for r in data), np.int64, len(data))
            cands = np.stack([np.asarray(r[1], dtype=np.int64) for r in data])
        else:
            us, cs = [], []
            for b_us, b_c in test_loader:
                us.append(torch.as_tensor(b_us).reshape(-1).to(torch.int64))
                cs.append(torch.as_tensor(b_c).to(torch.int64).reshape(us[-1].numel(), -1))
            if not us:
                return np.zeros((0,), np.float32)
            users, cands = torch.cat(us).numpy(), torch.cat(cs).numpy()
        self._ensure(1)
        d_users = torch.from_numpy(users).to(self.device)
        d_cands = torch.from_numpy(np.ascontiguousarray(cands)).to(self.device)
        scores = ops.neumf_scores(self._tabs(), self.tower, self._ws, d_users, d_cands, cands.shape[1], self._tower_dtype)
        k = min(self.topk, cands.shape[1])
        return ops.topk_from_scores(scores, d_cands, k).cpu().numpy()

    def full_rank(self, u):
        self._ensure(1)
        users = torch.tensor([int(u)], dtype=torch.int64, device=self.device)
        scores = ops.neumf_scores(self._tabs(), self.tower, self._ws, users, None, self.item_num, self._tower_dtype)
        return ops.topk_from_scores(scores, None, min(self.topk, self.item_num))[0].cpu().numpy()
