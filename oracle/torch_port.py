"""PyTorch-CPU port of the reference's BPR-MF training step -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python on top of PyTorch; /root/reference does not exist on the GPU
box, so the "reference arm" of bench.py (--impl reference) and the cpu_baseline leg time THIS
port: the same library calls in the same order as the reference (dense nn.Embedding-style
parameters, autograd backward producing table-sized dense gradients, torch.optim.SGD / Adam
over the full tables), written from the algorithm, not copied.  It is pinned against the
reference's own outputs by tests/test_oracle_golden.py::test_torch_port_matches_reference.

Restates: daisy/model/MFRecommender.py:63-97 (forward, calc_loss with un-squared L1 / Frobenius
regularisers of the gathered batch rows), daisy/utils/loss.py:5-13 (BPRLoss, gamma=1e-10, sum),
daisy/model/AbstractRecommender.py:48-67 (optimizer factory) and :119-128 (zero_grad, calc_loss,
isnan guard, backward, step, loss.item()).
"""
import torch


class TorchMFBaseline:
    def __init__(self, P0, Q0, lr=0.01, reg_1=0.001, reg_2=0.001, optimizer="sgd"):
        # dense (sparse=False) embedding weights, like nn.Embedding in MFRecommender.py:53-54
        self.P = torch.nn.Parameter(torch.as_tensor(P0, dtype=torch.float32).clone())
        self.Q = torch.nn.Parameter(torch.as_tensor(Q0, dtype=torch.float32).clone())
        self.reg_1, self.reg_2 = reg_1, reg_2
        if optimizer == "sgd":
            self.opt = torch.optim.SGD([self.P, self.Q], lr=lr)          # AbstractRecommender.py:55-56
        else:
            self.opt = torch.optim.Adam([self.P, self.Q], lr=lr)         # AbstractRecommender.py:53-54

    def _score(self, u, i):                                               # MFRecommender.py:63-68
        return (torch.nn.functional.embedding(u, self.P) * torch.nn.functional.embedding(i, self.Q)).sum(dim=-1)

    def loss(self, u, i, j):                                              # MFRecommender.py:70-97, loss.py:11
        emb = torch.nn.functional.embedding
        pos, neg = self._score(u, i), self._score(u, j)
        out = -(1e-10 + torch.sigmoid(pos - neg)).log().sum()
        out = out + self.reg_1 * (emb(i, self.Q).norm(p=1) + emb(j, self.Q).norm(p=1))
        out = out + self.reg_2 * (emb(i, self.Q).norm() + emb(j, self.Q).norm())
        out = out + self.reg_1 * emb(u, self.P).norm(p=1)
        out = out + self.reg_2 * emb(u, self.P).norm()
        return out

    def step(self, u, i, j):                                              # AbstractRecommender.py:119-128
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss(u, i, j)
        if torch.isnan(loss):
            raise ValueError("Loss=Nan or Infinity: current settings does not fit the recommender")
        loss.backward()
        self.opt.step()
        return loss.item()
