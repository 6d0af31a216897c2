"""Seeded synthetic implicit-feedback data of a given shape.

Data generation is setup, not hot path: it uses ordinary torch ops on whatever device it is given
(GPU for the bench, CPU for the reference arm / tests).  Users' degrees are log-normal (clipped to
[10, item_num/4], rescaled to the requested nnz); items follow a Zipf(1.0) popularity over a
random relabelling; (user, item) pairs are unique; COO rows come in a random "time" order.
"""
import torch

SHAPES = {
    # name: (user_num, item_num, nnz)   -- BASELINE.json configs
    "ml-20m": (138_493, 26_744, 20_000_000),
    "amazon-book": (52_643, 91_599, 3_000_000),
    "netflix": (480_189, 17_770, 100_000_000),
    "tiny": (2_000, 1_500, 60_000),
}


def make_interactions(user_num, item_num, nnz, seed=2022, device="cpu", alpha=1.0):
    """-> dict(row_ptr int64[U+1], col int32[nnz'] sorted CSR, coo_u/coo_i int32[nnz'] time order)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lo, hi = 10.0, max(10.0, item_num / 4.0)
    raw = torch.exp(torch.randn(user_num, generator=g, device=dev))
    rank = torch.arange(1, item_num + 1, device=dev, dtype=torch.float64)
    pop = (1.0 / rank.pow(alpha))
    pop = pop[torch.randperm(item_num, generator=g, device=dev)]
    cdf = torch.cumsum(pop / pop.sum(), 0)
    cdf[-1] = 1.0
    over = 1.25
    keys = None
    for _ in range(4):
        deg = raw * (nnz * over / raw.sum())
        deg = deg.clamp_(lo, hi).round_().to(torch.int64)
        users = torch.repeat_interleave(torch.arange(user_num, device=dev), deg)
        items = torch.searchsorted(cdf, torch.rand(users.numel(), generator=g, device=dev, dtype=torch.float64))
        items.clamp_(max=item_num - 1)
        keys = torch.unique(users * item_num + items)              # sorted: user-major, item-minor
        del users, items
        if keys.numel() >= nnz:
            break
        over *= 1.25 * nnz / max(1, keys.numel())
    if keys.numel() > nnz:                                           # trim uniformly, keep each user's first row
        u_of = keys // item_num
        first = torch.ones_like(keys, dtype=torch.bool)
        first[1:] = u_of[1:] != u_of[:-1]
        extra = keys.numel() - nnz
        score = torch.rand(keys.numel(), generator=g, device=dev)
        score[first] = 2.0
        drop = torch.topk(score, extra, largest=False).indices
        keep = torch.ones_like(first)
        keep[drop] = False
        keys = keys[keep]
    u = (keys // item_num).to(torch.int32)
    i = (keys % item_num).to(torch.int32)
    row_ptr = torch.zeros(user_num + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(torch.bincount(u.to(torch.int64), minlength=user_num), 0)
    order = torch.randperm(keys.numel(), generator=g, device=dev)
    return dict(row_ptr=row_ptr, col=i.contiguous(), coo_u=u[order].contiguous(), coo_i=i[order].contiguous(),
                user_num=user_num, item_num=item_num, nnz=int(keys.numel()))


def init_tables(user_num, item_num, factors, seed=2022, device="cpu", std=0.01):
    """normal(0, 0.01) tables (MF default init, AbstractRecommender.py:20,:69-77), seeded on CPU then moved."""
    g = torch.Generator()
    g.manual_seed(seed)
    P = torch.empty(user_num, factors).normal_(0.0, std, generator=g)
    Q = torch.empty(item_num, factors).normal_(0.0, std, generator=g)
    return P.to(device), Q.to(device)
