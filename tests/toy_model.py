"""A small CPU model that speaks the reference's pipeline tuple protocol the way models/sdxl.py does (BASELINE config 0,
"plumbing"): five inputs including int64 ids, a variable-length residual tuple between layers, a bool scalar tensor in
the tuple, a final layer returning (prediction, timesteps) and a loss_fn that unpacks it and applies per-sample
min-SNR-style weights (reference: models/sdxl.py:333-355, 579, 632-651, 695, 796-801, 995)."""
import torch
from torch import nn


class First(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.inp = nn.Linear(8, d)
        self.emb = nn.Embedding(16, d)
        self.tproj = nn.Linear(1, d)

    def forward(self, inputs):
        for item in inputs:
            if torch.is_floating_point(item):
                item.requires_grad_(True)
        x, ids, timesteps, cond, flag = inputs
        h = self.inp(x) + self.emb(ids).mean(1, keepdim=True) + self.tproj(timesteps[:, None, None].float())
        return h, cond, timesteps, flag


class Down(nn.Module):
    """appends its output to the residual tuple (variable-length tuple between layers)"""

    def __init__(self, d):
        super().__init__()
        self.lin = nn.Linear(d, d)
        # frozen "base" weight + trainable low-rank factors, LoRA style
        self.lin.weight.requires_grad_(False)
        self.a = nn.Parameter(torch.randn(4, d) * 0.1)
        self.b = nn.Parameter(torch.randn(d, 4) * 0.1)

    def forward(self, inputs):
        h, cond, timesteps, flag, *res = inputs
        h2 = torch.tanh(self.lin(h) + (h @ self.a.t()) @ self.b.t() + cond[:, None, :])
        return (h2, cond, timesteps, flag, *res, h)


class Up(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.lin = nn.Linear(2 * d, d)

    def forward(self, inputs):
        h, cond, timesteps, flag, *res = inputs
        skip = res.pop()
        h2 = torch.relu(self.lin(torch.cat([h, skip], dim=-1)))
        if bool(flag.item()):
            h2 = h2 * 1.5
        return (h2, cond, timesteps, flag, *res)


class Last(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.out = nn.Linear(d, 8)

    def forward(self, inputs):
        h, cond, timesteps, flag = inputs
        return self.out(h), timesteps


def loss_fn(output, label):
    pred, timesteps = output
    target, mask = label
    loss = torch.nn.functional.mse_loss(pred.float(), target.float(), reduction='none')
    if mask.numel() > 0:
        loss = loss * mask.float()
    loss = loss.mean(dim=list(range(1, loss.ndim)))
    snr = 1.0 / (1.0 + timesteps.float())
    return (loss * torch.clamp(snr, max=5.0)).mean()


def make_layers(d=16, seed=0):
    torch.manual_seed(seed)
    layers = [First(d), Down(d), Down(d), Up(d), Up(d), Last(d)]
    for li, l in enumerate(layers):
        for n, p in l.named_parameters():
            p.original_name = f'{li}.{n}'
    return layers


def make_micro_batches(n, bs, seed, with_mask=True):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        x = torch.randn(bs, 5, 8, generator=g)
        ids = torch.randint(0, 16, (bs, 3), generator=g)
        ts = torch.randint(0, 10, (bs,), generator=g)
        cond = torch.randn(bs, 16, generator=g)
        flag = torch.tensor(True)
        target = torch.randn(bs, 5, 8, generator=g)
        mask = (torch.rand(bs, 5, 8, generator=g) > 0.2).float() if with_mask else torch.tensor([])
        out.append(((x, ids, ts, cond, flag), (target, mask)))
    return out
