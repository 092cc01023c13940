"""ORACLE — test infrastructure only.  Single-process restatement of one DeepSpeed pipeline-engine step as the
reference drives it (train.py:915-918):

    for each of GAS micro-batches:  loss_mb = loss_fn(layers(inputs), labels);  (loss_mb / GAS).backward()
    [data-parallel mean of gradients is the caller's business: pass every rank's micro-batches]
    clip to max global L2 norm (utils/patches.py:175-246; skipped when gradient_clipping == 0)
    optimizer.step(); optimizer.zero_grad(); lr_scheduler.step()
    returns mean(loss_mb)

The stage partition does not change the arithmetic, so a multi-stage / multi-rank run of the product engine must
reproduce this loss and these gradients (fp32, CPU: to ~1e-6).
"""
import torch


class RefPipelineEngine:
    def __init__(self, layers, loss_fn, optimizer=None, lr_scheduler=None, gradient_accumulation_steps=1,
                 gradient_clipping=0.0):
        self.layers = list(layers)
        self.loss_fn = loss_fn
        self.optimizer = optimizer
        self.lr_scheduler = lr_scheduler
        self.gas = gradient_accumulation_steps
        self.clip = gradient_clipping
        self.grad_norm = None

    def parameters(self):
        seen = set()
        for l in self.layers:
            if isinstance(l, torch.nn.Module):
                for p in l.parameters():
                    if id(p) not in seen:
                        seen.add(id(p))
                        yield p

    def forward(self, inputs):
        x = tuple(t.clone().detach().requires_grad_(t.is_floating_point()) for t in inputs)
        x = x if len(x) > 1 else x[0]
        for l in self.layers:
            x = l(x)
        return x

    def train_batch(self, micro_batches, step=True):
        """micro_batches: list of (features_tuple, labels)."""
        assert len(micro_batches) == self.gas
        total = 0.0
        for feats, labels in micro_batches:
            loss = self.loss_fn(self.forward(feats), labels)
            (loss / self.gas).backward()
            total = total + loss.detach()
        params = [p for p in self.parameters() if p.requires_grad and p.grad is not None]
        if self.clip > 0:
            norms = [p.grad.detach().float().norm(2) for p in params]
            total_norm = torch.stack(norms).square().sum().sqrt()
            coef = torch.clamp(self.clip / (total_norm + 1e-6), max=1.0)
            for p in params:
                p.grad.mul_(coef)
            self.grad_norm = total_norm
        if step and self.optimizer is not None:
            self.optimizer.step()
            self.optimizer.zero_grad(set_to_none=True)
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()
        return total / self.gas
