"""AdamW wrapper with clipping and chained LR schedulers (reference: rl_sandbox/utils/optimizer.py:11-71).

B200 addition: when torch.distributed is initialised, gradients are averaged across
ranks with ONE flat-bucket all-reduce issued between backward and clip_grad_norm_, so every rank
clips by the same global norm.  World size 1 (or no process group) is the reference's behaviour.
"""
import typing as t
from collections.abc import Iterable

import torch
from torch import nn
from torch.optim.lr_scheduler import LambdaLR, LinearLR, LRScheduler


class WarmupScheduler(LinearLR):
    def __init__(self, optimizer, warmup_steps):
        super().__init__(optimizer, start_factor=1 / warmup_steps, total_iters=int(warmup_steps))


class DecayScheduler(LambdaLR):
    def __init__(self, optimizer, decay_steps, decay_rate):
        super().__init__(optimizer, lambda epoch: decay_rate ** (epoch / decay_steps))


def allreduce_grads_(params: t.Iterable[torch.Tensor]) -> None:
    """Average .grad over ranks through one flat fp32 bucket (NCCL over NVLink on B200, gloo in CPU tests)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(dist.get_world_size())
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


class Optimizer:
    def __init__(self, model, lr=1e-4, eps=1e-8, weight_decay=0.01,
                 lr_scheduler: t.Optional[t.Type[LRScheduler] | t.Iterable[t.Type[LRScheduler]]] = None,
                 scaler: bool = False, log_grad: bool = False, clip: t.Optional[float] = None):
        self.model = model
        self.optimizer = torch.optim.AdamW(model.parameters(), lr=lr, eps=eps, weight_decay=weight_decay)
        if isinstance(lr_scheduler, Iterable):
            lr_scheduler = torch.optim.lr_scheduler.ChainedScheduler(
                [make(optimizer=self.optimizer) for make in lr_scheduler])
        elif lr_scheduler is not None:
            lr_scheduler = lr_scheduler(optimizer=self.optimizer)
        self.lr_scheduler = lr_scheduler
        self.log_grad = log_grad
        self.scaler = torch.amp.GradScaler() if scaler else None
        self.clip = clip

    def step(self, loss):
        self.optimizer.zero_grad(set_to_none=True)
        (self.scaler.scale(loss) if self.scaler else loss).backward()
        if self.scaler:
            self.scaler.unscale_(self.optimizer)
        return self._apply()

    def step_with_grads(self):
        """Same as ``step`` from the all-reduce on, for gradients a kernel already wrote into ``.grad``
        (rlsb_ac_update replaces zero_grad + loss.backward(), optimizer.py:55-57)."""
        return self._apply()

    def _apply(self):
        metrics = {}
        allreduce_grads_(self.model.parameters())
        if self.log_grad:
            for tag, value in self.model.named_parameters():
                metrics[f"grad/{tag.replace('.', '/')}"] = value.detach()
        if self.clip:
            nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
        if self.scaler:
            self.scaler.step(self.optimizer)
            self.scaler.update()
        else:
            self.optimizer.step()
        if self.lr_scheduler:
            self.lr_scheduler.step()
            metrics[f'lr/{self.model.__class__.__name__}'] = torch.Tensor(self.lr_scheduler.get_last_lr())
        return metrics
