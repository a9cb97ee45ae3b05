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


class GradBucket:
    """One persistent flat fp32 buffer holding the gradients of several modules, each ``p.grad`` a VIEW into it
    ("one ncclAllReduce(sum) over a flat fp32 bucket of actor+critic grads").

    rlsb_ac_update writes the actor's and the critic's gradients straight into these views (ops._mlp_grads hands the
    kernels ``p.grad.data_ptr()``), so the data-parallel step is a single all-reduce of ``flat`` — no torch.cat, no
    copy-back — issued after both backward passes and before either ``clip_grad_norm_``.  The views are stable
    addresses, which is what a captured CUDA graph of the update needs."""

    def __init__(self, params: t.Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket needs at least one trainable parameter")
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=torch.float32)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
        self.attach()

    def attach(self) -> None:
        """(Re-)install the views as ``.grad``; a gradient somebody else put there meanwhile is copied in first."""
        if all(p.grad is None for p in self.params):   # the usual case after zero_grad(set_to_none=True): ONE fill, not one
            self.flat.zero_()                           # per parameter (~ 45 launches per update otherwise)
            for p, v in zip(self.params, self.views):
                p.grad = v
            return
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v

    def attached(self) -> bool:
        return all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(self.params, self.views))

    def all_reduce(self) -> bool:
        """Average the bucket over the ranks with ONE collective; returns False when there is nothing to do (no
        process group / a single rank)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return False
        if not self.attached():
            self.attach()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.div_(dist.get_world_size())
        return True


class Optimizer:
    def __init__(self, model, lr=1e-4, eps=1e-8, weight_decay=0.01,
                 lr_scheduler: t.Optional[t.Type[LRScheduler] | t.Iterable[t.Type[LRScheduler]]] = None,
                 scaler: bool = False, log_grad: bool = False, clip: t.Optional[float] = None):
        self.model = model
        params = list(model.parameters())
        # CUDA parameters: torch's fused AdamW (one multi-tensor kernel per step instead of ~10 foreach launches of
        # 20 us each — 10 % of the step at the configured 800 start states); same update rule, same state_dict layout
        fused = bool(params) and all(p.is_cuda and p.dtype == torch.float32 for p in params)
        self.optimizer = torch.optim.AdamW(params, lr=lr, eps=eps, weight_decay=weight_decay, **({"fused": True} if fused else {}))
        self._fused = fused
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

    def step_with_grads(self, reduced: bool = False):
        """Same as ``step`` from the all-reduce on, for gradients a kernel already wrote into ``.grad``
        (rlsb_ac_update replaces zero_grad + loss.backward(), optimizer.py:55-57).  ``reduced``: the caller has
        already averaged the gradients over the ranks (GradBucket.all_reduce)."""
        return self._apply(reduced)

    def _apply(self, reduced: bool = False):
        metrics = {}
        if not reduced:
            allreduce_grads_(self.model.parameters())
        if self.log_grad:
            for tag, value in self.model.named_parameters():
                metrics[f"grad/{tag.replace('.', '/')}"] = value.detach()
        if self.clip:
            nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
        if self._fused:
            # the fused kernel wants every gradient in its parameter's dtype and strides (autograd hands conv-weight
            # gradients back in whatever memory format cuDNN chose; the optimizer state follows the parameter)
            for p in self.model.parameters():
                g = p.grad
                if g is not None and (g.stride() != p.stride() or g.dtype != p.dtype):
                    p.grad = torch.empty_like(p).copy_(g)
        if self.scaler:
            self.scaler.step(self.optimizer)
            self.scaler.update()
        else:
            self.optimizer.step()
        if self.lr_scheduler:
            self.lr_scheduler.step()
            metrics[f'lr/{self.model.__class__.__name__}'] = torch.Tensor(self.lr_scheduler.get_last_lr())
        return metrics
