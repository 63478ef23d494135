"""Data-parallel step plumbing for the training config (cfg 5): ONE exchange step per iteration -
an all-reduce (mean) of a single flat fp32 gradient buffer (4.04 M parameters = 16.2 MB for the FiLM model) over NCCL /
NVLink - followed by gradient clipping and Adam as one fused kernel over the same flat buffers
(src/trainers/trainer.py:42-43,115-116; scripts/train/config_tss.yaml:36-39,59).

The reference has no distributed code (single process, single GPU); this is new functionality next to the path.  The
backward that fills the gradient buffer is train.py (DPRNN-TasNet and every DPRNN-Spe fusion); SpeTrainStep there drives
one iteration with the pieces of this module: FlatParams, broadcast_state (replicas start identical), allreduce_mean (the
exchange step) and ClipAdam.
"""
from __future__ import annotations

import torch

from ._lib import lib


class FlatParams:
    """The trainable parameters of a module as views into one contiguous fp32 buffer (plus a same-shaped gradient
    buffer), in ``named_parameters()`` order restricted to ``requires_grad`` - the set the reference hands to Adam
    (src/trainers/trainer.py:42-43: the frozen ``separation.average.*`` of the attention fusion are excluded)."""

    ALIGN = 64                                                      # elements

    def __init__(self, module: torch.nn.Module):
        self.named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if not self.named:
            raise ValueError('no trainable parameters')
        dev = self.named[0][1].device
        self.numel = sum(p.numel() for _, p in self.named)
        # every parameter starts on a 256-byte boundary (the kernels read weights with 16-byte vector loads); the
        # padding stays zero in both buffers, so norms, the all-reduce and Adam are unaffected by it
        pad = lambda n: (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.size = sum(pad(p.numel()) for _, p in self.named)
        self.flat = torch.zeros(self.size, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.size, dtype=torch.float32, device=dev)
        off = 0
        self.offsets = []
        for _, p in self.named:
            n = p.numel()
            self.offsets.append(off)
            self.flat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)              # parameters now alias the flat buffer
            p.grad = self.grad[off:off + n].view_as(p)
            off += pad(n)

    def zero_grad(self):
        self.grad.zero_()


def allreduce_mean(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """The exchange step: in-place mean over the ranks of the flat gradient buffer (one collective per iteration;
    NCCL on GPU tensors, gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.div_(dist.get_world_size(group))
    return flat_grad


def _world(group=None) -> int:
    import torch.distributed as dist
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def broadcast_state(tensors, group=None, src: int = 0):
    """In-place broadcast of rank `src`'s tensors (parameters, optimiser moments, BatchNorm buffers) to every replica;
    no-op without a process group.  Integer buffers (num_batches_tracked) are broadcast as they are."""
    import torch.distributed as dist
    if _world(group) > 1:
        for t in tensors:
            dist.broadcast(t, src=src, group=group)
    return tensors


def average_buffers(tensors, group=None):
    """In-place mean over the ranks (the BatchNorm running statistics before a rank-0 checkpoint)."""
    import torch.distributed as dist
    w = _world(group)
    if w > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            t.div_(w)
    return tensors


class ClipAdam:
    """clip_grad_norm_(max_norm) + Adam(lr, betas, eps, weight_decay) over FlatParams, one fused pass on the GPU."""

    def __init__(self, fp: FlatParams, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5, max_norm=5.0):
        if not fp.flat.is_cuda:
            raise RuntimeError('ClipAdam runs on the GPU (no CPU path)')
        self.fp, self.lr, self.betas, self.eps, self.wd, self.max_norm = fp, lr, betas, eps, weight_decay, max_norm
        self.exp_avg = torch.zeros_like(fp.flat)
        self.exp_avg_sq = torch.zeros_like(fp.flat)
        self.step_count = 0
        self.ws = torch.empty(lib().query('dprnn_clip_adam_workspace_bytes'), dtype=torch.uint8, device=fp.flat.device)
        self.total_norm = torch.zeros(1, device=fp.flat.device)

    def step(self):
        self.step_count += 1
        lib().call('dprnn_clip_adam_step', self.fp.flat, self.fp.grad, self.exp_avg, self.exp_avg_sq, self.fp.size,
                   float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.wd),
                   float(self.max_norm), self.step_count, self.ws, self.total_norm,
                   torch.cuda.current_stream().cuda_stream)
        return self.total_norm

    # ---- torch.optim.Adam.state_dict() layout, so that the reference's Trainer ({'epoch','optimizer','model'} checkpoints,
    # src/trainers/trainer.py:294-306) and this stepper can resume from each other's files
    def state_dict(self):
        state = {}
        if self.step_count > 0:
            for i, ((_, p), off) in enumerate(zip(self.fp.named, self.fp.offsets)):
                n = p.numel()
                state[i] = {'step': torch.tensor(float(self.step_count)),
                            'exp_avg': self.exp_avg[off:off + n].view_as(p).clone(),
                            'exp_avg_sq': self.exp_avg_sq[off:off + n].view_as(p).clone()}
        group = {'lr': self.lr, 'betas': tuple(self.betas), 'eps': self.eps, 'weight_decay': self.wd, 'amsgrad': False,
                 'maximize': False, 'foreach': None, 'capturable': False, 'differentiable': False, 'fused': None,
                 'decoupled_weight_decay': False, 'params': list(range(len(self.fp.named)))}
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        group = sd['param_groups'][0]
        if len(group['params']) != len(self.fp.named):
            raise ValueError(f"optimizer state has {len(group['params'])} parameters, the model has {len(self.fp.named)}")
        if group.get('amsgrad') or group.get('maximize'):
            raise NotImplementedError('amsgrad / maximize are not built (the reference trains with plain Adam)')
        self.lr, self.betas, self.eps, self.wd = group['lr'], tuple(group['betas']), group['eps'], group['weight_decay']
        self.exp_avg.zero_(); self.exp_avg_sq.zero_()
        self.step_count = 0
        for i, ((_, p), off) in enumerate(zip(self.fp.named, self.fp.offsets)):
            st = sd['state'].get(group['params'][i])
            if st is None:
                continue
            n = p.numel()
            self.exp_avg[off:off + n].copy_(st['exp_avg'].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
            self.step_count = max(self.step_count, int(float(st['step'])))
