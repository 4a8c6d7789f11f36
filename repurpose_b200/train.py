"""First pieces of the training step (BASELINE configs[4]; reference main.py:294-409): the
memory-bound kernels around the (not yet written) GEMM / attention backward, behind small Python mirrors.

  focal_loss_grad(masks, logits, labels, batch_size)  <- autograd of `model.losses(...)['cls_loss'] / batch_size`
  layernorm512_backward(x, dy, gamma)                 <- autograd of nn.LayerNorm(512)
  FlatAdam(params, lr, weight_decay)                  <- optim.Adam(model.parameters(), lr, weight_decay).step()
  allreduce_flat_(flat_grads, group)                  <- DDP's gradient averaging (utils/distributed.py:415-428):
                                                         ONE all-reduce over the flat gradient buffer

All compute is in librepurpose_b200.so; there is no PyTorch fallback.  What a full step still needs is listed
in DESIGN.md §7 (dgrad / wgrad GEMMs with MN-major operands, attention backward, dropout)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, cur_stream, ptr


def _cuda(t, what):
    if not t.is_cuda:
        raise _lib.RepurposeError(f"{what}: CUDA tensors only (no CPU path)")


def focal_loss_grad(masks, out_cls_logits, gt_cls_labels, batch_size: int = 1, alpha: float = 0.7, gamma: float = 2.0):
    """d(cls_loss / batch_size) / d(out_cls_logits) for `MMCTransformer.losses` (models/MMCTransformer.py:159-179,
    models/losses.py:5-53, main.py:326): masks [B,1,T], logits [B,T,1], labels [B,T] -> [B,T,1] fp32."""
    _cuda(out_cls_logits, "focal_loss_grad")
    dev = out_cls_logits.device
    logits = out_cls_logits.to(torch.float32).reshape(-1).contiguous()
    targets = gt_cls_labels.to(dev, torch.float32).reshape(-1).contiguous()
    mask = masks.to(dev).reshape(-1).ne(0).to(torch.uint8).contiguous()
    if not (logits.numel() == targets.numel() == mask.numel()):
        raise ValueError("focal_loss_grad: logits [B,T,1], labels [B,T] and masks [B,1,T] must cover the same steps")
    out = torch.empty_like(logits)
    with torch.cuda.device(dev):
        check(_lib.load().rp_focal_loss_grad(ptr(logits), ptr(targets), ptr(mask), logits.numel(), alpha, gamma,
                                             1.0 / float(batch_size), ptr(out), cur_stream()), "rp_focal_loss_grad")
    return out.view_as(out_cls_logits)


def layernorm512_backward(x, dy, gamma, eps: float = 1e-5):
    """x, dy [..., 512] fp32, gamma [512] -> (dx like x, dgamma [512], dbeta [512]); deterministic."""
    _cuda(x, "layernorm512_backward")
    if x.shape[-1] != 512 or dy.shape != x.shape:
        raise ValueError("layernorm512_backward: x and dy must be [..., 512]")
    dev = x.device
    x2 = x.to(torch.float32).reshape(-1, 512).contiguous()
    dy2 = dy.to(torch.float32).reshape(-1, 512).contiguous()
    g = gamma.to(dev, torch.float32).contiguous()
    dx = torch.empty_like(x2)
    dgamma = torch.empty(512, dtype=torch.float32, device=dev)
    dbeta = torch.empty(512, dtype=torch.float32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        nbytes = int(lib.rp_layernorm512_bwd_scratch_bytes())
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib.rp_layernorm512_bwd(ptr(x2), ptr(dy2), ptr(g), x2.shape[0], eps, ptr(dx), ptr(dgamma), ptr(dbeta),
                                      ptr(scratch), nbytes, cur_stream()), "rp_layernorm512_bwd")
    return dx.view_as(x), dgamma, dbeta


class FlatAdam:
    """torch.optim.Adam (L2 weight decay, no amsgrad) over ONE flat fp32 buffer that the parameters are views
    of — the layout the single gradient all-reduce wants too.  `params`: iterable of CUDA fp32 tensors; after
    construction each `p.data` is a view into `self.flat` and `self.grads_like(p)` is its slice of `self.grad`."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, bf16_copy=False):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        dev = self.params[0].device
        _cuda(self.params[0], "FlatAdam")
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_bf16 = torch.empty(n, dtype=torch.bfloat16, device=dev) if bf16_copy else None
        self._slices = []
        pos = 0
        for p in self.params:
            k = p.numel()
            self.flat[pos:pos + k].copy_(p.data.reshape(-1))
            p.data = self.flat[pos:pos + k].view_as(p.data)
            self._slices.append((pos, k))
            pos += k
        self.lr, self.betas, self.eps, self.weight_decay, self.step_count = lr, betas, eps, weight_decay, 0

    def grads_like(self, i: int):
        pos, k = self._slices[i]
        return self.grad[pos:pos + k].view_as(self.params[i].data)

    def zero_grad(self):
        self.grad.zero_()

    def step(self):
        self.step_count += 1
        with torch.cuda.device(self.flat.device):
            check(_lib.load().rp_adam_step(ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                           self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                                           self.weight_decay, self.step_count, ptr(self.flat_bf16), cur_stream()),
                  "rp_adam_step")


def allreduce_flat_(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """Average the flat gradient buffer over the data-parallel ranks with ONE collective (what DDP does in
    buckets, utils/distributed.py:415-428).  No-op outside a process group."""
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
            flat_grad.div_(world)
    return flat_grad
