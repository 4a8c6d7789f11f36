"""First pieces of the training step: focal-loss gradient against the reference's own autograd
(golden), LayerNorm backward and Adam against torch (the reference's third-party arithmetic for these:
nn.LayerNorm, optim.Adam — main.py:190-191), and the flat-buffer gradient all-reduce on two gloo ranks."""
from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden" / "losses_cases.npz"
DEV = "cuda"


def test_focal_grad_oracle_matches_reference_autograd():
    from oracle import losses as ol
    g = np.load(GOLD)
    for name in g["names"]:
        masks, logits, labels = (torch.from_numpy(g[f"{name}_{k}"]) for k in ("masks", "logits", "labels"))
        d = ol.losses_grad(masks, logits, labels, logits.shape[0])
        assert (d - torch.from_numpy(g[f"{name}_dlogits"])).abs().max().item() < 1e-6, name


@pytest.mark.gpu
def test_focal_loss_grad_matches_reference_autograd_golden():
    from repurpose_b200.train import focal_loss_grad
    g = np.load(GOLD)
    for name in g["names"]:
        masks, logits, labels = (torch.from_numpy(g[f"{name}_{k}"]).to(DEV) for k in ("masks", "logits", "labels"))
        got = focal_loss_grad(masks, logits, labels, batch_size=logits.shape[0])
        ref = torch.from_numpy(g[f"{name}_dlogits"])
        assert got.shape == logits.shape
        err = (got.cpu() - ref).abs().max().item()
        assert err <= 2e-6 * max(1.0, ref.abs().max().item()), (name, err)
        assert (got.cpu()[~masks.cpu().transpose(1, 2)] == 0).all()          # padded steps get no gradient


@pytest.mark.gpu
@pytest.mark.parametrize("M", [1, 7, 300, 4099, 57632])
def test_layernorm512_backward_matches_autograd(M):
    from repurpose_b200.train import layernorm512_backward
    gen = torch.Generator(device=DEV).manual_seed(M)
    x = (torch.randn(M, 512, device=DEV, generator=gen) * 2 + torch.randn(M, 1, device=DEV, generator=gen)).requires_grad_(True)
    gamma = (torch.rand(512, device=DEV, generator=gen) + 0.5).requires_grad_(True)
    beta = torch.randn(512, device=DEV, generator=gen).requires_grad_(True)
    dy = torch.randn(M, 512, device=DEV, generator=gen)
    torch.nn.functional.layer_norm(x, (512,), gamma, beta, 1e-5).backward(dy)
    dx, dgamma, dbeta = layernorm512_backward(x.detach(), dy, gamma.detach())
    assert torch.allclose(dx, x.grad, atol=2e-5, rtol=1e-4), (dx - x.grad).abs().max().item()
    tol = 2e-4 * max(1.0, M ** 0.5)                                            # fp32 sums over M rows
    assert torch.allclose(dgamma, gamma.grad, atol=tol, rtol=1e-4), (dgamma - gamma.grad).abs().max().item()
    assert torch.allclose(dbeta, beta.grad, atol=tol, rtol=1e-4), (dbeta - beta.grad).abs().max().item()
    dx2, dgamma2, dbeta2 = layernorm512_backward(x.detach(), dy, gamma.detach())
    assert torch.equal(dgamma, dgamma2) and torch.equal(dbeta, dbeta2) and torch.equal(dx, dx2)   # deterministic


@pytest.mark.gpu
def test_flat_adam_matches_torch_adam():
    from repurpose_b200.train import FlatAdam
    torch.manual_seed(0)
    shapes = [(512, 2944), (512,), (1536, 512), (3,), (1, 256)]
    ref_params = [torch.nn.Parameter(torch.randn(*s, device=DEV)) for s in shapes]
    our_params = [torch.nn.Parameter(p.detach().clone()) for p in ref_params]
    ref = torch.optim.Adam(ref_params, lr=1e-3, weight_decay=1e-4)           # main.py:190-191, configs/Repurpose.yaml:36-38
    ours = FlatAdam(our_params, lr=1e-3, weight_decay=1e-4, bf16_copy=True)
    for step in range(5):
        for i, (rp, s) in enumerate(zip(ref_params, shapes)):
            g = torch.randn(*s, device=DEV) * (0.1 + step)
            rp.grad = g.clone()
            ours.grads_like(i).copy_(g)
        ref.step()
        ours.step()
        for rp, op in zip(ref_params, our_params):
            assert torch.allclose(op.data, rp.data, atol=1e-6, rtol=1e-5), (step, (op.data - rp.data).abs().max().item())
    assert torch.equal(ours.flat_bf16, ours.flat.to(torch.bfloat16))
    assert our_params[0].data.data_ptr() == ours.flat.data_ptr()              # parameters are views of the flat buffer


def _allreduce_worker(rank, world, port, out):
    import torch.distributed as dist
    from repurpose_b200.train import allreduce_flat_
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    allreduce_flat_(g)
    out.put((rank, g[:4].tolist(), float(g.sum())))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + int(torch.randint(0, 300, (1,)))
    procs = [ctx.Process(target=_allreduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join()
    want = (torch.arange(1000, dtype=torch.float32) * 1.5)
    for _, head, total in res:
        assert head == want[:4].tolist() and total == float(want.sum())
