"""The training step (BASELINE configs[4]; reference main.py:294-409) behind small Python mirrors:

  TrainStep(model, lr, weight_decay).step(batch)      <- one iteration of main.py's loop: forward (train graph), masked
                                                         focal loss / batch_size, backward, gradient all-reduce, Adam

  focal_loss_grad(masks, logits, labels, batch_size)  <- autograd of `model.losses(...)['cls_loss'] / batch_size`
  layernorm512_backward(x, dy, gamma)                 <- autograd of nn.LayerNorm(512)
  FlatAdam(params, lr, weight_decay)                  <- optim.Adam(model.parameters(), lr, weight_decay).step()
  allreduce_flat_(flat_grads, group)                  <- DDP's gradient averaging (utils/distributed.py:415-428):
                                                         ONE all-reduce over the flat gradient buffer

All compute is in librepurpose_b200.so; there is no PyTorch fallback.  Dropout: the reference trains under
model.train() (main.py:285) with nn.Dropout(0.1) in the encoder layers, on the attention weights and in the heads;
`TrainStep(dropout=0.1)` applies all of them with counter-based masks (csrc/dropout.cuh; `dropout_keys` below derives a
stream per site and step).  The draws differ from torch's Philox stream, the distribution does not; the parity tests
feed these masks to autograd over the reference graph."""
from __future__ import annotations

import torch
import torch.distributed as dist

import ctypes

from . import _lib
from ._lib import check, cur_stream, ptr

_M32 = 0xFFFFFFFF


def _fmix32(x: int) -> int:
    """MurmurHash3's 32-bit finaliser (the device's drop_hash uses the same rounds)"""
    x &= _M32
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & _M32
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & _M32
    x ^= x >> 16
    return x


def dropout_keys(seed: int, step: int, site: int):
    """(key_a, key_b) of one dropout site at one training step: every (seed, step, site) draws from its own stream"""
    a = _fmix32((seed & _M32) ^ _fmix32(site * 0x9E3779B1 + step * 0x632BE5AB + 1))
    b = _fmix32(((seed >> 32) & _M32) ^ _fmix32(site * 0x85EBCA6B + step * 0x27D4EB2F + 2))
    return a, b


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


LOG2E_OVER_8 = 1.4426950408889634 / 8.0


class TrainStep:
    """One training iteration of the reference loop (main.py:318-368) on the device:

        output = model(batch); loss = model.losses(*output)['cls_loss'] / batch_size
        optimizer.zero_grad(); loss.backward(); optimizer.step()          (+ DDP's gradient averaging)

    `model` is this package's MMCTransformer (reference state-dict schema).  Its parameters become views of one flat
    fp32 master buffer (FlatAdam); a bf16 copy feeds the tensor-core GEMMs.  The forward keeps, per encoder layer, the
    residual stream before each LayerNorm (fp32), the LayerNorm outputs, q|k|v, the attention output and its
    log-sum-exp, and the ReLU output (bf16): 14.3 KB per token and layer.  The backward is autograd's chain written
    out with rp_gemm_bwd / rp_fmha_bwd / rp_layernorm512_bwd_acc / rp_colsum_bf16 / rp_relu_bwd / rp_head_out_bwd.
    reg_head receives no gradient in the reference (the loss is the focal classification term only,
    models/MMCTransformer.py:159-179) — torch's Adam skips such parameters, and so does this step."""

    def __init__(self, model, lr=1e-3, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, group=None, dropout=0.1,
                 seed=0, deterministic=None):
        """dropout: the p of every nn.Dropout of the reference graph (0.1 there; 0 = the eval-mode graph);
        seed: base of the dropout streams (use a different one per data-parallel rank, as torch's per-process
        generators are);
        deterministic: True = bit-reproducible gradients (the attention backward runs its two deterministic kernels,
        about 20 % slower), False = the fused attention backward whose dQ is summed by fp32 adds in L2 in no fixed order,
        None = leave the process-wide setting (rp_set_attn_bwd_deterministic / RP_FMHA_BWD_FUSED) as it is"""
        if not 0.0 <= dropout < 1.0:
            raise ValueError("TrainStep: 0 <= dropout < 1")
        if deterministic is not None:
            check(_lib.load().rp_set_attn_bwd_deterministic(1 if deterministic else 0), "rp_set_attn_bwd_deterministic")
        self.p, self.seed, self.step_index = float(dropout), int(seed), 0
        self.model = model
        self.cfg = c = model._cfg
        if c["d_model"] != 512 or c["head_hidden"] != 256 or c["num_heads"] * 64 != c["d_model"]:
            raise ValueError("TrainStep: kernels are specialised for d_model 512, heads of 64, head hidden 256")
        named = [(n, p) for n, p in model.named_parameters()]
        _cuda(named[0][1], "TrainStep")
        self.dev = named[0][1].device
        trainable = [(n, p) for n, p in named if not n.startswith("reg_head.")]
        self.opt = FlatAdam([p for _, p in trainable], lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                            bf16_copy=True)
        self._index = {n: i for i, (n, _) in enumerate(trainable)}
        self._frozen16 = {n: p.data.to(torch.bfloat16) for n, p in named if n.startswith("reg_head.") and p.dim() == 2}
        self._frozen32 = {n: p.data for n, p in named if n.startswith("reg_head.")}
        self.group = group
        # bucketed, overlapped gradient all-reduce (what DDP's buckets do, utils/distributed.py:415-428): the flat buffer is in
        # parameter order and the backward pass fills it from the end, so once the backward of layer l is through, everything
        # from that layer's first parameter to the end of the (not yet reduced) buffer is final and can travel while the
        # earlier layers still compute.  Cuts every four layers; the head of the buffer goes last, in step().
        self.overlap_allreduce = True
        self._layer_start = {}
        for n, i in self._index.items():
            if n.startswith("multimodal_encoder.layers."):
                l = int(n.split(".")[2])
                pos = self.opt._slices[i][0]
                self._layer_start[l] = min(self._layer_start.get(l, pos), pos)
        self._ar_pending, self._ar_hi = [], None
        self.lib = _lib.load()
        self._bufs = None
        self._scratch = torch.empty(int(self.lib.rp_train_scratch_bytes()), dtype=torch.uint8, device=self.dev)
        self._part = torch.empty(8 << 20, dtype=torch.float32, device=self.dev)  # split-K partials (32 MB)
        L = c["num_layers"]
        self._wqkv16 = [torch.empty(1536, 512, dtype=torch.bfloat16, device=self.dev) for _ in range(L)]
        self._bqkv = [torch.empty(1536, dtype=torch.float32, device=self.dev) for _ in range(L)]
        self.refresh_weights()

    # ---- parameter views -------------------------------------------------------------------------------------
    def w32(self, name):
        return self._frozen32[name] if name in self._frozen32 else self.opt.params[self._index[name]].data

    def w16(self, name):
        if name in self._frozen16:
            return self._frozen16[name]
        pos, k = self.opt._slices[self._index[name]]
        return self.opt.flat_bf16[pos:pos + k].view_as(self.opt.params[self._index[name]].data)

    def grad(self, name):
        return self.opt.grads_like(self._index[name])

    def refresh_weights(self):
        """bf16 operands after a parameter update: the flat copy, and in_proj with the q rows scaled by log2(e)/8."""
        st, lib = cur_stream(), self.lib
        with torch.cuda.device(self.dev):
            check(lib.rp_cast_scaled(ptr(self.opt.flat), self.opt.flat.numel(), 0, 1.0, ptr(self.opt.flat_bf16), 0, st),
                  "rp_cast_scaled")
            for l in range(self.cfg["num_layers"]):
                pre = f"multimodal_encoder.layers.{l}.self_attn."
                check(lib.rp_cast_scaled(ptr(self.w32(pre + "in_proj_weight")), 1536 * 512, 512 * 512, LOG2E_OVER_8,
                                         ptr(self._wqkv16[l]), 0, st), "rp_cast_scaled")
                check(lib.rp_cast_scaled(ptr(self.w32(pre + "in_proj_bias")), 1536, 512, LOG2E_OVER_8, 0,
                                         ptr(self._bqkv[l]), st), "rp_cast_scaled")

    # ---- buffers -----------------------------------------------------------------------------------------------
    def _buffers(self, B, T):
        if self._bufs is not None and self._bufs["shape"] == (B, T):
            return self._bufs
        c, dev, M, L = self.cfg, self.dev, B * T, self.cfg["num_layers"]
        f32 = dict(dtype=torch.float32, device=dev)
        b16 = dict(dtype=torch.bfloat16, device=dev)
        Cin = c["vis_dim"] + c["aud_dim"] + c["text_dim"]
        d = dict(shape=(B, T))
        d["xcat"] = torch.empty(M, Cin, **b16)
        d["xproj"] = torch.empty(M, 512, **f32)
        d["h"] = [torch.empty(M, 512, **f32) for _ in range(L + 1)]
        d["hmid"] = [torch.empty(M, 512, **f32) for _ in range(L)]
        d["u1"] = [torch.empty(M, 512, **b16) for _ in range(L + 1)]   # u1[L] = encoder_norm output
        d["u2"] = [torch.empty(M, 512, **b16) for _ in range(L)]
        d["qkv"] = [torch.empty(M, 1536, **b16) for _ in range(L)]
        d["attn"] = [torch.empty(M, 512, **b16) for _ in range(L)]
        d["ffn"] = [torch.empty(M, c["d_ff"], **b16) for _ in range(L)]
        d["lse"] = [torch.empty(B, c["num_heads"], T, **f32) for _ in range(L)]
        d["fm"] = torch.empty(M, 512, **f32)
        d["feats"] = torch.empty(B, T, 512, **f32)
        d["uc"], d["ur"] = torch.empty(M, 512, **b16), torch.empty(M, 512, **b16)
        d["a1c"], d["a2c"] = torch.empty(M, 256, **b16), torch.empty(M, 256, **b16)
        d["a1r"], d["a2r"] = torch.empty(M, 256, **b16), torch.empty(M, 256, **b16)
        d["logits"] = torch.empty(B, T, 1, **f32)
        d["offsets"] = torch.empty(B, T, 2, **f32)
        # backward
        d["dh"], d["dh16"] = torch.empty(M, 512, **f32), torch.empty(M, 512, **b16)
        d["du"] = torch.empty(M, 512, **f32)                      # fp32 gradient entering a LayerNorm
        d["dwide"] = torch.empty(M, c["d_ff"], **b16)             # dffn / dqkv / head activations' gradients
        d["d512"] = torch.empty(M, 512, **b16)                    # dattn
        d["dsum"] = torch.empty(B, c["num_heads"], T, **f32)
        d["lens"] = torch.empty(B, dtype=torch.int32, device=dev)
        if self.p > 0:
            # keep bits of the attention-weight dropout, one buffer per layer (the backward reads them again): rows of
            # round_up(T, 128) / 32 words per (batch, head, query) — 55 MB per layer at B = 16, T = 1801
            d["bits_ld"] = ((T + 127) // 128) * 4
            d["bits"] = [torch.empty(B * c["num_heads"] * T, d["bits_ld"], dtype=torch.int32, device=dev) for _ in range(L)]
        self._bufs = d
        return d

    # ---- dropout sites -----------------------------------------------------------------------------------------------
    def site(self, name, layer=0):
        """stream index of a dropout site: 'attn' / 'drop1' / 'ffn' / 'drop2' of an encoder layer, or 'feats',
        'cls1', 'cls2', 'reg1', 'reg2' (feature_map[3], cls_head[3] / [6], reg_head[3] / [6])"""
        L = self.cfg["num_layers"]
        per_layer = {"attn": 0, "drop1": 1, "ffn": 2, "drop2": 3}
        if name in per_layer:
            return 4 * layer + per_layer[name]
        return 4 * L + ("feats", "cls1", "cls2", "reg1", "reg2").index(name)

    def _drop(self, site):
        """rp_dropout of a site at the current step (None when dropout is off)"""
        if self.p <= 0 or site is None:
            return None
        a, b = dropout_keys(self.seed, self.step_index, site)
        return _lib.RpDropout(a, b, self.p)

    @property
    def drop_scale(self):
        return 1.0 / (1.0 - self.p) if self.p > 0 else 1.0

    def keep_mask(self, site, shape):
        """the keep mask (bool, `shape`) of an element-wise site at the current step — what the kernels apply"""
        n = 1
        for k in shape:
            n *= int(k)
        out = torch.empty(n, dtype=torch.uint8, device=self.dev)
        with torch.cuda.device(self.dev):
            check(self.lib.rp_dropout_mask_u8(ctypes.byref(self._drop(site)), n, ptr(out), cur_stream()), "rp_dropout_mask_u8")
        return out.view(*shape).bool()

    def attention_keep_mask(self, layer):
        """[B, H, T, T] bool keep mask of layer `layer`'s attention weights, unpacked from the bits of the last forward"""
        d = self._bufs
        B, T = d["shape"]
        H = self.cfg["num_heads"]
        words = d["bits"][layer].view(B, H, T, d["bits_ld"]).to(torch.int64) & _M32
        shifts = torch.arange(32, device=self.dev, dtype=torch.int64)
        bits = ((words[..., None] >> shifts) & 1).reshape(B, H, T, d["bits_ld"] * 32)
        return bits[..., :T].bool()

    # ---- per-kernel-class device timing (CUDA events on the launch stream) ----------------------------------------
    _PROFILED = {"_gemm": "fwd_gemm", "_ln": "fwd_layernorm", "_dgrad": "bwd_dgrad", "_wgrad": "bwd_wgrad",
                 "_colsum": "bwd_bias_colsum", "_ln_bwd": "bwd_layernorm", "_relu_bwd": "bwd_relu",
                 "_relu_bwd_bias": "bwd_relu"}

    def profile_begin(self):
        """from now on every wrapped launch is bracketed by CUDA events; `profile_end()` -> {tag: (ms, launches)}"""
        self._prof = []
        for meth, tag in self._PROFILED.items():
            inner = getattr(type(self), meth)

            def wrapped(*a, _inner=inner, _tag=tag, **k):
                with self._timed(_tag):
                    return _inner(self, *a, **k)
            setattr(self, meth, wrapped)

    def _timed(self, tag):
        import contextlib

        @contextlib.contextmanager
        def cm():
            if getattr(self, "_prof", None) is None:
                yield
                return
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            yield
            e1.record()
            self._prof.append((tag, e0, e1))
        return cm()

    def profile_end(self):
        torch.cuda.synchronize(self.dev)
        out = {}
        for tag, e0, e1 in self._prof or []:
            ms, n = out.get(tag, (0.0, 0))
            out[tag] = (ms + e0.elapsed_time(e1), n + 1)
        self._prof = None
        for meth in self._PROFILED:
            if meth in self.__dict__:
                delattr(self, meth)
        return out

    # ---- thin wrappers over the C ABI ------------------------------------------------------------------------------
    def _gemm(self, epi, A, W, D, bias, resid=None, site=None):
        """site: the nn.Dropout that follows this Linear (after its ReLU / before the residual add) in train mode"""
        M, K = A.shape
        N = W.shape[0]
        drop = self._drop(site)
        if drop is not None:
            check(self.lib.rp_gemm_bf16_dropout(epi, ptr(A), K, ptr(W), K, ptr(D), N, ptr(bias), ptr(resid),
                                                N if resid is not None else 0, M, N, K, ctypes.byref(drop), cur_stream()),
                  "rp_gemm_bf16_dropout")
            return
        check(self.lib.rp_gemm_bf16(epi, ptr(A), K, ptr(W), K, ptr(D), N, ptr(bias), ptr(resid), N if resid is not None else 0,
                                    M, N, K, cur_stream()), "rp_gemm_bf16")

    def _ln(self, mode, x, M, T, g0, b0, g1=None, b1=None, g2=None, b2=None, pe=None, out_f32=None, y=None, y2=None,
            site=None):
        drop = self._drop(site)
        if drop is not None:
            check(self.lib.rp_layernorm512_dropout(mode, ptr(x), M, T, ptr(g0), ptr(b0), ptr(g1), ptr(b1), ptr(g2), ptr(b2),
                                                   ptr(pe), ptr(out_f32), ptr(y), ptr(y2), ctypes.byref(drop), cur_stream()),
                  "rp_layernorm512_dropout")
            return
        check(self.lib.rp_layernorm512(mode, ptr(x), M, T, ptr(g0), ptr(b0), ptr(g1), ptr(b1), ptr(g2), ptr(b2), ptr(pe),
                                       ptr(out_f32), ptr(y), ptr(y2), cur_stream()), "rp_layernorm512")

    def _dgrad(self, dY, W, out):
        """out[M, in] = dY[M, out] W[out, in]"""
        M, Nout = dY.shape
        Kin = W.shape[1]
        check(self.lib.rp_gemm_bwd(1, 1 if out.dtype == torch.float32 else 0, ptr(dY), Nout, ptr(W), Kin, ptr(out), Kin,
                                   M, Kin, Nout, 1, cur_stream()), "rp_gemm_bwd dgrad")

    def _wgrad(self, dY, X, name):
        """grad(name)[out, in] = dY[tok, out]^T X[tok, in];  grad(name bias)[out] = column sums of dY"""
        tok, Nout = dY.shape
        Kin = X.shape[1]
        gw = self.grad(name + "weight") if name.endswith(".") else self.grad(name)
        tiles = ((Nout + 255) // 256) * ((Kin + 255) // 256)
        kblocks = (tok + 63) // 64
        splits = max(1, min(kblocks, 74 // tiles, self._part.numel() // (Nout * Kin)))
        per = (kblocks + splits - 1) // splits
        splits = (kblocks + per - 1) // per          # no empty split
        st = cur_stream()
        if splits == 1:
            check(self.lib.rp_gemm_bwd(3, 1, ptr(dY), Nout, ptr(X), Kin, ptr(gw), Kin, Nout, Kin, tok, 1, st), "rp_gemm_bwd wgrad")
        else:
            check(self.lib.rp_gemm_bwd(3, 1, ptr(dY), Nout, ptr(X), Kin, ptr(self._part), Kin, Nout, Kin, tok, splits, st),
                  "rp_gemm_bwd wgrad")
            check(self.lib.rp_splitk_reduce(ptr(self._part), splits, Nout * Kin, ptr(gw), st), "rp_splitk_reduce")

    def _colsum(self, X, out):
        M, N = X.shape
        check(self.lib.rp_colsum_bf16(ptr(X), M, N, ptr(out), ptr(self._scratch), self._scratch.numel(), cur_stream()),
              "rp_colsum_bf16")

    def _ln_bwd(self, x, dy, prefix, dh, dh16, accumulate, bias_grad=None, site=None):
        """dh (+)= LayerNorm backward of the branch `prefix` (…weight / …bias receive their gradients); dh16 = bf16 copy
        of the resulting dh; bias_grad (a parameter name) receives its column sums — the bias gradient of the Linear
        whose output gradient dh is.  site: that Linear's output passed through this nn.Dropout before it joined the
        residual stream (dropout1 / dropout2): dh16 and the bias gradient then carry dropout's backward of dh"""
        M = x.shape[0] if x.dim() == 2 else x.numel() // 512
        drop = self._drop(site)
        args = (ptr(x), ptr(dy), ptr(self.w32(prefix + "weight")), M, 1e-5, 1 if accumulate else 0, ptr(dh), ptr(dh16),
                ptr(self.grad(bias_grad)) if bias_grad else 0, ptr(self.grad(prefix + "weight")),
                ptr(self.grad(prefix + "bias")), ptr(self._scratch), self._scratch.numel())
        if drop is not None:
            check(self.lib.rp_layernorm512_bwd_acc_dropout(*args, ctypes.byref(drop), cur_stream()),
                  "rp_layernorm512_bwd_acc_dropout")
            return
        check(self.lib.rp_layernorm512_bwd_acc(*args, cur_stream()), "rp_layernorm512_bwd_acc")

    def _relu_bwd_bias(self, dy, act, bias_name):
        """dy = act > 0 ? dy * drop_scale : 0 in place (act = dropout(relu(.)) as stored: the ReLU and the Dropout behind it
        in one mask), and grad(bias_name) = column sums of the result"""
        M, N = dy.shape
        check(self.lib.rp_relu_bwd_colsum_scaled(ptr(dy), ptr(act), M, N, self.drop_scale, ptr(self.grad(bias_name)),
                                                 ptr(self._scratch), self._scratch.numel(), cur_stream()),
              "rp_relu_bwd_colsum_scaled")

    def _relu_bwd(self, dy, act):
        check(self.lib.rp_relu_bwd_scaled(ptr(dy), ptr(act), dy.numel(), 1 if dy.dtype == torch.float32 else 0,
                                          self.drop_scale, cur_stream()), "rp_relu_bwd_scaled")

    # ---- forward (train graph, activations kept) -------------------------------------------------------------------
    def forward(self, batch):
        """-> the reference's 6-tuple (models/MMCTransformer.py:109-151), computed by the same kernels as inference but
        with every tensor the backward needs kept (no in-place residual updates, no LayerNorm-in-GEMM fusion)."""
        m, c, lib = self.model, self.cfg, self.lib
        vis, aud, txt = (batch[k].to(self.dev, torch.float32).contiguous() for k in ("visual_feats", "audio_feats", "text_feats"))
        B, T = vis.shape[0], vis.shape[1]
        M, L, H = B * T, c["num_layers"], c["num_heads"]
        d = self._buffers(B, T)
        masks = batch["masks"].to(self.dev)
        with torch.cuda.device(self.dev):
            st = cur_stream()
            lens = m._lens_from_masks(masks, B, T)
            d["lens"].copy_(lens)
            check(lib.rp_concat_cast(ptr(vis), ptr(aud), ptr(txt), c["vis_dim"], c["aud_dim"], c["text_dim"], ptr(d["xcat"]), M, st),
                  "rp_concat_cast")
            self._gemm(2, d["xcat"], self.w16("input_projection.weight"), d["xproj"], self.w32("input_projection.bias"))
            pe = m.positional_encoding.pe[0]
            lay = "multimodal_encoder.layers.{}."
            self._ln(1, d["xproj"], M, T, self.w32("input_norm.weight"), self.w32("input_norm.bias"),
                     self.w32(lay.format(0) + "norm1.weight"), self.w32(lay.format(0) + "norm1.bias"), pe=pe,
                     out_f32=d["h"][0], y=d["u1"][0])
            for l in range(L):
                p = lay.format(l)
                self._gemm(0, d["u1"][l], self._wqkv16[l], d["qkv"][l], self._bqkv[l])
                q = d["qkv"][l]
                if self.p > 0:
                    with self._timed("dropout_bits"):
                        check(lib.rp_attn_dropout_bits(ctypes.byref(self._drop(self.site("attn", l))), d["bits"][l].numel(),
                                                       ptr(d["bits"][l]), st), "rp_attn_dropout_bits")
                with self._timed("fwd_fmha"):
                    if self.p > 0:
                        check(lib.rp_fmha_train_dropout(ptr(q), ptr(q) + 1024, ptr(q) + 2048, ptr(d["attn"][l]), 1536, 512,
                                                        B, H, T, ptr(d["lens"]), ptr(d["lse"][l]), ptr(d["bits"][l]),
                                                        d["bits_ld"], self.p, st), "rp_fmha_train_dropout")
                    else:
                        check(lib.rp_fmha_train(ptr(q), ptr(q) + 1024, ptr(q) + 2048, ptr(d["attn"][l]), 1536, 512, B, H, T,
                                                ptr(d["lens"]), ptr(d["lse"][l]), st), "rp_fmha_train")
                self._gemm(3, d["attn"][l], self.w16(p + "self_attn.out_proj.weight"), d["hmid"][l],
                           self.w32(p + "self_attn.out_proj.bias"), resid=d["h"][l], site=self.site("drop1", l))
                self._ln(0, d["hmid"][l], M, T, self.w32(p + "norm2.weight"), self.w32(p + "norm2.bias"), y=d["u2"][l])
                self._gemm(1, d["u2"][l], self.w16(p + "linear1.weight"), d["ffn"][l], self.w32(p + "linear1.bias"),
                           site=self.site("ffn", l))
                self._gemm(3, d["ffn"][l], self.w16(p + "linear2.weight"), d["h"][l + 1], self.w32(p + "linear2.bias"),
                           resid=d["hmid"][l], site=self.site("drop2", l))
                nxt = lay.format(l + 1) + "norm1." if l + 1 < L else "encoder_norm."
                self._ln(0, d["h"][l + 1], M, T, self.w32(nxt + "weight"), self.w32(nxt + "bias"), y=d["u1"][l + 1])
            self._gemm(2, d["u1"][L], self.w16("feature_map.0.weight"), d["fm"], self.w32("feature_map.0.bias"))
            self._ln(2, d["fm"], M, T, self.w32("feature_map.1.weight"), self.w32("feature_map.1.bias"),
                     self.w32("cls_head.0.weight"), self.w32("cls_head.0.bias"), self.w32("reg_head.0.weight"),
                     self.w32("reg_head.0.bias"), out_f32=d["feats"], y=d["uc"], y2=d["ur"], site=self.site("feats"))
            self._gemm(1, d["uc"], self.w16("cls_head.1.weight"), d["a1c"], self.w32("cls_head.1.bias"), site=self.site("cls1"))
            self._gemm(1, d["a1c"], self.w16("cls_head.4.weight"), d["a2c"], self.w32("cls_head.4.bias"), site=self.site("cls2"))
            self._gemm(1, d["ur"], self.w16("reg_head.1.weight"), d["a1r"], self.w32("reg_head.1.bias"), site=self.site("reg1"))
            self._gemm(1, d["a1r"], self.w16("reg_head.4.weight"), d["a2r"], self.w32("reg_head.4.bias"), site=self.site("reg2"))
            check(lib.rp_head_out(ptr(d["a2c"]), ptr(d["a2r"]), ptr(self.w32("cls_head.7.weight")), ptr(self.w32("cls_head.7.bias")),
                                  ptr(self.w32("reg_head.7.weight")), ptr(self.w32("reg_head.7.bias")), ptr(d["logits"]),
                                  ptr(d["offsets"]), M, st), "rp_head_out")
        return masks, d["logits"], d["offsets"], batch.get("labels"), batch.get("segments"), d["feats"]

    # ---- backward ----------------------------------------------------------------------------------------------------
    def _reduce_tail_async(self, lo):
        """start the all-reduce of grad[lo : end of the not yet reduced part] (NCCL's own stream; it waits for the kernels
        enqueued so far on the current stream and runs beside the ones enqueued after)"""
        hi = self._ar_hi
        if hi is None or lo >= hi:
            return
        with self._timed("allreduce"):
            self._ar_pending.append(dist.all_reduce(self.opt.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                                    async_op=True))
        self._ar_hi = lo

    def backward(self, dlogits, overlap_allreduce=False):
        """gradients of every trainable parameter into the flat gradient buffer, given d loss / d logits [B,T,1];
        overlap_allreduce (step() sets it under a multi-rank NCCL group): finished tails of the buffer are all-reduced
        while the earlier layers' backward still runs — the caller must finish with _finish_allreduce()"""
        self._ar_pending, self._ar_hi = [], (self.opt.grad.numel() if overlap_allreduce else None)
        c, lib, d = self.cfg, self.lib, self._bufs
        B, T = d["shape"]
        M, L, H = B * T, c["num_layers"], c["num_heads"]
        dh, dh16, du = d["dh"], d["dh16"], d["du"]
        lay = "multimodal_encoder.layers.{}."
        with torch.cuda.device(self.dev):
            st = cur_stream()
            # cls head: Linear(256,1) <- ReLU <- Linear(256,256) <- ReLU <- Linear(512,256) <- LayerNorm
            da2 = d["dwide"].view(-1)[:M * 256].view(M, 256)
            da1 = d["dwide"].view(-1)[M * 256:2 * M * 256].view(M, 256)
            check(lib.rp_head_out_bwd_scaled(ptr(dlogits), ptr(d["a2c"]), ptr(self.w32("cls_head.7.weight")), M,
                                             self.drop_scale, ptr(da2), ptr(self.grad("cls_head.7.weight")),
                                             ptr(self.grad("cls_head.7.bias")), ptr(self._scratch), self._scratch.numel(), st),
                  "rp_head_out_bwd_scaled")
            self._wgrad(da2, d["a1c"], "cls_head.4.")
            self._colsum(da2, self.grad("cls_head.4.bias"))
            self._dgrad(da2, self.w16("cls_head.4.weight"), da1)
            self._relu_bwd_bias(da1, d["a1c"], "cls_head.1.bias")
            self._wgrad(da1, d["uc"], "cls_head.1.")
            self._dgrad(da1, self.w16("cls_head.1.weight"), du)
            feats = d["feats"].view(M, 512)
            self._ln_bwd(feats, du, "cls_head.0.", dh, None, accumulate=False)       # dh = d loss / d feats
            self._relu_bwd(dh, feats)                                                # feats = relu(LN(fm))
            self._ln_bwd(d["fm"], dh, "feature_map.1.", du, dh16, accumulate=False,  # du = d fm, dh16 = its bf16 copy
                         bias_grad="feature_map.0.bias")
            self._wgrad(dh16, d["u1"][L], "feature_map.0.")
            self._dgrad(dh16, self.w16("feature_map.0.weight"), du)                  # du = d encoder_norm output
            self._ln_bwd(d["h"][L], du, "encoder_norm.", dh, dh16, accumulate=False,
                         bias_grad=lay.format(L - 1) + "linear2.bias",               # dh = d h[L]; dh16 = dY of the last linear2
                         site=self.site("drop2", L - 1))
            dffn, dattn = d["dwide"], d["d512"]
            dqkv = d["dwide"].view(-1)[:M * 1536].view(M, 1536)
            for l in range(L - 1, -1, -1):
                p = lay.format(l)
                # FFN: h_out = hmid + relu(u2 W1^T + b1) W2^T + b2
                self._wgrad(dh16, d["ffn"][l], p + "linear2.")
                self._dgrad(dh16, self.w16(p + "linear2.weight"), dffn)
                self._relu_bwd_bias(dffn, d["ffn"][l], p + "linear1.bias")
                self._wgrad(dffn, d["u2"][l], p + "linear1.")
                self._dgrad(dffn, self.w16(p + "linear1.weight"), du)
                self._ln_bwd(d["hmid"][l], du, p + "norm2.", dh, dh16, accumulate=True,
                             bias_grad=p + "self_attn.out_proj.bias",                # dh = d hmid; dh16 = dY of out_proj
                             site=self.site("drop1", l))
                # attention: hmid = h + attn Wo^T + bo
                self._wgrad(dh16, d["attn"][l], p + "self_attn.out_proj.")
                self._dgrad(dh16, self.w16(p + "self_attn.out_proj.weight"), dattn)
                q = d["qkv"][l]
                with self._timed("bwd_fmha"):
                    if self.p > 0:
                        check(lib.rp_fmha_bwd_dropout(ptr(q), ptr(q) + 1024, ptr(q) + 2048, ptr(d["attn"][l]), ptr(dattn),
                                                      ptr(d["lse"][l]), ptr(d["dsum"]), ptr(dqkv), ptr(dqkv) + 1024,
                                                      ptr(dqkv) + 2048, 1536, 512, 1536, B, H, T, ptr(d["lens"]),
                                                      ptr(d["bits"][l]), d["bits_ld"], self.p, st), "rp_fmha_bwd_dropout")
                    else:
                        check(lib.rp_fmha_bwd(ptr(q), ptr(q) + 1024, ptr(q) + 2048, ptr(d["attn"][l]), ptr(dattn),
                                              ptr(d["lse"][l]), ptr(d["dsum"]), ptr(dqkv), ptr(dqkv) + 1024, ptr(dqkv) + 2048,
                                              1536, 512, 1536, B, H, T, ptr(d["lens"]), st), "rp_fmha_bwd")
                self._wgrad(dqkv, d["u1"][l], p + "self_attn.in_proj_weight")
                self._colsum(dqkv, self.grad(p + "self_attn.in_proj_bias"))
                self._dgrad(dqkv, self.w16(p + "self_attn.in_proj_weight"), du)
                self._ln_bwd(d["h"][l], du, p + "norm1.", dh, dh16, accumulate=True,  # dh = d h[l]; dh16 = dY of layer l-1's linear2
                             bias_grad=(lay.format(l - 1) + "linear2.bias") if l > 0 else None,
                             site=self.site("drop2", l - 1) if l > 0 else None)
                if overlap_allreduce and l > 0 and l % 4 == 0:
                    self._reduce_tail_async(self._layer_start[l])
            # h0 = LayerNorm(x W_in^T + b_in) + PE
            self._ln_bwd(d["xproj"], dh, "input_norm.", du, dh16, accumulate=False, bias_grad="input_projection.bias")
            self._wgrad(dh16, d["xcat"], "input_projection.")

    # ---- one iteration ------------------------------------------------------------------------------------------------
    def loss_and_grads(self, batch, batch_size=None, _overlap_allreduce=False):
        """forward + loss + backward (no parameter update): -> cls_loss / batch_size as a device scalar"""
        out = self.forward(batch)
        masks, logits, _, labels, _, _ = out
        bs = int(batch_size if batch_size is not None else logits.shape[0])
        loss = self.model.losses(*out)["cls_loss"] / bs
        dlogits = focal_loss_grad(masks, logits, labels, batch_size=bs)
        self.backward(dlogits, overlap_allreduce=_overlap_allreduce)
        return loss

    def _multi_rank_nccl(self):
        return (dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
                and dist.get_backend(self.group) == "nccl")

    def _finish_allreduce(self):
        """the head of the buffer, then wait for every bucket and average"""
        self._reduce_tail_async(0)
        for w in self._ar_pending:
            w.wait()
        self._ar_pending, self._ar_hi = [], None
        self.opt.grad.div_(dist.get_world_size(self.group))

    def step(self, batch, batch_size=None):
        overlap = self.overlap_allreduce and self._multi_rank_nccl()
        loss = self.loss_and_grads(batch, batch_size, _overlap_allreduce=overlap)
        with self._timed("allreduce"):
            if overlap:
                self._finish_allreduce()
            else:
                allreduce_flat_(self.opt.grad, self.group)
        with self._timed("adam_and_recast"):
            self.opt.step()
            self.refresh_weights()
        self.step_index += 1     # the next iteration draws fresh dropout masks
        return loss
