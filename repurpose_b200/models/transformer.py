"""`MultiHeadAttention` with the call signature of the reference's models/transformer.py:37-81,
executed by the sm_100a GEMM + fused-attention kernels (general-mask mode of the FMHA kernel).

    forward(q, k, v, mask=None) -> [B, Tq, d_model]
      q [B,Tq,d_model]; k, v [B,Tk,d_model] (self- or cross-attention)
      mask broadcastable to [B,1,Tq,Tk] after unsqueeze(1): [B,1,Tk] (padding) or [B,Tq,Tk]
      (band / arbitrary); positions where mask == 0 get masked_fill(-1e9) as in the reference, so a
      fully masked row attends uniformly instead of producing NaN.

Only this class of the (otherwise dead) reference module is provided; d_model must be a multiple of
256 with head dim 64.  Parameter names (`q_linear`, `k_linear`, `v_linear`, `out`, buffer `scale`)
match the reference so its state dicts load.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import _lib
from .._lib import check, cur_stream, ptr

_LOG2E = 1.4426950408889634


class MultiHeadAttention(nn.Module):
    def __init__(self, d_model, num_heads):
        super().__init__()
        if d_model % num_heads != 0 or d_model // num_heads != 64 or d_model % 256 != 0:
            raise ValueError("repurpose_b200 MultiHeadAttention needs head dim 64 and d_model % 256 == 0")
        self.num_heads = num_heads
        self.d_model = d_model
        self.d_k = d_model // num_heads
        self.q_linear = nn.Linear(d_model, d_model)
        self.k_linear = nn.Linear(d_model, d_model)
        self.v_linear = nn.Linear(d_model, d_model)
        self.out = nn.Linear(d_model, d_model)
        self.register_buffer("scale", torch.sqrt(torch.FloatTensor([self.d_k])))
        self._packed = None
        self._sig = None

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _pack(self):
        sig = self._signature()
        if self._packed is not None and self._sig == sig:
            return self._packed
        qs = _LOG2E / float(self.scale.item())
        bf = torch.bfloat16
        self._packed = dict(
            wq=(self.q_linear.weight.detach().float() * qs).to(bf).contiguous(),
            bq=(self.q_linear.bias.detach().float() * qs).contiguous(),
            wk=self.k_linear.weight.detach().to(bf).contiguous(),
            bk=self.k_linear.bias.detach().float().contiguous(),
            wv=self.v_linear.weight.detach().to(bf).contiguous(),
            bv=self.v_linear.bias.detach().float().contiguous(),
            wo=self.out.weight.detach().to(bf).contiguous(),
            bo=self.out.bias.detach().float().contiguous())
        self._sig = sig
        return self._packed

    @torch.no_grad()
    def forward(self, q, k, v, mask=None):
        dev = q.device
        if dev.type != "cuda":
            raise _lib.RepurposeError("MultiHeadAttention runs on CUDA tensors only (no CPU path)")
        lib = _lib.load()
        w = self._pack()
        B, Tq, D = q.shape
        Tk = k.shape[1]
        st_args = dict(dtype=torch.bfloat16, device=dev)

        def project(x, wgt, bias, T):
            x32 = x.to(torch.float32).contiguous()
            xb = torch.empty(B * T, D, **st_args)
            check(lib.rp_cast_bf16(ptr(x32), ptr(xb), x32.numel(), cur_stream()), "rp_cast_bf16")
            y = torch.empty(B * T, D, **st_args)
            check(lib.rp_gemm_bf16(0, ptr(xb), D, ptr(wgt), D, ptr(y), D, ptr(bias), 0, 0, B * T, D, D,
                                   cur_stream()), "rp_gemm_bf16")
            return y

        with torch.cuda.device(dev):
            qp = project(q, w["wq"], w["bq"], Tq)
            kp = project(k, w["wk"], w["bk"], Tk)
            vp = project(v, w["wv"], w["bv"], Tk)
            o = torch.empty(B * Tq, D, **st_args)
            mask_u8 = None
            mode, mb, mq = 0, 0, 0
            if mask is not None:
                m = mask.to(dev)
                if m.dim() != 3 or m.shape[0] != B or m.shape[-1] != Tk or m.shape[1] not in (1, Tq):
                    raise ValueError(f"mask shape {tuple(mask.shape)} not [B,1,Tk] or [B,Tq,Tk]")
                mask_u8 = m.ne(0).to(torch.uint8).contiguous()
                mode = 1
                mb = mask_u8.stride(0)
                mq = 0 if m.shape[1] == 1 else mask_u8.stride(1)
            check(lib.rp_fmha(ptr(qp), ptr(kp), ptr(vp), ptr(o), D, D, D, D, Tq * D, Tk * D, Tk * D,
                              Tq * D, B, self.num_heads, Tq, Tk, 0, mode, ptr(mask_u8), mb, mq,
                              cur_stream()), "rp_fmha")
            out = torch.empty(B, Tq, D, dtype=torch.float32, device=dev)
            check(lib.rp_gemm_bf16(2, ptr(o), D, ptr(w["wo"]), D, ptr(out), D, ptr(w["bo"]), 0, 0,
                                   B * Tq, D, D, cur_stream()), "rp_gemm_bf16")
        return out
