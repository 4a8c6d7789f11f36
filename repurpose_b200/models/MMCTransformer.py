"""Drop-in `MMCTransformer` for the Repurpose inference hot path, executed by hand-written sm_100a
kernels through the C ABI (include/repurpose_b200.h).

Mirrors the reference interface (paths relative to YosubShin/Repurpose):
  * constructor              models/MMCTransformer.py:26   (same argument names; `text_num_layers`
                                                            and `cross_num_layers` are ignored there too)
  * forward(batch)           models/MMCTransformer.py:109-151  -> the same 6-tuple
  * inference_(batch, cfg)   models/MMCTransformer.py:231-275  -> list of dicts
  * inference_single_video   models/MMCTransformer.py:181-229
  * state_dict keys / shapes models/MMCTransformer.py:32-93    (checkpoints load unchanged)

The torch sub-modules below exist only as the parameter store (same construction order as the
reference, so the same `torch.manual_seed` yields the same random init); no torch op of theirs is
ever executed.  All compute happens in librepurpose_b200.so — there is no PyTorch/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from .._lib import RpDecodeCfg, RpModelCfg, check, cur_stream, ptr
from .softnms import soft_nms_intervals_cpu  # noqa: F401  (re-exported like the reference module)

HEAD_HIDDEN = 256
_STRICT_MASKS = os.environ.get("RP_STRICT_MASKS", "0") == "1"


class PositionalEncoding(nn.Module):
    """Sinusoidal table, same formula/layout as models/MMCTransformer.py:9-22 (buffer `pe`
    [1, max_len, d_model]); only the table is used — the add is fused into a LayerNorm kernel."""

    def __init__(self, d_model: int, max_len: int = 5000):
        super().__init__()
        self.register_buffer("pe", self.table(d_model, max_len).unsqueeze(0))

    @staticmethod
    def table(d_model: int, max_len: int) -> torch.Tensor:
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        return pe


class _GraphedInference:
    """forward -> decode -> Soft-NMS -> slot packing of one (batch, padded length, decode settings) shape as ONE
    CUDA graph over static buffers: the reference's call pattern is batch size 1 with a host synchronisation per
    video (inference.py:31, 39-47), where the ~125 kernel launches and the Python around them — not the GPU —
    set the latency.  T is padded to a multiple of 128 (15 shapes cover every length up to 1801); rows beyond a
    video's length never reach a valid row (keys are masked by length, everything else is row-wise)."""

    def __init__(self, model, B, Tp, settings):
        from ..scheduler import pack_slots
        dev = model.device
        c = model._cfg
        self.B, self.Tp = B, Tp
        f32 = dict(dtype=torch.float32, device=dev)
        self.vis = torch.zeros(B, Tp, c["vis_dim"], **f32)
        self.aud = torch.zeros(B, Tp, c["aud_dim"], **f32)
        self.txt = torch.zeros(B, Tp, c["text_dim"], **f32)
        self.mask8 = torch.zeros(B, Tp, dtype=torch.uint8, device=dev)
        self.max_seg = torch.zeros(B, dtype=torch.int32, device=dev)
        self.lens = torch.zeros(B, dtype=torch.int32, device=dev)
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.logits = torch.empty(B, Tp, 1, **f32)
        self.offsets = torch.empty(B, Tp, 2, **f32)
        self.feats = torch.empty(B, Tp, c["d_model"], **f32)
        self.kcap = kcap = max(1, int(np.ceil((Tp // 60) * settings["max_seg_per_min"])))
        self.segs = torch.zeros(B, kcap, 2, **f32)
        self.scores = torch.zeros(B, kcap, **f32)
        self.dscores = torch.zeros(B, kcap, **f32)
        self.labels = torch.zeros(B, kcap, dtype=torch.int32, device=dev)
        self.counts = torch.zeros(B, dtype=torch.int32, device=dev)
        self.ncand = torch.zeros(B, dtype=torch.int32, device=dev)
        self.slots_host = torch.zeros(B, 1 + 4 * kcap, dtype=torch.float32).pin_memory()
        self.flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.max_seg_host = torch.zeros(B, dtype=torch.int32).pin_memory()
        cfg = model._decode_cfg(settings)
        lib = _lib.load()
        ws = model._get_workspace(B, Tp)

        def launch():
            st = cur_stream()
            check(lib.rp_mask_lens(ptr(self.mask8), B, Tp, ptr(self.lens), ptr(self.flag), st), "rp_mask_lens")
            check(lib.rp_forward(model._handle, ptr(self.vis), ptr(self.aud), ptr(self.txt), ptr(self.lens), B, Tp,
                                 ptr(self.logits), ptr(self.offsets), ptr(self.feats), ptr(ws), ws.numel(), st),
                  "rp_forward")
            check(lib.rp_decode_nms(ptr(self.logits), ptr(self.offsets), ptr(self.lens), ptr(self.max_seg), B, Tp,
                                    C.byref(cfg), kcap, ptr(self.segs), ptr(self.scores), ptr(self.dscores),
                                    ptr(self.labels), ptr(self.counts), ptr(self.ncand), 0, 0, 0, st), "rp_decode_nms")
            self.slots_host.copy_(pack_slots(self.segs, self.scores, self.labels, self.counts), non_blocking=True)
            self.flag_host.copy_(self.flag, non_blocking=True)

        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                launch()                       # warm-up outside the capture (lazy kernel configuration)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            # (an explicit capture stream: torch's default one is a process-wide singleton created on the first
            # capturing device, which a second device of the same process cannot capture on)
            with torch.cuda.graph(self.graph, stream=side):
                launch()
        self.workspace = ws                    # the graph holds its address

    def run(self, vis, aud, txt, masks, max_seg):
        """copies the inputs into the static buffers (H2D or D2D), replays and waits; -> host slots [B, 1+4K]"""
        T = vis.shape[1]
        self.vis[:, :T].copy_(vis, non_blocking=True)
        self.aud[:, :T].copy_(aud, non_blocking=True)
        self.txt[:, :T].copy_(txt, non_blocking=True)
        self.mask8.zero_()
        self.mask8[:, :T].copy_(masks.reshape(self.B, T), non_blocking=True)
        self.max_seg_host.copy_(torch.as_tensor(max_seg, dtype=torch.int32))
        self.max_seg.copy_(self.max_seg_host, non_blocking=True)
        self.graph.replay()
        torch.cuda.current_stream(self.vis.device).synchronize()
        if int(self.flag_host[0]) != 0:
            raise _lib.RepurposeError(MMCTransformer._MASK_MSG)
        return self.slots_host


def _invalidate_hook(module, _incompatible_keys):
    module._invalidate()


_NATIVE_STATE = dict(_handle=None, _weights_sig=None, _workspace=None, _last_masks=None, _last_lens=None, _graphs=None)


class MMCTransformer(nn.Module):
    def __init__(self, vis_dim, aud_dim, text_dim, d_model, self_num_layers, text_num_layers,
                 cross_num_layers, num_heads, d_ff=2048, max_len=5000):
        super().__init__()
        self._cfg = dict(vis_dim=vis_dim, aud_dim=aud_dim, text_dim=text_dim, d_model=d_model,
                         num_layers=self_num_layers, num_heads=num_heads, d_ff=d_ff,
                         head_hidden=HEAD_HIDDEN, max_len=max_len)
        # ---- parameter store (names, shapes and RNG order of the reference) -------------------
        self.input_projection = nn.Linear(vis_dim + aud_dim + text_dim, d_model)
        self.input_norm = nn.LayerNorm(d_model)
        self.positional_encoding = PositionalEncoding(d_model, max_len)
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=num_heads, dim_feedforward=d_ff,
                                           dropout=0.1, activation="relu", batch_first=True,
                                           norm_first=True)
        self.multimodal_encoder = nn.TransformerEncoder(layer, num_layers=self_num_layers,
                                                        enable_nested_tensor=False)
        self.encoder_norm = nn.LayerNorm(d_model)
        self.feature_map = nn.Sequential(nn.Linear(d_model, d_model), nn.LayerNorm(d_model),
                                         nn.ReLU(), nn.Dropout(0.1))
        self.cls_head = nn.Sequential(
            nn.LayerNorm(d_model), nn.Linear(d_model, HEAD_HIDDEN), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(HEAD_HIDDEN, HEAD_HIDDEN), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(HEAD_HIDDEN, 1))
        self.reg_head = nn.Sequential(
            nn.LayerNorm(d_model), nn.Linear(d_model, HEAD_HIDDEN), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(HEAD_HIDDEN, HEAD_HIDDEN), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(HEAD_HIDDEN, 2), nn.ReLU())
        self._init_weights()
        # ---- native state ---------------------------------------------------------------------
        self._handle = None
        self._weights_sig = None
        self._workspace = None
        self._mask_checks = []
        self._last_masks = self._last_lens = None
        self._graphs = None
        # "auto": CUDA-graph replay for launch-bound calls (at most GRAPH_MAX_BATCH videos); True / False force it
        self.cuda_graphs = "auto" if os.environ.get("RP_CUDA_GRAPHS", "1") != "0" else False
        self.register_load_state_dict_post_hook(_invalidate_hook)

    def _init_weights(self):
        # models/MMCTransformer.py:98-107: Xavier for every nn.Linear, zero bias, LN gamma=1 beta=0
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    # ------------------------------------------------------------------------------- plumbing
    @property
    def device(self):
        return next(self.parameters()).device

    def _invalidate(self):
        self._weights_sig = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._weights_sig = None
        return out

    def __getstate__(self):
        # copy.deepcopy / pickle / torch.save(model): the native handle (a ctypes pointer), the workspace and
        # pending events stay behind; the copy repacks its weights lazily on its first forward
        st = self.__dict__.copy()
        st.update(_NATIVE_STATE)
        st["_mask_checks"] = []
        return st

    def __del__(self):
        try:
            h = self.__dict__.get("_handle")
            if h is not None:
                self.__dict__["_handle"] = None
                _lib.load().rp_destroy(h)
        except Exception:
            pass

    def _signature(self):
        sig = [self.device]
        for p in self.parameters():
            sig.append((p.data_ptr(), p._version))
        return tuple(sig)

    def refresh_weights(self):
        """(Re)pack the current parameters into the native handle.  Called automatically when a
        parameter's storage or autograd version changed (`load_state_dict`, `.to()`, optimizer
        steps); call it by hand after editing weights through `.data`."""
        dev = self.device
        if dev.type != "cuda":
            raise _lib.RepurposeError(
                "repurpose_b200.MMCTransformer runs only on a CUDA (sm_100) device; "
                "move the module with .to('cuda') — there is no CPU path")
        lib = _lib.load()
        with torch.cuda.device(dev):
            if self._handle is None:
                cfg = RpModelCfg(**self._cfg)
                h = C.c_void_p()
                check(lib.rp_create(C.byref(cfg), C.byref(h)), "rp_create")
                self._handle = h
            st = cur_stream()
            for name, t in self.state_dict().items():
                t32 = t.detach().to(device=dev, dtype=torch.float32).contiguous()
                check(lib.rp_load_weight(self._handle, name.encode(), ptr(t32), t32.numel(), st),
                      f"rp_load_weight({name})")
            check(lib.rp_weights_complete(self._handle), "rp_weights_complete")
            torch.cuda.current_stream().synchronize()  # staging copies above may be temporaries
        self._weights_sig = self._signature()

    def _ensure_ready(self):
        if self._handle is None or self._weights_sig != self._signature():
            self.refresh_weights()

    def set_skip_padding(self, on: bool):
        """Extension: whether forward skips the blocks / attention tiles that hold nothing but padding and returns zeros
        at padded steps (default on; include/repurpose_b200.h rp_set_skip_padding).  Valid steps are bit-identical."""
        self._ensure_ready()
        check(_lib.load().rp_set_skip_padding(self._handle, 1 if on else 0), "rp_set_skip_padding")
        self._graphs = None   # captured graphs hold the old setting

    def _get_workspace(self, B, T):
        need = _lib.load().rp_workspace_bytes(self._handle, B, T)
        if need < 0:
            raise _lib.RepurposeError("rp_workspace_bytes failed")
        ws = self._workspace
        if ws is None or ws.numel() < need or ws.device != self.device:
            self._workspace = ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return ws

    GRAPH_MAX_BATCH = 4

    def _graphed(self, B, T, settings):
        """the CUDA-graph entry for this shape, or None when the call is not launch-bound / graphs are off"""
        if self.cuda_graphs is False or (self.cuda_graphs == "auto" and B > self.GRAPH_MAX_BATCH):
            return None
        Tp = -(-T // 128) * 128
        if Tp > self._cfg["max_len"] or int(np.ceil((Tp // 60) * settings["max_seg_per_min"])) > 64:
            return None
        key = (B, Tp, tuple(sorted((k, float(v)) for k, v in settings.items())))
        if self._graphs is None:
            self._graphs = {}
        e = self._graphs.get(key)
        if e is None:
            if len(self._graphs) >= 64:         # a pathological mix of shapes: start over rather than grow
                self._graphs.clear()
            e = self._graphs[key] = _GraphedInference(self, B, Tp, settings)
        return e

    @staticmethod
    def _as_f32(t, dev):
        return t.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()

    # ------------------------------------------------------------------------ masks -> lengths
    _MASK_MSG = ("masks must be left-aligned (masks[b, 0, t] == t < valid_len[b], as dataset/RepurposeClip.py:"
                 "528-531 builds them): the attention and decode kernels take lengths; a mask with holes or "
                 "leading padding would silently attend to / decode the wrong steps")

    def _lens_from_masks(self, masks, B, T):
        """Valid steps per video (int32 [B] on the device) from a [B,1,T] / [B,T] mask, and the check that
        the reference's key-padding semantics (models/MMCTransformer.py:132-138 hands the MASK to the
        encoder) are expressible as lengths.  Host masks are checked on the spot; device masks by one
        kernel whose flag is read back asynchronously and raises at the next host synchronisation point of
        this module (or at once with RP_STRICT_MASKS=1) — a forward never waits for the GPU."""
        dev = self.device
        m = masks.reshape(B, -1)
        if m.shape[1] != T:
            raise ValueError(f"masks cover {m.shape[1]} steps, features {T}")
        if not m.is_cuda:
            mb = m.ne(0)
            lens = mb.sum(dim=1, dtype=torch.int32)
            if not torch.equal(mb, torch.arange(T)[None, :] < lens[:, None]):
                raise _lib.RepurposeError(self._MASK_MSG)
            return lens.to(dev, non_blocking=True)
        m = m.to(dev)
        m8 = (m if m.dtype in (torch.bool, torch.uint8) else m.ne(0)).contiguous().view(torch.uint8)
        lens = torch.empty(B, dtype=torch.int32, device=dev)
        flag = torch.empty(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load().rp_mask_lens(ptr(m8), B, T, ptr(lens), ptr(flag), cur_stream()), "rp_mask_lens")
            host = torch.empty(1, dtype=torch.int32, pin_memory=True)
            host.copy_(flag, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self._mask_checks.append((ev, host))
        self._poll_mask_checks(wait=_STRICT_MASKS)
        return lens

    def _poll_mask_checks(self, wait=False):
        """Raise if a finished alignment check found a bad mask (wait=True: finish them all first)."""
        pending = []
        for ev, host in self._mask_checks:
            if wait:
                ev.synchronize()
            if ev.query():
                if int(host[0]) != 0:
                    self._mask_checks = []
                    raise _lib.RepurposeError(self._MASK_MSG)
            else:
                pending.append((ev, host))
        self._mask_checks = pending

    # -------------------------------------------------------------------------------- forward
    def forward(self, batch):
        """batch: dict with visual_feats [B,T,Cv], audio_feats [B,T,Ca], text_feats [B,T,Ct] fp32,
        masks [B,1,T] bool (True = valid, left-aligned as built by dataset/RepurposeClip.py:528-531),
        labels, segments (passed through).  Returns (masks, cls_logits [B,T,1], offsets [B,T,2],
        labels, segments, feats [B,T,d_model]) like the reference.
        A batch from `repurpose_b200.features.collate_ragged` (unpadded rows + offsets) is accepted
        too: the padding then happens on the device."""
        self._ensure_ready()
        self._poll_mask_checks()
        dev = self.device
        if batch.get("parts") is not None:  # zero-copy ragged batch used outside the pipeline
            batch = dict(batch)
            for k, ts in batch.pop("parts").items():
                batch[k] = torch.cat([t.to(dev, non_blocking=True) for t in ts]) if sum(int(t.shape[0]) for t in ts) \
                    else torch.zeros(1, int(ts[0].shape[1]), device=dev, dtype=ts[0].dtype)
        # ragged batches may carry pre-converted bf16 feature rows (half the PCIe bytes, identical results)
        feats_bf16 = bool(batch.get("ragged")) and all(
            torch.is_tensor(batch[k]) and batch[k].dtype == torch.bfloat16
            for k in ("visual_feats", "audio_feats", "text_feats"))
        conv = (lambda t, d: t.to(d, non_blocking=True).contiguous()) if feats_bf16 else self._as_f32
        vis = conv(batch["visual_feats"], dev)
        aud = conv(batch["audio_feats"], dev)
        txt = conv(batch["text_feats"], dev)
        lib = _lib.load()
        if batch.get("ragged"):
            i32 = lambda t: t.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
            lens = i32(batch["lens"])
            B, T = len(batch["duration"]), int(batch.get("max_len") or max(batch["duration"]))
            masks = (torch.arange(T, device=dev)[None, :] < lens[:, None]).unsqueeze(1)
            roff, toff, tlen = i32(batch["row_offsets"]), i32(batch["text_offsets"]), i32(batch["text_lens"])
        else:
            masks = batch["masks"]
            B, T = vis.shape[0], vis.shape[1]
            lens = self._lens_from_masks(masks, B, T)
        logits = torch.empty(B, T, 1, dtype=torch.float32, device=dev)
        offsets = torch.empty(B, T, 2, dtype=torch.float32, device=dev)
        feats = torch.empty(B, T, self._cfg["d_model"], dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = self._get_workspace(B, T)
            if batch.get("ragged"):
                fwd = lib.rp_forward_ragged_bf16 if feats_bf16 else lib.rp_forward_ragged
                check(fwd(self._handle, ptr(vis), ptr(aud), ptr(txt), ptr(roff), ptr(toff),
                          ptr(tlen), ptr(lens), B, T, ptr(logits), ptr(offsets), ptr(feats),
                          ptr(ws), ws.numel(), cur_stream()), "rp_forward_ragged")
            else:
                check(lib.rp_forward(self._handle, ptr(vis), ptr(aud), ptr(txt), ptr(lens), B, T,
                                     ptr(logits), ptr(offsets), ptr(feats), ptr(ws), ws.numel(),
                                     cur_stream()), "rp_forward")
        self._last_masks, self._last_lens = masks, lens
        return masks, logits, offsets, batch.get("labels"), batch.get("segments"), feats

    def profile_begin(self):
        """Bracket every kernel of the following forward passes with CUDA events (bench.py)."""
        self._ensure_ready()
        check(_lib.load().rp_profile_begin(self._handle), "rp_profile_begin")

    def profile_end(self):
        """-> {kernel class: (total device ms, launches)} since profile_begin(); synchronises."""
        n = len(_lib.PROFILE_TAGS)
        ms = (C.c_float * n)()
        cnt = (C.c_int32 * n)()
        check(_lib.load().rp_profile_end(self._handle, ms, cnt), "rp_profile_end")
        return {t: (float(ms[i]), int(cnt[i])) for i, t in enumerate(_lib.PROFILE_TAGS)}

    @torch.no_grad()
    def losses(self, masks, out_cls_logits, out_offsets, gt_cls_labels, gt_offsets, feats):
        """Forward VALUE of the reference's loss (models/MMCTransformer.py:159-179): masked sigmoid focal
        loss (alpha 0.7, gamma 2, models/losses.py:5-53) summed over valid steps, as main.py's evaluation
        loop logs it (main.py:626-634).  Returns {'cls_loss': 0-dim fp32 tensor on the device}.  No
        autograd graph: the training step is outside this path."""
        dev = self.device
        logits = out_cls_logits.to(dev, torch.float32).reshape(-1).contiguous()
        targets = gt_cls_labels.to(dev, torch.float32).reshape(-1).contiguous()
        mask = masks.to(dev).reshape(-1).ne(0).to(torch.uint8).contiguous()
        if not (logits.numel() == targets.numel() == mask.numel()):
            raise ValueError("losses: logits [B,T,1], labels [B,T] and masks [B,1,T] must cover the same steps")
        out = torch.empty(1, dtype=torch.float32, device=dev)
        scratch = torch.empty(296, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load().rp_focal_loss_sum(ptr(logits), ptr(targets), ptr(mask), logits.numel(), 0.7, 2.0,
                                                ptr(scratch), ptr(out), cur_stream()), "rp_focal_loss_sum")
        return {"cls_loss": out[0]}

    # --------------------------------------------------------------------------------- decode
    @staticmethod
    def _decode_cfg(s):
        return RpDecodeCfg(int(s["pre_nms_topk"]), float(s["pre_nms_thresh"]),
                           float(s["duration_thresh"]), float(s["duration_thresh_max"]),
                           float(s.get("nms_sigma", 0.5)), float(s.get("min_score", 0.001)))

    def _run_decode(self, logits, offsets, lens, max_seg, settings, want_candidates=False):
        """logits [B,T] f32, offsets [B,T,2] f32, lens [B] i32 (device), max_seg: list[int]."""
        dev = logits.device
        B, T = logits.shape
        kcap = max(1, max(max_seg))
        if kcap > 64:
            return self._run_decode_wide(logits, offsets, lens, max_seg, settings, want_candidates)
        max_seg_t = torch.tensor(max_seg, dtype=torch.int32).to(dev, non_blocking=True)
        segs = torch.empty(B, kcap, 2, dtype=torch.float32, device=dev)
        scores = torch.empty(B, kcap, dtype=torch.float32, device=dev)
        dscores = torch.empty(B, kcap, dtype=torch.float32, device=dev)
        labels = torch.empty(B, kcap, dtype=torch.int32, device=dev)
        counts = torch.empty(B, dtype=torch.int32, device=dev)
        ncand = torch.empty(B, dtype=torch.int32, device=dev)
        cseg = cscore = clabel = None
        if want_candidates:
            ccap = min(int(settings["pre_nms_topk"]), T)
            cseg = torch.empty(B, ccap, 2, dtype=torch.float32, device=dev)
            cscore = torch.empty(B, ccap, dtype=torch.float32, device=dev)
            clabel = torch.empty(B, ccap, dtype=torch.int32, device=dev)
        cfg = self._decode_cfg(settings)
        with torch.cuda.device(dev):
            check(_lib.load().rp_decode_nms(ptr(logits), ptr(offsets), ptr(lens), ptr(max_seg_t), B, T,
                                            C.byref(cfg), kcap, ptr(segs), ptr(scores), ptr(dscores),
                                            ptr(labels), ptr(counts), ptr(ncand), ptr(cseg),
                                            ptr(cscore), ptr(clabel), cur_stream()), "rp_decode_nms")
        return dict(segments=segs, scores=scores, dscores=dscores, labels=labels, counts=counts,
                    ncand=ncand, cand_segments=cseg, cand_scores=cscore, cand_labels=clabel)

    def _run_decode_wide(self, logits, offsets, lens, max_seg, settings, want_candidates):
        """max_seg_num above the 64 slots the fused decode kernel keeps in shared memory (max_seg_per_min >= 1
        on an hour-long video; the reference has no such limit): the fused kernel only builds the candidate
        lists, the stand-alone Soft-NMS kernel (Kcap up to the candidate count) selects, and the kept rows
        are gathered on the device.  Same outputs as `_run_decode`."""
        from .softnms import soft_nms_batched
        dev = logits.device
        r = self._run_decode(logits, offsets, lens, [1] * len(max_seg), settings, want_candidates=True)
        ccap = r["cand_scores"].shape[1]
        kcap = max(1, min(max(max_seg), ccap))
        max_seg_t = torch.tensor(max_seg, dtype=torch.int32).to(dev, non_blocking=True)
        keep, dscores, counts = soft_nms_batched(r["cand_scores"], r["cand_segments"], r["ncand"], max_seg_t,
                                                 float(settings.get("nms_sigma", 0.5)),
                                                 float(settings.get("min_score", 0.001)), kcap=kcap)
        idx = keep.clamp_min(0).to(torch.int64)
        out = dict(segments=torch.gather(r["cand_segments"], 1, idx[..., None].expand(-1, -1, 2)),
                   scores=torch.gather(r["cand_scores"], 1, idx), dscores=dscores,
                   labels=torch.gather(r["cand_labels"], 1, idx), counts=counts, ncand=r["ncand"],
                   cand_segments=None, cand_scores=None, cand_labels=None)
        if want_candidates:
            out.update(cand_segments=r["cand_segments"], cand_scores=r["cand_scores"], cand_labels=r["cand_labels"])
        return out

    @torch.no_grad()
    def inference_single_video(self, masks, out_cls_logits, out_offsets, inference_settings):
        """Pre-NMS candidates of one video (reference :181-229): dict(segments [N,2], scores [N]
        descending, labels [N] int64)."""
        dev = out_cls_logits.device
        logits = out_cls_logits.reshape(1, -1).to(torch.float32).contiguous()
        T = logits.shape[1]
        offsets = out_offsets.reshape(1, T, 2).to(torch.float32).contiguous()
        lens = self._lens_from_masks(masks, 1, T)
        r = self._run_decode(logits, offsets, lens, [1], inference_settings, want_candidates=True)
        n = int(r["ncand"][0].item())
        self._poll_mask_checks()
        return {"segments": r["cand_segments"][0, :n], "scores": r["cand_scores"][0, :n],
                "labels": r["cand_labels"][0, :n].to(torch.int64)}

    @torch.no_grad()
    def inference_device(self, batch, inference_settings):
        """forward + decode + Soft-NMS with no host synchronisation: returns the fixed-slot device
        tensors of `_run_decode` (segments [B,K,2], scores, dscores, labels [B,K], counts [B])."""
        return self.decode_device(self.forward(batch), batch, inference_settings)

    def decode_device(self, output, batch, inference_settings):
        """decode + Soft-NMS of an existing `forward(batch)` result (the 6-tuple): main.py's evaluation loop
        calls `model(batch)` for the loss and then `inference_(batch, ...)`, which in the reference runs the
        whole forward a second time (main.py:626, :673-676); `inference_(batch, cfg, output=output)` reuses it."""
        masks, logits, offsets = output[0], output[1], output[2]
        dev = self.device
        max_seg = [int(np.ceil((int(v) // 60) * inference_settings["max_seg_per_min"]))
                   for v in batch["duration"]]
        B, T = logits.shape[0], logits.shape[1]
        # the forward that produced `output` already derived (and checked) the lengths of this mask
        lens = self._last_lens if masks is self._last_masks and self._last_lens is not None \
            and self._last_lens.shape[0] == B else self._lens_from_masks(masks, B, T)
        return self._run_decode(logits.reshape(B, T).contiguous(), offsets.contiguous(), lens, max_seg,
                                inference_settings)

    @torch.no_grad()
    def inference_(self, batch, inference_settings, to_host=False, output=None):
        """forward -> per-video decode -> Soft-NMS, all on the device (reference :231-275).
        Returns one dict per video: segments [K,2] f32 (feature-grid seconds), scores [K] f32
        (the candidate's probability, i.e. the reference's CUDA semantics),
        labels [K] int64, video_id, duration — in Soft-NMS selection order.
        Tensors live on the model's device like the reference's; `to_host=True` (extension) returns
        CPU tensors taken from one packed device->host copy, which is what callers that immediately
        do `.tolist()` (inference.py:47, main.py:689) want.  `output` (extension): a `forward(batch)` result to
        decode instead of running the forward again."""
        vid_idxs = batch["video_id"]
        vid_lens = batch["duration"]
        results = []
        if output is None and not batch.get("ragged") and batch.get("parts") is None:
            self._ensure_ready()
            vis = batch["visual_feats"]
            g = self._graphed(int(vis.shape[0]), int(vis.shape[1]), inference_settings)
            if g is not None:
                from ..scheduler import unpack_slots
                max_seg = [int(np.ceil((int(v) // 60) * inference_settings["max_seg_per_min"])) for v in vid_lens]
                with torch.cuda.device(self.device):
                    host = g.run(vis, batch["audio_feats"], batch["text_feats"], batch["masks"], max_seg)
                for o, vidx, vlen in zip(unpack_slots(host), vid_idxs, vid_lens):
                    if not to_host:             # like the reference: tensors on the model's device
                        o = {k: v.to(self.device) for k, v in o.items()}
                    o["video_id"], o["duration"] = vidx, vlen
                    results.append(o)
                return results
        r = (self.inference_device(batch, inference_settings) if output is None
             else self.decode_device(output, batch, inference_settings))
        if to_host:
            from ..scheduler import pack_slots, unpack_slots
            slots = pack_slots(r["segments"], r["scores"], r["labels"], r["counts"])
            for o, vidx, vlen in zip(unpack_slots(slots), vid_idxs, vid_lens):
                o["video_id"], o["duration"] = vidx, vlen
                results.append(o)
            self._poll_mask_checks()
            return results
        counts = r["counts"].tolist()  # the single device->host sync of the whole batch
        self._poll_mask_checks()
        labels64 = r["labels"].to(torch.int64)
        for i, (vidx, vlen) in enumerate(zip(vid_idxs, vid_lens)):
            k = counts[i]
            results.append({"segments": r["segments"][i, :k], "scores": r["scores"][i, :k],
                            "labels": labels64[i, :k], "video_id": vidx, "duration": vlen})
        return results
