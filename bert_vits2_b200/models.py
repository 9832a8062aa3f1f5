"""Drop-in for the reference `models.SynthesizerTrn` on the inference path.

Mirrors (names, argument meaning, return values, error behaviour):
  * constructor: reference models.py:816-935 as called by infer.get_net_g (infer.py:95-101):
        SynthesizerTrn(len(symbols), filter_length // 2 + 1, segment_size // hop_length,
                       n_speakers=hps.data.n_speakers, **hps.model)
  * .to(device) / .eval() / .state_dict() / .load_state_dict(): the parameters are registered under the
    reference's state_dict keys (spec.py), so utils.load_checkpoint (reference utils.py:65-120) works
    unmodified; `enc_q.*` keys in a checkpoint are ignored (strict=False there), as compress_model.py drops them.
  * .infer(...): reference models.py:1026-1074, same signature, returns
        (o [B,1,L], attn [B,1,F,T], y_mask [B,1,F], (z, z_p, m_p, logs_p))

Only tensor plumbing happens here.  All arithmetic runs in libbv2.so (CUDA, sm_100a); there is no PyTorch or
CPU fallback — calling .infer() on a CPU module raises.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import synth
from .engine import Bv2Error, Engine
from .spec import ModelConfig, param_specs


class LazyAttn:
    """Stand-in for the `attn` tensor infer() returns (reference models.py:1074): the dense [B,1,F,T] one-hot path is only
    written when somebody reads it (the reference's own callers never do: infer.py:302-318 uses `o` alone).  Any attribute
    access, indexing or torch function materialises it once through bv2_attn_path; `.materialize()` returns the tensor."""

    def __init__(self, engine, shape, token):
        self._engine, self._shape, self._token, self._t = engine, tuple(shape), token, None

    def materialize(self) -> torch.Tensor:
        if self._t is None:
            if getattr(self._engine, "_attn_token", None) is not self._token:
                raise RuntimeError("attn of an earlier infer() call: materialise it before the next call on the same module")
            self._t = self._engine.attn_path()
        return self._t

    @property
    def shape(self):
        return torch.Size(self._shape)

    def size(self, *a):
        return self.shape if not a else self.shape[a[0]]

    def dim(self):
        return len(self._shape)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __len__(self):
        return self._shape[0]

    def __repr__(self):
        return f"LazyAttn(shape={self._shape}, materialized={self._t is not None})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        conv = lambda a: a.materialize() if isinstance(a, LazyAttn) else a  # noqa: E731
        return func(*[conv(a) for a in args], **{k: conv(v) for k, v in (kwargs or {}).items()})


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's dotted state_dict key tree."""


class SynthesizerTrn(nn.Module):
    def __init__(self, n_vocab, spec_channels, segment_size, inter_channels, hidden_channels, filter_channels, n_heads,
                 n_layers, kernel_size, p_dropout, resblock, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates,
                 upsample_initial_channel, upsample_kernel_sizes, n_speakers=256, gin_channels=256, use_sdp=True,
                 n_flow_layer=4, n_layers_trans_flow=4, flow_share_parameter=False, use_transformer_flow=True,
                 precision: str = "fp16", init_seed: Optional[int] = 0, **kwargs):
        super().__init__()
        if n_speakers < 1:
            raise ValueError("n_speakers == 0 (ReferenceEncoder path, models.py:752-808) is a training-only configuration")
        if flow_share_parameter:
            raise ValueError("flow_share_parameter=True references attentions.FFT, which does not exist in the reference")
        if not kwargs.get("use_spk_conditioned_encoder", True):
            raise ValueError("use_spk_conditioned_encoder=False is not supported")
        self.n_vocab, self.spec_channels, self.segment_size = n_vocab, spec_channels, segment_size
        self.inter_channels, self.hidden_channels, self.filter_channels = inter_channels, hidden_channels, filter_channels
        self.n_heads, self.n_layers, self.kernel_size, self.p_dropout = n_heads, n_layers, kernel_size, p_dropout
        self.n_speakers, self.gin_channels, self.use_sdp = n_speakers, gin_channels, use_sdp
        self.precision = precision
        model = dict(inter_channels=inter_channels, hidden_channels=hidden_channels, filter_channels=filter_channels,
                     n_heads=n_heads, n_layers=n_layers, kernel_size=kernel_size, resblock=resblock,
                     resblock_kernel_sizes=list(resblock_kernel_sizes),
                     resblock_dilation_sizes=[list(d) for d in resblock_dilation_sizes], upsample_rates=list(upsample_rates),
                     upsample_initial_channel=upsample_initial_channel, upsample_kernel_sizes=list(upsample_kernel_sizes),
                     gin_channels=gin_channels, use_sdp=use_sdp, n_flow_layer=n_flow_layer,
                     n_layers_trans_flow=n_layers_trans_flow, use_transformer_flow=use_transformer_flow)
        self.cfg = ModelConfig.from_hps_model(model, n_vocab=n_vocab, n_speakers=n_speakers)
        init = synth.synthetic_state_dict(self.cfg, init_seed) if init_seed is not None else None
        for p in param_specs(self.cfg):
            parts = p.key.split(".")
            node = self
            for name in parts[:-1]:
                if name not in node._modules:
                    node.add_module(name, _Node())
                node = node._modules[name]
            t = init[p.key] if init is not None else torch.zeros(p.shape)
            node.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))
        self._engines = {}
        self._weights_version = 0
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    # -- weight changes invalidate the packed device copy ------------------------------------------------
    def _invalidate(self):
        self._weights_version += 1
        self._engines.clear()

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        self._invalidate()
        return r

    def remove_weight_norm(self):
        """No-op: weight-norm is folded once when the engine packs its weights (the reference re-evaluates it on
        every forward because nobody calls this)."""

    def _engine(self, device: torch.device) -> Engine:
        key = (str(device), self._weights_version)
        eng = self._engines.get(key)
        if eng is None:
            sd = {k: v for k, v in self.state_dict().items()}
            eng = Engine(self.cfg, sd, device=device, precision=self.precision)
            self._engines = {key: eng}
        return eng

    # ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def infer(self, x, x_lengths, sid, tone, language, bert, ja_bert, en_bert, noise_scale=0.667, length_scale=1,
              noise_scale_w=0.8, max_len=None, sdp_ratio=0, y=None, *, noise_w=None, noise_z=None, w_ceil_override=None, pcm16=False):
        """reference models.py:1026-1074.  Keyword-only extras (not in the reference): explicit noise tensors
        `noise_w` [B,2,T] / `noise_z` [B,inter,>=F] replacing the two in-model RNG draws (models.py:249, 1071), and
        `w_ceil_override` [B,T] to teacher-force durations in parity harnesses, `pcm16=True` to get `o` as int16 converted like
        the reference's callers do (gradio convert_to_16_bit_wav, webui.py:86).  `attn` comes back as a LazyAttn (see above)."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise Bv2Error("SynthesizerTrn.infer: module is on CPU; bert_vits2_b200 has no CPU path — call .to('cuda')")
        if x.dim() != 2 or bert.dim() != 3 or bert.shape[-1] != x.shape[1]:
            raise ValueError("expected x [B,T] and bert features [B,1024,T]")
        eng = self._engine(dev)
        B, T = x.shape
        if noise_w is None:  # same draw order/shape as the reference: SDP first (models.py:249)
            noise_w = torch.randn(B, 2, T, device=dev, dtype=torch.float32)
        y_lengths, F = eng.infer_begin(x, x_lengths, sid, tone, language, bert, ja_bert, en_bert, noise_w, noise_scale_w,
                                       length_scale, sdp_ratio, w_ceil_override)
        if noise_z is None:  # torch.randn_like(m_p), m_p: [B, inter, F] (models.py:1071)
            noise_z = torch.randn(B, self.inter_channels, F, device=dev, dtype=torch.float32)
        o, _, y_mask, aux = eng.infer_finish(B, T, F, noise_z, noise_scale, max_len, want_attn=False, pcm16=pcm16)
        eng._attn_token = token = object()
        self.last_y_lengths = y_lengths
        return o, LazyAttn(eng, (B, 1, F, T), token), y_mask, aux

    def forward(self, *a, **kw):
        raise NotImplementedError("training forward (reference models.py:937-1024) is out of scope; use .infer()")
