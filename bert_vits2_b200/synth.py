"""Seeded synthetic checkpoints, inputs and noise for the VITS2 infer path.

There is no network for trained checkpoints or datasets, so parity tests and bench.py use synthetic
state_dicts with the reference's key names/shapes (spec.py) and synthetic phoneme/BERT inputs of the
shapes BASELINE.json names.  Everything is drawn from numpy's Philox generator so the
same bits are produced in the build container and on the GPU box (torch's CPU RNG stream is not
guaranteed stable across builds).

Deviations from the reference's default init, on purpose:
  * layers the reference zero-initialises (ConvFlow.proj modules.py:483-484, coupling post
    modules.py:434-435/558-559, ElementwiseAffine modules.py:388-389) are re-drawn from N(0, 0.05^2)
    so that splines and couplings are data dependent;
  * dp.proj.bias = log(fpt), sdp.flows.0.m[0] = -log(fpt): mean duration ~fpt frames per token.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .spec import ModelConfig, param_specs


def _rng(seed: int, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed, stream]))


def synthetic_state_dict(cfg: ModelConfig, seed: int = 0, frames_per_token: float = 4.0) -> Dict[str, torch.Tensor]:
    """fp32 CPU state_dict with the reference's keys (weight-norm kept as weight_g / weight_v)."""
    specs = param_specs(cfg)
    sd: Dict[str, np.ndarray] = {}
    pending_g = []
    for idx, p in enumerate(specs):
        r = _rng(seed, idx + 1)
        shp = p.shape
        if p.init == "conv":  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
            fan_in = int(np.prod(shp[1:]))
            b = 1.0 / math.sqrt(fan_in)
            a = r.uniform(-b, b, size=shp)
        elif p.init in ("conv_b",):
            b = 1.0 / math.sqrt(max(1, p.aux.get("fan_in", 1)))
            a = r.uniform(-b, b, size=shp)
        elif p.init == "wn_v":
            std = p.aux.get("std")
            if std is None:
                fan_in = int(np.prod(shp[1:]))
                b = 1.0 / math.sqrt(fan_in)
                a = r.uniform(-b, b, size=shp)
            else:
                a = r.normal(0.0, std, size=shp)
        elif p.init == "wn_g":
            pending_g.append((p, idx))
            continue
        elif p.init == "ln_g":
            a = 1.0 + 0.1 * r.normal(size=shp)
        elif p.init == "ln_b":
            a = 0.1 * r.normal(size=shp)
        elif p.init == "emb":
            a = r.normal(0.0, shp[1] ** -0.5, size=shp)
        elif p.init == "rel":
            a = r.normal(0.0, shp[2] ** -0.5, size=shp)
        elif p.init == "small":
            a = 0.1 * r.normal(size=shp)
        elif p.init in ("zero_rerand", "zero_rerand_b"):
            a = 0.05 * r.normal(size=shp)
        elif p.init == "spk":
            a = r.normal(size=shp)
        else:
            raise ValueError(p.init)
        sd[p.key] = a.astype(np.float32)
    for p, idx in pending_g:
        # weight_norm init: g = ||v|| over every dim except 0, then perturbed so g != ||v||.
        v = sd[p.aux["of"]]
        n = np.sqrt((v.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
        r = _rng(seed, idx + 1)
        gain = p.aux.get("gain")
        if gain is not None:
            # The reference's N(0, 0.01) init (commons.init_weights) crushes the signal path of a random-init
            # Generator (output becomes bias-only, i.e. input independent).  Rescale g so the folded weight is
            # variance preserving: std(w) = gain / sqrt(fan_in_eff).
            v_rms = float(np.sqrt((v.astype(np.float64) ** 2).mean()))
            n = n * (gain / (v_rms * math.sqrt(p.aux["fan_in_eff"])))
        sd[p.key] = (n * (1.0 + 0.1 * r.normal(size=n.shape))).astype(np.float32)
    out = {p.key: torch.from_numpy(np.ascontiguousarray(sd[p.key])) for p in specs}
    lf = math.log(frames_per_token)
    out["dp.proj.bias"] = torch.full((1,), lf, dtype=torch.float32)
    m = out["sdp.flows.0.m"].clone()
    m[0, 0] = -lf
    out["sdp.flows.0.m"] = m
    return out


_TONE_RANGES = {0: (0, 6), 1: (6, 8), 2: (8, 12)}  # ZH / JP / EN, reference text/symbols.py:100,152-172


def synthetic_inputs(cfg: ModelConfig, lengths, languages=None, seed: int = 1, sid: int = 0):
    """Padded batch of synthetic get_text() outputs (reference infer.py:107-148 shapes).

    Returns dict of CPU tensors: x,tone,language [B,T] i64, x_lengths [B] i64, sid [B] i64,
    bert/ja_bert/en_bert [B,1024,T] f32 (N(0,1) everywhere, as get_text fills unused slots with randn).
    Padding positions hold zeros.
    """
    lengths = [int(t) for t in lengths]
    B, T = len(lengths), max(lengths)
    if languages is None:
        languages = [0] * B
    r = _rng(seed, 0)
    x = np.zeros((B, T), np.int64)
    tone = np.zeros((B, T), np.int64)
    lang = np.zeros((B, T), np.int64)
    berts = np.zeros((3, B, cfg.bert_dim, T), np.float32)
    for b, (t, lg) in enumerate(zip(lengths, languages)):
        x[b, :t] = r.integers(0, cfg.n_vocab, size=t)
        lo, hi = _TONE_RANGES[lg]
        tone[b, :t] = r.integers(lo, hi, size=t)
        lang[b, :t] = lg
        berts[:, b, :, :t] = r.standard_normal(size=(3, cfg.bert_dim, t), dtype=np.float32)
    return {
        "x": torch.from_numpy(x),
        "x_lengths": torch.tensor(lengths, dtype=torch.int64),
        "sid": torch.full((B,), sid, dtype=torch.int64),
        "tone": torch.from_numpy(tone),
        "language": torch.from_numpy(lang),
        "bert": torch.from_numpy(berts[0]),
        "ja_bert": torch.from_numpy(berts[1]),
        "en_bert": torch.from_numpy(berts[2]),
    }


def synthetic_noise(cfg: ModelConfig, B: int, T: int, F_cap: int, seed: int = 2):
    """Explicit noise for the two RNG draws of infer(): randn(B,2,T) (reference models.py:249) and
    randn_like(m_p)[B,192,F] (models.py:1071).  noise_z is drawn at capacity F_cap; infer uses [:, :, :F]."""
    r = _rng(seed, 0)
    nw = r.standard_normal(size=(B, 2, T), dtype=np.float32)
    r2 = _rng(seed, 1)
    nz = r2.standard_normal(size=(B, cfg.inter_channels, F_cap), dtype=np.float32)
    return torch.from_numpy(nw), torch.from_numpy(nz)


def synthetic_generator_inputs(cfg: ModelConfig, B: int, F: int, seed: int = 5):
    """Config 5 (Generator-only microbench): z ~ N(0,1)[B,192,F], g ~ N(0,1)[B,512,1]."""
    r = _rng(seed, 0)
    z = r.standard_normal(size=(B, cfg.inter_channels, F), dtype=np.float32)
    g = r.standard_normal(size=(B, cfg.gin_channels, 1), dtype=np.float32)
    return torch.from_numpy(z), torch.from_numpy(g)
