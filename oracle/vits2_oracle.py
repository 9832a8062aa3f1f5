"""CPU oracle: a functional fp32 restatement of Bert-VITS2 v2.3 `SynthesizerTrn.infer()`.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this module; the product path (bert_vits2_b200/)
never does and fails loudly if its CUDA library is missing.

Parity status: the reference ships NO tests or golden vectors for this path, so the
oracle is pinned against outputs of the unmodified reference itself, generated in the build container
by tests/golden/make_golden.py (fixtures in tests/golden/*.npz) and re-checked by
oracle/validate_against_reference.py.  tests/test_oracle_golden.py asserts the agreement on CPU.

Every function cites the reference file:line it restates.  It works on a plain state_dict with the
reference's key names (weight-norm kept as weight_g/weight_v and folded on every call, exactly as the
reference serves it: nobody calls remove_weight_norm).  Noise is an explicit input
(the reference draws it inside the model: models.py:249, 1071).

Arithmetic: fp32 throughout, torch CPU (ATen) ops — the same third-party library the reference's own
arithmetic executes in.  Relative-position attention is restated in its banded form
(|j-i| <= window) instead of the reference's pad/reshape skew trick; both are algebraically identical
 and the difference is covered by the golden check.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # reference modules.py:14


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
def wn_weight(sd, name: str) -> torch.Tensor:
    """torch.nn.utils.weight_norm, dim=0: w = g * v / ||v||_{dims != 0} (applied at reference
    models.py:513, modules.py:160,172,182,226-292).  For ConvTranspose1d dim 0 is in_channels."""
    v, g = sd[name + ".weight_v"], sd[name + ".weight_g"]
    return torch._weight_norm(v, g, 0)


def sequence_mask(length: torch.Tensor, max_length: Optional[int] = None) -> torch.Tensor:
    """reference commons.py:119-123"""
    if max_length is None:
        max_length = int(length.max())
    x = torch.arange(max_length, dtype=length.dtype)
    return x.unsqueeze(0) < length.unsqueeze(1)


def layer_norm_c(x: torch.Tensor, gamma, beta, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over the channel dim of [B,C,T] (reference modules.py:26-29 == attentions.py:21-24)."""
    return F.layer_norm(x.transpose(1, -1), (x.shape[1],), gamma, beta, eps).transpose(1, -1)


def conv1d(sd, name, x, dilation=1, padding=0, groups=1):
    return F.conv1d(x, sd[name + ".weight"], sd.get(name + ".bias"), 1, padding, dilation, groups)


# --------------------------------------------------------------------------------------------------
# attentions.py
# --------------------------------------------------------------------------------------------------
def mha_rel(sd, name: str, x: torch.Tensor, attn_mask: torch.Tensor, n_heads: int, window: int) -> torch.Tensor:
    """MultiHeadAttention.forward + .attention with windowed relative positions
    (reference attentions.py:263-322; relative terms :285-290, :311-318 in banded form)."""
    B, C, T = x.shape
    dk = C // n_heads
    q = conv1d(sd, name + ".conv_q", x).view(B, n_heads, dk, T).transpose(2, 3)  # [B,h,T,dk]
    k = conv1d(sd, name + ".conv_k", x).view(B, n_heads, dk, T).transpose(2, 3)
    v = conv1d(sd, name + ".conv_v", x).view(B, n_heads, dk, T).transpose(2, 3)
    qs = q / math.sqrt(dk)
    scores = torch.matmul(qs, k.transpose(-2, -1))  # [B,h,T,T]
    rel_k = sd[name + ".emb_rel_k"][0]  # [2w+1, dk]; heads_share=True
    rel_v = sd[name + ".emb_rel_v"][0]
    rel_logits = torch.matmul(qs, rel_k.t())  # [B,h,T,2w+1] : q_i . E_k[r], r = j-i+w
    idx = torch.arange(T)
    for r in range(2 * window + 1):
        off = r - window
        i = idx[(idx + off >= 0) & (idx + off < T)]
        scores[:, :, i, i + off] = scores[:, :, i, i + off] + rel_logits[:, :, i, r]
    scores = scores.masked_fill(attn_mask == 0, -1e4)  # reference attentions.py:297
    p = F.softmax(scores, dim=-1)
    out = torch.matmul(p, v)  # [B,h,T,dk]
    for r in range(2 * window + 1):
        off = r - window
        i = idx[(idx + off >= 0) & (idx + off < T)]
        out[:, :, i, :] = out[:, :, i, :] + p[:, :, i, i + off].unsqueeze(-1) * rel_v[r]
    out = out.transpose(2, 3).contiguous().view(B, C, T)
    return conv1d(sd, name + ".conv_o", out)


def ffn(sd, name: str, x, x_mask, k: int):
    """FFN.forward, activation=None -> ReLU, same padding (reference attentions.py:438-464)."""
    pl, pr = (k - 1) // 2, k // 2
    h = conv1d(sd, name + ".conv_1", F.pad(x * x_mask, (pl, pr)))
    h = torch.relu(h)
    h = conv1d(sd, name + ".conv_2", F.pad(h * x_mask, (pl, pr)))
    return h * x_mask


def encoder(sd, name: str, x, x_mask, g, n_layers: int, n_heads: int, window: int, k: int, cond_layer_idx: int = 2):
    """attentions.Encoder.forward (reference attentions.py:103-120)."""
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    x = x * x_mask
    for i in range(n_layers):
        if i == cond_layer_idx and g is not None:
            gg = F.linear(g.transpose(1, 2), sd[name + ".spk_emb_linear.weight"], sd[name + ".spk_emb_linear.bias"])
            x = (x + gg.transpose(1, 2)) * x_mask
        y = mha_rel(sd, f"{name}.attn_layers.{i}", x, attn_mask, n_heads, window)
        x = layer_norm_c(x + y, sd[f"{name}.norm_layers_1.{i}.gamma"], sd[f"{name}.norm_layers_1.{i}.beta"])
        y = ffn(sd, f"{name}.ffn_layers.{i}", x, x_mask, k)
        x = layer_norm_c(x + y, sd[f"{name}.norm_layers_2.{i}.gamma"], sd[f"{name}.norm_layers_2.{i}.beta"])
    return x * x_mask


# --------------------------------------------------------------------------------------------------
# models.py : TextEncoder
# --------------------------------------------------------------------------------------------------
def text_encoder(sd, cfg, x, x_lengths, tone, language, bert, ja_bert, en_bert, g):
    """TextEncoder.forward (reference models.py:377-400)."""
    H = cfg.hidden_channels
    e = (
        F.embedding(x, sd["enc_p.emb.weight"])
        + F.embedding(tone, sd["enc_p.tone_emb.weight"])
        + F.embedding(language, sd["enc_p.language_emb.weight"])
        + conv1d(sd, "enc_p.bert_proj", bert).transpose(1, 2)
        + conv1d(sd, "enc_p.ja_bert_proj", ja_bert).transpose(1, 2)
        + conv1d(sd, "enc_p.en_bert_proj", en_bert).transpose(1, 2)
    ) * math.sqrt(H)
    h = e.transpose(1, -1)
    x_mask = sequence_mask(x_lengths, h.size(2)).unsqueeze(1).to(h.dtype)
    h = encoder(sd, "enc_p.encoder", h * x_mask, x_mask, g, cfg.n_layers, cfg.n_heads, cfg.window_size,
                cfg.kernel_size, cfg.cond_layer_idx)
    stats = conv1d(sd, "enc_p.proj", h) * x_mask
    m, logs = torch.split(stats, cfg.inter_channels, dim=1)
    return h, m, logs, x_mask


# --------------------------------------------------------------------------------------------------
# modules.py : DDSConv, ConvFlow ; transforms.py : rational-quadratic spline (inverse)
# --------------------------------------------------------------------------------------------------
def dds_conv(sd, name: str, x, x_mask, g, k: int, n_layers: int):
    """DDSConv.forward (reference modules.py:118-130); F.gelu default = exact erf form."""
    if g is not None:
        x = x + g
    for i in range(n_layers):
        d = k ** i
        pad = (k * d - d) // 2
        y = conv1d(sd, f"{name}.convs_sep.{i}", x * x_mask, dilation=d, padding=pad, groups=x.shape[1])
        y = F.gelu(layer_norm_c(y, sd[f"{name}.norms_1.{i}.gamma"], sd[f"{name}.norms_1.{i}.beta"]))
        y = conv1d(sd, f"{name}.convs_1x1.{i}", y)
        y = F.gelu(layer_norm_c(y, sd[f"{name}.norms_2.{i}.gamma"], sd[f"{name}.norms_2.{i}.beta"]))
        x = x + y
    return x * x_mask


def rq_spline_inverse(inputs, uw, uh, ud, tail_bound: float = 5.0, min_bw=1e-3, min_bh=1e-3, min_d=1e-3):
    """piecewise_rational_quadratic_transform(inverse=True, tails='linear')
    (reference transforms.py:11-41 -> 49-96 -> 99-173).  Dense (no boolean-mask gather): the spline is
    evaluated everywhere with clamped inputs and the identity is selected outside [-B, B] (:61-74)."""
    nb = uw.shape[-1]
    inside = (inputs >= -tail_bound) & (inputs <= tail_bound)
    const = float(np.log(np.exp(1 - min_d) - 1))  # transforms.py:69
    ud = F.pad(ud, (1, 1))
    ud[..., 0] = const
    ud[..., -1] = const
    x = torch.where(inside, inputs, torch.zeros_like(inputs))
    left = bottom = -tail_bound
    right = top = tail_bound

    widths = F.softmax(uw, dim=-1)
    widths = min_bw + (1 - min_bw * nb) * widths
    cumwidths = F.pad(torch.cumsum(widths, dim=-1), (1, 0), value=0.0)
    cumwidths = (right - left) * cumwidths + left
    cumwidths[..., 0] = left
    cumwidths[..., -1] = right
    widths = cumwidths[..., 1:] - cumwidths[..., :-1]

    derivatives = min_d + F.softplus(ud)

    heights = F.softmax(uh, dim=-1)
    heights = min_bh + (1 - min_bh * nb) * heights
    cumheights = F.pad(torch.cumsum(heights, dim=-1), (1, 0), value=0.0)
    cumheights = (top - bottom) * cumheights + bottom
    cumheights[..., 0] = bottom
    cumheights[..., -1] = top
    heights = cumheights[..., 1:] - cumheights[..., :-1]

    locs = cumheights.clone()
    locs[..., -1] += 1e-6  # searchsorted eps, transforms.py:44-46 (in-place on cumheights in the reference)
    bin_idx = (torch.sum(x[..., None] >= locs, dim=-1) - 1)[..., None]
    cumheights = locs  # the reference's in-place "+= eps" leaks into cumheights used below (:151)

    in_cw = cumwidths.gather(-1, bin_idx)[..., 0]
    in_bw = widths.gather(-1, bin_idx)[..., 0]
    in_ch = cumheights.gather(-1, bin_idx)[..., 0]
    delta = heights / widths
    in_delta = delta.gather(-1, bin_idx)[..., 0]
    in_d = derivatives.gather(-1, bin_idx)[..., 0]
    in_d1 = derivatives[..., 1:].gather(-1, bin_idx)[..., 0]
    in_h = heights.gather(-1, bin_idx)[..., 0]

    a = (x - in_ch) * (in_d + in_d1 - 2 * in_delta) + in_h * (in_delta - in_d)
    b = in_h * in_d - (x - in_ch) * (in_d + in_d1 - 2 * in_delta)
    c = -in_delta * (x - in_ch)
    disc = b.pow(2) - 4 * a * c
    assert bool((disc[inside] >= 0).all())  # transforms.py:170
    root = (2 * c) / (-b - torch.sqrt(disc))
    out = root * in_bw + in_cw
    return torch.where(inside, out, inputs)


def conv_flow_reverse(sd, name: str, z, x_mask, g, cfg):
    """ConvFlow.forward(reverse=True) (reference modules.py:486-516)."""
    x0, x1 = torch.split(z, [1, 1], 1)
    h = conv1d(sd, name + ".pre", x0)
    h = dds_conv(sd, name + ".convs", h, x_mask, g, cfg.sdp_kernel, cfg.sdp_dds_layers)
    h = conv1d(sd, name + ".proj", h) * x_mask
    b, c, t = x0.shape
    h = h.reshape(b, c, -1, t).permute(0, 1, 3, 2)
    nb = cfg.sdp_num_bins
    s = math.sqrt(cfg.sdp_filter)
    x1 = rq_spline_inverse(x1, h[..., :nb] / s, h[..., nb:2 * nb] / s, h[..., 2 * nb:], cfg.sdp_tail_bound)
    return torch.cat([x0, x1], 1) * x_mask


def sdp_reverse(sd, cfg, x, x_mask, g, noise_w, noise_scale_w):
    """StochasticDurationPredictor.forward(reverse=True) (reference models.py:197-204, 245-256).
    noise_w [B,2,T] replaces torch.randn at :249."""
    h = conv1d(sd, "sdp.pre", x) + conv1d(sd, "sdp.cond", g)
    h = dds_conv(sd, "sdp.convs", h, x_mask, None, cfg.sdp_kernel, cfg.sdp_dds_layers)
    h = conv1d(sd, "sdp.proj", h) * x_mask
    z = noise_w.to(x.dtype) * noise_scale_w
    # reversed(flows) minus the "useless vflow" (flows[1]) : Flip,CF(2n-1),...,Flip,CF3,Flip, EA   (:246-247)
    n = cfg.sdp_n_flows
    for i in range(n, 1, -1):
        z = torch.flip(z, [1])  # Flip at index 2i
        z = conv_flow_reverse(sd, f"sdp.flows.{2 * i - 1}", z, x_mask, h, cfg)
    z = torch.flip(z, [1])  # Flip at index 2
    z = (z - sd["sdp.flows.0.m"]) * torch.exp(-sd["sdp.flows.0.logs"]) * x_mask  # ElementwiseAffine, modules.py:397-399
    return z[:, :1]


def duration_predictor(sd, cfg, x, x_mask, g):
    """DurationPredictor.forward (reference models.py:285-299); dropout is identity in eval."""
    p = cfg.dp_kernel // 2
    h = x + conv1d(sd, "dp.cond", g)
    h = torch.relu(conv1d(sd, "dp.conv_1", h * x_mask, padding=p))
    h = layer_norm_c(h, sd["dp.norm_1.gamma"], sd["dp.norm_1.beta"])
    h = torch.relu(conv1d(sd, "dp.conv_2", h * x_mask, padding=p))
    h = layer_norm_c(h, sd["dp.norm_2.gamma"], sd["dp.norm_2.beta"])
    return conv1d(sd, "dp.proj", h * x_mask) * x_mask


# --------------------------------------------------------------------------------------------------
# length regulation (models.py:1055-1071, commons.py:126-140)
# --------------------------------------------------------------------------------------------------
def generate_path(duration, mask):
    """commons.generate_path (reference commons.py:126-140)."""
    b, _, t_y, t_x = mask.shape
    cum = torch.cumsum(duration, -1).view(b * t_x)
    path = sequence_mask(cum, t_y).to(mask.dtype).view(b, t_x, t_y)
    path = path - F.pad(path, (0, 0, 1, 0))[:, :-1]
    return path.unsqueeze(1).transpose(2, 3) * mask


def length_regulate(logw, x_mask, m_p, logs_p, length_scale, w_ceil_override=None):
    """reference models.py:1055-1069.  Returns w_ceil, y_lengths, y_mask, attn, expanded m_p/logs_p."""
    w = torch.exp(logw) * x_mask * length_scale
    w_ceil = torch.ceil(w) if w_ceil_override is None else w_ceil_override
    y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
    y_mask = sequence_mask(y_lengths, None).unsqueeze(1).to(x_mask.dtype)
    attn_mask = x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)
    attn = generate_path(w_ceil, attn_mask)
    m_e = torch.matmul(attn.squeeze(1), m_p.transpose(1, 2)).transpose(1, 2)
    l_e = torch.matmul(attn.squeeze(1), logs_p.transpose(1, 2)).transpose(1, 2)
    return w_ceil, y_lengths, y_mask, attn, m_e, l_e


# --------------------------------------------------------------------------------------------------
# flows
# --------------------------------------------------------------------------------------------------
def transformer_coupling_reverse(sd, name, x, x_mask, g, cfg):
    """TransformerCouplingLayer.forward(reverse=True), mean_only (reference modules.py:561-580)."""
    half = cfg.inter_channels // 2
    x0, x1 = torch.split(x, [half, half], 1)
    h = conv1d(sd, name + ".pre", x0) * x_mask
    h = encoder(sd, name + ".enc", h, x_mask, g, cfg.n_layers_trans_flow, cfg.n_heads, cfg.window_size,
                cfg.flow_kernel_size, cfg.cond_layer_idx)
    m = conv1d(sd, name + ".post", h) * x_mask
    x1 = (x1 - m) * x_mask  # logs == 0 -> exp(-logs) == 1
    return torch.cat([x0, x1], 1)


def wn(sd, name, x, x_mask, g, cfg):
    """modules.WN.forward, dilation_rate=1 (reference modules.py:185-210) with
    commons.fused_add_tanh_sigmoid_multiply (commons.py:98-105)."""
    H, L, k = cfg.hidden_channels, cfg.wn_layers, cfg.flow_kernel_size
    out = torch.zeros_like(x)
    gg = F.conv1d(g, wn_weight(sd, name + ".cond_layer"), sd[name + ".cond_layer.bias"])
    for i in range(L):
        x_in = F.conv1d(x, wn_weight(sd, f"{name}.in_layers.{i}"), sd[f"{name}.in_layers.{i}.bias"], padding=(k - 1) // 2)
        a = x_in + gg[:, 2 * H * i: 2 * H * (i + 1)]
        acts = torch.tanh(a[:, :H]) * torch.sigmoid(a[:, H:])
        rs = F.conv1d(acts, wn_weight(sd, f"{name}.res_skip_layers.{i}"), sd[f"{name}.res_skip_layers.{i}.bias"])
        if i < L - 1:
            x = (x + rs[:, :H]) * x_mask
            out = out + rs[:, H:]
        else:
            out = out + rs
    return out * x_mask


def residual_coupling_reverse(sd, name, x, x_mask, g, cfg):
    """ResidualCouplingLayer.forward(reverse=True), mean_only (reference modules.py:437-456)."""
    half = cfg.inter_channels // 2
    x0, x1 = torch.split(x, [half, half], 1)
    h = conv1d(sd, name + ".pre", x0) * x_mask
    h = wn(sd, name + ".enc", h, x_mask, g, cfg)
    m = conv1d(sd, name + ".post", h) * x_mask
    x1 = (x1 - m) * x_mask
    return torch.cat([x0, x1], 1)


def flow_reverse(sd, cfg, z_p, y_mask, g):
    """{Transformer,Residual}CouplingBlock.forward(reverse=True) (reference models.py:142-145, 442-445):
    for flow in reversed([L0,Flip,L1,Flip,...]) -> Flip, L(n-1), Flip, L(n-2), ..., Flip, L0."""
    layer = transformer_coupling_reverse if cfg.use_transformer_flow else residual_coupling_reverse
    x = z_p
    for i in range(cfg.n_flows - 1, -1, -1):  # n_flow_layer (transformer flow) / 4 (WN flow: models.py:918-919 passes n_flow_layer as n_layers)
        x = torch.flip(x, [1])
        x = layer(sd, f"flow.flows.{2 * i}", x, y_mask, g, cfg)
    return x


# --------------------------------------------------------------------------------------------------
# Generator (HiFi-GAN), models.py:538-557 + modules.ResBlock1.forward modules.py:296-309
# --------------------------------------------------------------------------------------------------
def resblock1(sd, name, x, k, dils):
    for n, d in enumerate(dils):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, wn_weight(sd, f"{name}.convs1.{n}"), sd[f"{name}.convs1.{n}.bias"],
                      padding=(k * d - d) // 2, dilation=d)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, wn_weight(sd, f"{name}.convs2.{n}"), sd[f"{name}.convs2.{n}.bias"], padding=(k - 1) // 2)
        x = xt + x
    return x


def generator(sd, cfg, z, g):
    x = conv1d(sd, "dec.conv_pre", z, padding=3)
    if g is not None:
        x = x + conv1d(sd, "dec.cond", g)
    nk = len(cfg.resblock_kernel_sizes)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, wn_weight(sd, f"dec.ups.{i}"), sd[f"dec.ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            r = resblock1(sd, f"dec.resblocks.{i * nk + j}", x, cfg.resblock_kernel_sizes[j], cfg.resblock_dilation_sizes[j])
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01 (reference models.py:553)
    x = F.conv1d(x, sd["dec.conv_post.weight"], None, padding=3)
    return torch.tanh(x)


# --------------------------------------------------------------------------------------------------
# SynthesizerTrn.infer (reference models.py:1026-1074)
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def infer(sd: Dict[str, torch.Tensor], cfg, x, x_lengths, sid, tone, language, bert, ja_bert, en_bert,
          noise_w, noise_z, noise_scale=0.667, length_scale=1.0, noise_scale_w=0.8, max_len=None, sdp_ratio=0.0,
          w_ceil_override=None, return_stages=False):
    """noise_w [B,2,T] and noise_z [B,192,>=F] are the two RNG draws (models.py:249, 1071)."""
    g = F.embedding(sid, sd["emb_g.weight"]).unsqueeze(-1)  # models.py:1045-1046
    h, m_p, logs_p, x_mask = text_encoder(sd, cfg, x, x_lengths, tone, language, bert, ja_bert, en_bert, g)
    logw_sdp = sdp_reverse(sd, cfg, h, x_mask, g, noise_w, noise_scale_w)
    logw_dp = duration_predictor(sd, cfg, h, x_mask, g)
    logw = logw_sdp * sdp_ratio + logw_dp * (1 - sdp_ratio)  # models.py:1052-1054
    w_ceil, y_lengths, y_mask, attn, m_e, l_e = length_regulate(logw, x_mask, m_p, logs_p, length_scale, w_ceil_override)
    Fr = m_e.shape[2]
    z_p = m_e + noise_z[:, :, :Fr].to(m_e.dtype) * torch.exp(l_e) * noise_scale  # models.py:1071
    z = flow_reverse(sd, cfg, z_p, y_mask, g)
    o = generator(sd, cfg, (z * y_mask)[:, :, :max_len], g)
    if return_stages:
        return dict(o=o, attn=attn, y_mask=y_mask, z=z, z_p=z_p, m_p=m_e, logs_p=l_e, g=g, x=h, m_p_tok=m_p,
                    logs_p_tok=logs_p, x_mask=x_mask, logw_sdp=logw_sdp, logw_dp=logw_dp, logw=logw, w_ceil=w_ceil,
                    y_lengths=y_lengths)
    return o, attn, y_mask, (z, z_p, m_e, l_e)


# --------------------------------------------------------------------------------------------------
# 16-bit PCM conversion applied by the reference's callers to every infer() result (reference webui.py:86,
# 129, 198; hiyoriUI.py:343): gradio.processing_utils.convert_to_16_bit_wav.  gradio is a third-party
# dependency that is absent from /root/reference and from this image (requirements.txt pins gradio==3.50.2);
# its published float branch is restated here:   data = data / np.abs(data).max(); data = data * 32767;
# data = data.astype(np.int16)   -- float32 arithmetic throughout, astype truncates toward zero.
# --------------------------------------------------------------------------------------------------
def convert_to_16_bit_wav(data):
    import numpy as np
    data = np.asarray(data)
    assert data.dtype == np.float32
    data = data / np.abs(data).max()
    data = data * 32767
    return data.astype(np.int16)

