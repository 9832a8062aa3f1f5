"""Parameter specification of the VITS2 synthesizer (Bert-VITS2 v2.3) for the infer() path.

The engine must accept the reference's checkpoints: `utils.load_checkpoint` (reference utils.py:65-120)
walks `model.state_dict()` and then calls `model.load_state_dict(new, strict=False)`.  This module lists
every state_dict key the reference `models.SynthesizerTrn` owns on the inference path, with its shape,
in the reference's registration order (reference models.py:816-935).  `enc_q.*` (PosteriorEncoder,
training only, reference models.py:448-487) is deliberately absent: released checkpoints drop it
(reference compress_model.py:44-53) and infer() never touches it.

Each entry is (key, shape, init) where `init` is only used by synth.py to build seeded synthetic
checkpoints (there is no network to fetch trained ones).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple


@dataclass
class ModelConfig:
    """`hps.model` + the positional ctor args of reference models.SynthesizerTrn.__init__ (models.py:816-841)."""

    n_vocab: int = 112  # len(text.symbols.symbols), reference text/symbols.py:168
    num_tones: int = 12  # reference text/symbols.py:172
    num_languages: int = 3  # reference text/symbols.py:176
    bert_dim: int = 1024  # reference models.py:366-368
    inter_channels: int = 192
    hidden_channels: int = 192
    filter_channels: int = 768
    n_heads: int = 2
    n_layers: int = 6
    kernel_size: int = 3
    window_size: int = 4  # attentions.Encoder default, reference attentions.py:46
    resblock: str = "1"
    resblock_kernel_sizes: Tuple[int, ...] = (3, 7, 11)
    resblock_dilation_sizes: Tuple[Tuple[int, ...], ...] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    upsample_rates: Tuple[int, ...] = (8, 8, 2, 2, 2)
    upsample_initial_channel: int = 512
    upsample_kernel_sizes: Tuple[int, ...] = (16, 16, 8, 2, 2)
    n_speakers: int = 850
    gin_channels: int = 512
    use_sdp: bool = True
    n_flow_layer: int = 4
    n_layers_trans_flow: int = 4
    use_transformer_flow: bool = True
    flow_kernel_size: int = 5  # hard-wired in reference models.py:905, 917
    wn_layers: int = 4  # ResidualCouplingBlock n_layers = n_flow_layer (reference models.py:918-919)
    sdp_filter: int = 192  # forced to in_channels, reference models.py:159
    sdp_kernel: int = 3
    sdp_n_flows: int = 4
    sdp_dds_layers: int = 3
    sdp_num_bins: int = 10
    sdp_tail_bound: float = 5.0
    dp_filter: int = 256
    dp_kernel: int = 3
    cond_layer_idx: int = 2  # reference attentions.py:67-69
    sampling_rate: int = 44100
    hop_length: int = 512

    @staticmethod
    def from_hps_model(model: dict, n_vocab: int = 112, n_speakers: int = 850, **extra) -> "ModelConfig":
        c = ModelConfig()
        c.n_vocab = n_vocab
        c.n_speakers = n_speakers
        for k, v in dict(model).items():
            if hasattr(c, k):
                if isinstance(v, list):
                    v = tuple(tuple(x) if isinstance(x, list) else x for x in v)
                setattr(c, k, v)
        for k, v in extra.items():
            if hasattr(c, k):
                setattr(c, k, v)
        c.wn_layers = c.n_flow_layer
        return c

    @property
    def n_flows(self) -> int:
        """Couplings in `flow`: TransformerCouplingBlock builds n_flow_layer of them (reference models.py:82-145);
        ResidualCouplingBlock is constructed as (inter, hidden, 5, 1, n_flow_layer, gin_channels=...) so n_flow_layer lands
        in its n_layers (WN depth) and n_flows keeps its default 4 (reference models.py:403-445, 918-919)."""
        return self.n_flow_layer if self.use_transformer_flow else 4

    @property
    def hop(self) -> int:
        h = 1
        for u in self.upsample_rates:
            h *= u
        return h


@dataclass
class ParamSpec:
    key: str
    shape: Tuple[int, ...]
    init: str  # conv | convT_v | wn_v | wn_g | wn_gT | ln_g | ln_b | emb | rel | small | zero_rerand | spk | special
    aux: dict = field(default_factory=dict)


def _conv(out: List[ParamSpec], name: str, co: int, ci: int, k: int, bias: bool = True, init: str = "conv"):
    out.append(ParamSpec(f"{name}.weight", (co, ci, k), init))
    if bias:
        out.append(ParamSpec(f"{name}.bias", (co,), init + "_b", {"fan_in": ci * k}))


def _linear(out, name, co, ci):
    out.append(ParamSpec(f"{name}.weight", (co, ci), "conv"))
    out.append(ParamSpec(f"{name}.bias", (co,), "conv_b", {"fan_in": ci}))


def _wn_conv(out, name, co, ci, k, std=None, gain=None):
    # torch.nn.utils.weight_norm registers bias, weight_g, weight_v in this order.
    out.append(ParamSpec(f"{name}.bias", (co,), "conv_b", {"fan_in": ci * k}))
    aux = {"of": f"{name}.weight_v"}
    if gain is not None:
        aux.update(gain=gain, fan_in_eff=ci * k)
    out.append(ParamSpec(f"{name}.weight_g", (co, 1, 1), "wn_g", aux))
    out.append(ParamSpec(f"{name}.weight_v", (co, ci, k), "wn_v", {"std": std}))


def _ln(out, name, c, gname="gamma", bname="beta"):
    out.append(ParamSpec(f"{name}.{gname}", (c,), "ln_g"))
    out.append(ParamSpec(f"{name}.{bname}", (c,), "ln_b"))


def _encoder(out, name, cfg: ModelConfig, n_layers: int, kernel: int):
    """attentions.Encoder parameters (reference attentions.py:37-101)."""
    H, Fc, dk = cfg.hidden_channels, cfg.filter_channels, cfg.hidden_channels // cfg.n_heads
    _linear(out, f"{name}.spk_emb_linear", H, cfg.gin_channels)
    for i in range(n_layers):
        a = f"{name}.attn_layers.{i}"
        out.append(ParamSpec(f"{a}.emb_rel_k", (1, 2 * cfg.window_size + 1, dk), "rel"))
        out.append(ParamSpec(f"{a}.emb_rel_v", (1, 2 * cfg.window_size + 1, dk), "rel"))
        for nm in ("conv_q", "conv_k", "conv_v", "conv_o"):
            _conv(out, f"{a}.{nm}", H, H, 1)
    for i in range(n_layers):
        _ln(out, f"{name}.norm_layers_1.{i}", H)
    for i in range(n_layers):
        _conv(out, f"{name}.ffn_layers.{i}.conv_1", Fc, H, kernel)
        _conv(out, f"{name}.ffn_layers.{i}.conv_2", H, Fc, kernel)
    for i in range(n_layers):
        _ln(out, f"{name}.norm_layers_2.{i}", H)


def _dds(out, name, c, k, n_layers):
    """modules.DDSConv parameters (reference modules.py:89-116)."""
    for i in range(n_layers):
        out.append(ParamSpec(f"{name}.convs_sep.{i}.weight", (c, 1, k), "conv"))
        out.append(ParamSpec(f"{name}.convs_sep.{i}.bias", (c,), "conv_b", {"fan_in": k}))
    for i in range(n_layers):
        _conv(out, f"{name}.convs_1x1.{i}", c, c, 1)
    for i in range(n_layers):
        _ln(out, f"{name}.norms_1.{i}", c)
    for i in range(n_layers):
        _ln(out, f"{name}.norms_2.{i}", c)


def _sdp_flows(out, name, cfg: ModelConfig):
    """[ElementwiseAffine(2)] + n_flows x [ConvFlow, Flip] (reference models.py:167-174, 181-187)."""
    out.append(ParamSpec(f"{name}.0.m", (2, 1), "small"))
    out.append(ParamSpec(f"{name}.0.logs", (2, 1), "small"))
    for i in range(cfg.sdp_n_flows):
        f = f"{name}.{1 + 2 * i}"
        _conv(out, f"{f}.pre", cfg.sdp_filter, 1, 1)
        _dds(out, f"{f}.convs", cfg.sdp_filter, cfg.sdp_kernel, cfg.sdp_dds_layers)
        nb = 3 * cfg.sdp_num_bins - 1
        out.append(ParamSpec(f"{f}.proj.weight", (nb, cfg.sdp_filter, 1), "zero_rerand"))
        out.append(ParamSpec(f"{f}.proj.bias", (nb,), "zero_rerand"))


def param_specs(cfg: ModelConfig) -> List[ParamSpec]:
    out: List[ParamSpec] = []
    H, I, G = cfg.hidden_channels, cfg.inter_channels, cfg.gin_channels
    # ---- enc_p: TextEncoder (reference models.py:333-375)
    out.append(ParamSpec("enc_p.emb.weight", (cfg.n_vocab, H), "emb"))
    out.append(ParamSpec("enc_p.tone_emb.weight", (cfg.num_tones, H), "emb"))
    out.append(ParamSpec("enc_p.language_emb.weight", (cfg.num_languages, H), "emb"))
    for nm in ("bert_proj", "ja_bert_proj", "en_bert_proj"):
        _conv(out, f"enc_p.{nm}", H, cfg.bert_dim, 1)
    _encoder(out, "enc_p.encoder", cfg, cfg.n_layers, cfg.kernel_size)
    _conv(out, "enc_p.proj", 2 * I, H, 1)
    # ---- dec: Generator (reference models.py:490-536)
    C0 = cfg.upsample_initial_channel
    _conv(out, "dec.conv_pre", C0, I, 7)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        ci, co = C0 >> i, C0 >> (i + 1)
        # ConvTranspose1d weight is [in, out, k]; weight_norm dim=0 -> g is per INPUT channel.
        out.append(ParamSpec(f"dec.ups.{i}.bias", (co,), "conv_b", {"fan_in": ci * k // u}))
        out.append(ParamSpec(f"dec.ups.{i}.weight_g", (ci, 1, 1), "wn_g",
                             {"of": f"dec.ups.{i}.weight_v", "gain": 1.0, "fan_in_eff": ci * k // u}))
        out.append(ParamSpec(f"dec.ups.{i}.weight_v", (ci, co, k), "wn_v", {"std": 0.01}))
    for i in range(len(cfg.upsample_rates)):
        ch = C0 >> (i + 1)
        for j, (k, ds) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            rb = f"dec.resblocks.{i * len(cfg.resblock_kernel_sizes) + j}"
            for n in range(len(ds)):
                _wn_conv(out, f"{rb}.convs1.{n}", ch, ch, k, std=0.01, gain=0.8)
            for n in range(len(ds)):
                _wn_conv(out, f"{rb}.convs2.{n}", ch, ch, k, std=0.01, gain=0.8)
    out.append(ParamSpec("dec.conv_post.weight", (1, C0 >> len(cfg.upsample_rates), 7), "conv"))
    _conv(out, "dec.cond", C0, G, 1)
    # ---- flow (reference models.py:903-926)
    half = I // 2
    for i in range(cfg.n_flows):
        f = f"flow.flows.{2 * i}"
        _conv(out, f"{f}.pre", H, half, 1)
        if cfg.use_transformer_flow:
            _encoder(out, f"{f}.enc", cfg, cfg.n_layers_trans_flow, cfg.flow_kernel_size)
        else:
            # modules.WN (reference modules.py:133-183): cond_layer, then (in_layer_i, res_skip_i) per layer.
            # Registration order: in_layers ModuleList, res_skip_layers ModuleList, cond_layer.
            L = cfg.wn_layers
            for n in range(L):
                _wn_conv(out, f"{f}.enc.in_layers.{n}", 2 * H, H, cfg.flow_kernel_size)
            for n in range(L):
                rs = 2 * H if n < L - 1 else H
                _wn_conv(out, f"{f}.enc.res_skip_layers.{n}", rs, H, 1)
            _wn_conv(out, f"{f}.enc.cond_layer", 2 * H * L, G, 1)
        out.append(ParamSpec(f"{f}.post.weight", (half, H, 1), "zero_rerand"))
        out.append(ParamSpec(f"{f}.post.bias", (half,), "zero_rerand"))
    # ---- sdp (reference models.py:148-195)
    _sdp_flows(out, "sdp.flows", cfg)
    _conv(out, "sdp.post_pre", cfg.sdp_filter, 1, 1)
    _conv(out, "sdp.post_proj", cfg.sdp_filter, cfg.sdp_filter, 1)
    _dds(out, "sdp.post_convs", cfg.sdp_filter, cfg.sdp_kernel, cfg.sdp_dds_layers)
    _sdp_flows(out, "sdp.post_flows", cfg)
    _conv(out, "sdp.pre", cfg.sdp_filter, H, 1)
    _conv(out, "sdp.proj", cfg.sdp_filter, cfg.sdp_filter, 1)
    _dds(out, "sdp.convs", cfg.sdp_filter, cfg.sdp_kernel, cfg.sdp_dds_layers)
    _conv(out, "sdp.cond", cfg.sdp_filter, G, 1)
    # ---- dp (reference models.py:259-283)
    _conv(out, "dp.conv_1", cfg.dp_filter, H, cfg.dp_kernel)
    _ln(out, "dp.norm_1", cfg.dp_filter)
    _conv(out, "dp.conv_2", cfg.dp_filter, cfg.dp_filter, cfg.dp_kernel)
    _ln(out, "dp.norm_2", cfg.dp_filter)
    _conv(out, "dp.proj", 1, cfg.dp_filter, 1)
    _conv(out, "dp.cond", H, G, 1)
    out.append(ParamSpec("emb_g.weight", (cfg.n_speakers, G), "spk"))
    return out


def param_shapes(cfg: ModelConfig) -> Dict[str, Tuple[int, ...]]:
    return {p.key: p.shape for p in param_specs(cfg)}


#: keys consumed by infer(); the rest (sdp.post_*, sdp.flows.1.* i.e. the dropped "useless vflow",
#: reference models.py:247) are carried only for state_dict compatibility.
def is_infer_key(key: str) -> bool:
    if key.startswith(("sdp.post_", "enc_q.")):
        return False
    if key.startswith("sdp.flows.1."):
        return False
    return True
