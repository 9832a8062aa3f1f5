"""GPU parity: the CUDA engine (through the C ABI / drop-in class) against the CPU oracle and against the
committed reference-generated goldens.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

from bert_vits2_b200 import synth
from bert_vits2_b200.spec import ModelConfig
from util import GOLDEN_CASES, case_inputs, load_golden, model_for, rms

pytestmark = pytest.mark.gpu

# Stated tolerances (fp32 path: different summation order only; tf32 / fp16 paths: 11-bit-significand operands, fp32 accumulate)
TOL_FP32 = 2e-4        # max-abs per stage, pre-Generator stages (values are O(1))
TOL_WAV_FP32 = 2e-5    # waveform RMS, fp32 Generator
TOL_WAV_TF32 = 1e-3    # waveform RMS, tf32 / fp16-operand tcgen05 Generator (north_star bar)
PRECISIONS = ["fp32", "tf32", "fp16g", "fp16"]


@pytest.fixture(scope="module")
def engines():
    from bert_vits2_b200.engine import Engine
    cache = {}

    def get(tflow: bool, precision: str):
        key = (tflow, precision)
        if key not in cache:
            cfg, sd = model_for(tflow, 0)
            cache[key] = Engine(cfg, sd, device="cuda:0", precision=precision)
        return cache[key]

    yield get
    cache.clear()


def _oracle_stages(meta):
    from oracle import vits2_oracle as O
    cfg, sd, inp, nw, nz, kw = case_inputs(meta)
    st = O.infer(sd, cfg, **inp, noise_w=nw, noise_z=nz, return_stages=True, **kw)
    return cfg, sd, inp, nw, nz, kw, st


@pytest.mark.parametrize("precision", PRECISIONS)  # every engine runs this stage on FP32 FMA (it feeds ceil(durations))
@pytest.mark.parametrize("name", ["tflow_b1", "tflow_b3"])
def test_text_encoder_stage(engines, name, precision):
    meta, gold = load_golden(name)
    cfg, sd, inp, nw, nz, kw = case_inputs(meta)
    eng = engines(True, precision)
    x, m, logs = eng.text_encoder(inp["x"], inp["x_lengths"], inp["sid"], inp["tone"], inp["language"], inp["bert"],
                                  inp["ja_bert"], inp["en_bert"])
    for got, key in ((x, "x"), (m, "m_p_tok"), (logs, "logs_p_tok")):
        err = float((got.cpu() - gold[key]).abs().max())
        print(f"[{name}/{precision}] {key} max-abs err {err:.2e}")
        assert err < TOL_FP32, (key, err)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", ["tflow_b1", "tflow_b3"])
def test_duration_stage(engines, name, precision):
    meta, gold = load_golden(name)
    cfg, sd, inp, nw, nz, kw = case_inputs(meta)
    eng = engines(True, precision)
    a, b = eng.duration(gold["x"], inp["x_lengths"], inp["sid"], nw, kw["noise_scale_w"])
    assert float((a.cpu() - gold["logw_sdp"]).abs().max()) < TOL_FP32
    assert float((b.cpu() - gold["logw_dp"]).abs().max()) < TOL_FP32


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_flow_stage(engines, name):
    meta, gold = load_golden(name)
    cfg, sd, inp, nw, nz, kw = case_inputs(meta)
    eng = engines(meta["use_transformer_flow"], "fp32")
    z = eng.flow_reverse(gold["z_p"], gold["y_lengths"], inp["sid"])
    err = float((z.cpu() - gold["z"]).abs().max())
    assert err < TOL_FP32, err


@pytest.mark.parametrize("precision", PRECISIONS)
def test_generator_stage(engines, precision):
    meta, gold = load_golden("tflow_b3")
    cfg, sd, inp, nw, nz, kw = case_inputs(meta)
    eng = engines(True, precision)
    g = torch.nn.functional.embedding(inp["sid"], sd["emb_g.weight"])
    zin = gold["z"] * gold["y_mask"]
    o = eng.generator(zin, g).cpu()
    assert o.shape == gold["o"].shape
    e = rms(o, gold["o"])
    assert e < (TOL_WAV_FP32 if precision == "fp32" else TOL_WAV_TF32), e


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_full_infer_vs_reference_golden(engines, name, precision):
    """Whole SynthesizerTrn.infer through the C ABI against outputs of the unmodified reference."""
    meta, gold = load_golden(name)
    cfg, sd, inp, nw, nz, kw = case_inputs(meta)
    eng = engines(meta["use_transformer_flow"], precision)
    B, T = inp["x"].shape
    ylen, F = eng.infer_begin(inp["x"], inp["x_lengths"], inp["sid"], inp["tone"], inp["language"], inp["bert"], inp["ja_bert"],
                              inp["en_bert"], nw, kw["noise_scale_w"], kw["length_scale"], kw["sdp_ratio"])
    w_ceil = eng.debug_read("w_ceil", (B, 1, T))
    flips = int((w_ceil != gold["w_ceil"]).sum())
    if flips:  # ceil() is discontinuous; teacher-force and report
        print(f"[{name}] {flips} duration flips; teacher-forcing reference w_ceil")
        ylen, F = eng.infer_begin(inp["x"], inp["x_lengths"], inp["sid"], inp["tone"], inp["language"], inp["bert"],
                                  inp["ja_bert"], inp["en_bert"], nw, kw["noise_scale_w"], kw["length_scale"], kw["sdp_ratio"],
                                  w_ceil_override=gold["w_ceil"][:, 0])
    assert flips <= 1
    assert ylen.tolist() == gold["y_lengths"].tolist()
    o, attn, y_mask, (z, z_p, m_p, logs_p) = eng.infer_finish(B, T, F, nz, kw["noise_scale"])
    assert torch.equal(y_mask.cpu(), gold["y_mask"])
    assert torch.equal(attn.cpu().sum(2), gold["w_ceil"])
    for got, key in ((m_p, "m_p"), (logs_p, "logs_p"), (z_p, "z_p")):
        assert float((got.cpu() - gold[key]).abs().max()) < TOL_FP32, key
    ztol = TOL_FP32 if precision == "fp32" else 5e-3
    assert float((z.cpu() - gold["z"]).abs().max()) < ztol
    e = rms(o.cpu(), gold["o"])
    print(f"[{name}/{precision}] waveform RMS err {e:.3e}")
    assert e < (5e-5 if precision == "fp32" else TOL_WAV_TF32), e


def test_dropin_class_like_get_net_g():
    """Construct / load / infer exactly as reference infer.get_net_g + infer.infer do (infer.py:95-104, 302-318)."""
    from bert_vits2_b200.models import SynthesizerTrn
    from oracle import vits2_oracle as O
    cfg, sd = model_for(True, 0)
    hps_model = dict(inter_channels=192, hidden_channels=192, filter_channels=768, n_heads=2, n_layers=6, kernel_size=3,
                     p_dropout=0.1, resblock="1", resblock_kernel_sizes=[3, 7, 11],
                     resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], upsample_rates=[8, 8, 2, 2, 2],
                     upsample_initial_channel=512, upsample_kernel_sizes=[16, 16, 8, 2, 2], n_layers_q=3, use_spectral_norm=False,
                     gin_channels=512, use_spk_conditioned_encoder=True, use_noise_scaled_mas=True, slm={"x": 1})
    net = SynthesizerTrn(112, 1025, 32, n_speakers=850, init_seed=None, precision="fp32", **hps_model).to("cuda:0")
    net.eval()
    missing = net.load_state_dict(sd, strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys
    inp = synth.synthetic_inputs(cfg, [19], [0], seed=11)
    dev = {k: v.to("cuda:0") for k, v in inp.items()}
    torch.manual_seed(123)
    o, attn, y_mask, (z, z_p, m_p, logs_p) = net.infer(dev["x"], dev["x_lengths"], dev["sid"], dev["tone"], dev["language"],
                                                       dev["bert"], dev["ja_bert"], dev["en_bert"], sdp_ratio=0.5,
                                                       noise_scale=0.6, noise_scale_w=0.9, length_scale=1.0)
    audio = o[0, 0].data.cpu().float().numpy()  # as infer.py:315-318
    assert audio.ndim == 1 and np.isfinite(audio).all()
    F = int(y_mask.sum())
    assert audio.shape[0] == F * 512 and attn.shape == (1, 1, F, 19)
    # same device RNG stream => reproduce the noise and check against the oracle
    torch.manual_seed(123)
    nw = torch.randn(1, 2, 19, device="cuda:0").cpu()
    nz = torch.randn(1, 192, F, device="cuda:0").cpu()
    ref, _, _, _ = O.infer(sd, cfg, **inp, noise_w=nw, noise_z=nz, sdp_ratio=0.5, noise_scale=0.6, noise_scale_w=0.9,
                           length_scale=1.0)
    assert ref.shape == o.shape and rms(o.cpu(), ref) < 5e-5


def test_max_len_and_cpu_module_raises():
    from bert_vits2_b200.models import SynthesizerTrn
    from bert_vits2_b200.engine import Bv2Error
    cfg, sd = model_for(True, 0)
    net = SynthesizerTrn(112, 1025, 32, 192, 192, 768, 2, 6, 3, 0.1, "1", [3, 7, 11], [[1, 3, 5]] * 3, [8, 8, 2, 2, 2], 512,
                         [16, 16, 8, 2, 2], n_speakers=850, gin_channels=512, init_seed=None, precision="fp32")
    net.load_state_dict(sd, strict=False)
    inp = synth.synthetic_inputs(cfg, [12], [2], seed=3)
    with pytest.raises(Bv2Error):
        net.infer(**inp)
    net = net.to("cuda:0")
    dev = {k: v.to("cuda:0") for k, v in inp.items()}
    torch.manual_seed(5)
    o_full, _, y_mask, _ = net.infer(**dev, sdp_ratio=0.2)
    torch.manual_seed(5)
    o_cut, _, _, _ = net.infer(**dev, sdp_ratio=0.2, max_len=10)
    assert o_cut.shape[-1] == 10 * 512
    # the Generator's receptive field (~14 frames/side) means only the early samples agree exactly
    assert torch.allclose(o_cut[..., : 2 * 512], o_full[..., : 2 * 512], atol=2e-2)
    with pytest.raises(ValueError):
        net.infer(dev["x"], dev["x_lengths"], dev["sid"], dev["tone"], dev["language"], dev["bert"][:, :, :5], dev["ja_bert"],
                  dev["en_bert"])


def test_single_token_and_determinism(engines):
    cfg, sd = model_for(True, 0)
    eng = engines(True, "fp32")
    inp = synth.synthetic_inputs(cfg, [1, 3], [0, 1], seed=4)
    nw, nz = synth.synthetic_noise(cfg, 2, 3, 256, seed=9)
    outs = []
    for _ in range(2):
        ylen, F = eng.infer_begin(inp["x"], inp["x_lengths"], inp["sid"], inp["tone"], inp["language"], inp["bert"], inp["ja_bert"],
                                  inp["en_bert"], nw, 0.9, 1.0, 0.5)
        o, *_ = eng.infer_finish(2, 3, F, nz, 0.6)
        outs.append(o.cpu())
    assert torch.equal(outs[0], outs[1]) and torch.isfinite(outs[0]).all()
    from oracle import vits2_oracle as O
    ref, _, _, _ = O.infer(sd, cfg, **inp, noise_w=nw, noise_z=nz, sdp_ratio=0.5, noise_scale=0.6, noise_scale_w=0.9, length_scale=1.0)
    assert ref.shape == outs[0].shape and rms(outs[0], ref) < 5e-5


def test_peer_slab_output_pointer_world1(engines):
    """SURVEY.md section 8e exchange step, single-process part: the engine stores the waveform batch through a raw output
    pointer into a libbv2-owned slab (the address a peer rank would have mapped with CUDA IPC), the meta record follows by
    bv2_peer_write, and the root-side views equal an ordinary infer() bit for bit.  Both the fused (out_ptr) and the
    API-level (copy of a finished tensor) variants, two slots."""
    from bert_vits2_b200.sharding import PeerWaveSlab
    cfg, sd = model_for(True, 0)
    eng = engines(True, "tf32")
    inp = synth.synthetic_inputs(cfg, [7, 12], [0, 2], seed=31)
    nw, nz = synth.synthetic_noise(cfg, 2, 12, 512, seed=32)
    ylen, F = eng.infer_begin(inp["x"], inp["x_lengths"], inp["sid"], inp["tone"], inp["language"], inp["bert"], inp["ja_bert"],
                              inp["en_bert"], nw, 0.9, 1.0, 0.5)
    o_ref, *_ = eng.infer_finish(2, 12, F, nz, 0.6)
    L = F * cfg.hop
    slab = PeerWaveSlab("cuda:0", 2, 2 * L, slots=2)
    try:
        assert slab.fits(2, L) and not slab.fits(3, L) and not slab.fits(2, 5 * L)
        ylen2, F2 = eng.infer_begin(inp["x"], inp["x_lengths"], inp["sid"], inp["tone"], inp["language"], inp["bert"], inp["ja_bert"],
                                    inp["en_bert"], nw, 0.9, 1.0, 0.5)
        assert F2 == F
        o_none, *_ = eng.infer_finish(2, 12, F, nz, 0.6, out_ptr=slab.wave_ptr(0))
        assert o_none is None
        slab.publish(0, 2, L, ylen2 * cfg.hop)
        slab.publish(1, 2, L, ylen * cfg.hop, wave=o_ref)
        for slot in (0, 1):
            waves, counts = slab.collect(slot)
            assert len(waves) == 1 and tuple(waves[0].shape) == (2, 1, L)
            assert torch.equal(waves[0], o_ref)
            assert counts[0].tolist() == [int(v) * cfg.hop for v in ylen]
        with pytest.raises(ValueError):
            slab.publish(0, 3, L, [1, 2, 3])
    finally:
        slab.close()


@pytest.mark.parametrize("precision", PRECISIONS)
def test_config2_full_size_against_oracle(engines, precision):
    """BASELINE.json config 2: B=1, 256-phoneme ZH utterance, full path; waveform RMS vs the CPU oracle < 1e-3."""
    from oracle import vits2_oracle as O
    cfg, sd = model_for(True, 0)
    eng = engines(True, precision)
    inp = synth.synthetic_inputs(cfg, [256], [0], seed=2)
    nw, nz = synth.synthetic_noise(cfg, 1, 256, 4096, seed=2)
    st = O.infer(sd, cfg, **inp, noise_w=nw, noise_z=nz, sdp_ratio=0.5, noise_scale=0.6, noise_scale_w=0.9, length_scale=1.0,
                 return_stages=True)
    ylen, F = eng.infer_begin(inp["x"], inp["x_lengths"], inp["sid"], inp["tone"], inp["language"], inp["bert"], inp["ja_bert"],
                              inp["en_bert"], nw, 0.9, 1.0, 0.5)
    w_ceil = eng.debug_read("w_ceil", (1, 1, 256))
    flips = int((w_ceil != st["w_ceil"]).sum())
    print(f"config2: frames ref {int(st['y_lengths'][0])} got {int(ylen[0])}, duration flips {flips}")
    if flips:
        ylen, F = eng.infer_begin(inp["x"], inp["x_lengths"], inp["sid"], inp["tone"], inp["language"], inp["bert"],
                                  inp["ja_bert"], inp["en_bert"], nw, 0.9, 1.0, 0.5, w_ceil_override=st["w_ceil"][:, 0])
    assert flips <= 2 and int(ylen[0]) == int(st["y_lengths"][0])
    o, *_ = eng.infer_finish(1, 256, F, nz, 0.6)
    e = rms(o.cpu(), st["o"])
    print(f"config2/{precision}: waveform RMS err {e:.3e} (signal RMS {float(st['o'].pow(2).mean().sqrt()):.3f})")
    assert e < (5e-5 if precision == "fp32" else TOL_WAV_TF32)


def test_infer_batch_api_matches_single_calls():
    """batched caller-side API; each utterance equals its own B=1 result except in the tail
    the padded-batch semantics of the reference touch (receptive field of flow + Generator)."""
    from bert_vits2_b200.infer_api import infer_batch
    from bert_vits2_b200.models import SynthesizerTrn
    cfg, sd = model_for(True, 0)
    net = SynthesizerTrn(112, 1025, 32, 192, 192, 768, 2, 6, 3, 0.1, "1", [3, 7, 11], [[1, 3, 5]] * 3, [8, 8, 2, 2, 2], 512,
                         [16, 16, 8, 2, 2], n_speakers=850, gin_channels=512, init_seed=None, precision="fp32")
    net.load_state_dict(sd, strict=False)
    net = net.to("cuda:0").eval()
    items = []
    for k, t in enumerate([9, 14, 11, 14]):
        inp = synth.synthetic_inputs(cfg, [t], [k % 3], seed=20 + k)
        items.append((inp["bert"][0], inp["ja_bert"][0], inp["en_bert"][0], inp["x"][0], inp["tone"][0], inp["language"][0]))
    outs = infer_batch(net, items, sid=0, batch_size=2, sdp_ratio=0.0, noise_scale=0.0, noise_scale_w=0.0)  # deterministic: no noise
    assert len(outs) == 4 and all(o.ndim == 1 and o.size % 512 == 0 and np.isfinite(o).all() for o in outs)
    for k, it in enumerate(items):
        d = {n: v.unsqueeze(0).to("cuda:0") for n, v in zip(("bert", "ja_bert", "en_bert", "x", "tone", "language"), it)}
        o1, _, ym, _ = net.infer(d["x"], torch.tensor([it[3].shape[0]], device="cuda:0"), torch.zeros(1, dtype=torch.int64, device="cuda:0"),
                                 d["tone"], d["language"], d["bert"], d["ja_bert"], d["en_bert"], sdp_ratio=0.0, noise_scale=0.0, noise_scale_w=0.0)
        ref = o1[0, 0].cpu().numpy()
        assert ref.shape == outs[k].shape
        body = slice(0, max(0, ref.size - 20 * 512))
        assert np.abs(ref[body] - outs[k][body]).max() < 1e-4 if ref.size > 20 * 512 else True


def test_long_utterance_T512_and_batch8(engines):
    """BASELINE.json config 4 upper bound (512 phonemes, ~3k frames) and a B=8 ragged batch: full path vs the CPU oracle."""
    from oracle import vits2_oracle as O
    cfg, sd = model_for(True, 0)
    eng = engines(True, "tf32")
    for lengths, langs, seed in (([512], [0], 31), ([64, 57, 33, 64, 12, 40, 64, 25], [0, 1, 2, 0, 1, 2, 0, 1], 32)):
        B, T = len(lengths), max(lengths)
        inp = synth.synthetic_inputs(cfg, lengths, langs, seed=seed)
        nw, nz = synth.synthetic_noise(cfg, B, T, 8192, seed=seed)
        st = O.infer(sd, cfg, **inp, noise_w=nw, noise_z=nz, sdp_ratio=0.5, noise_scale=0.6, noise_scale_w=0.9, length_scale=1.0,
                     return_stages=True)
        ylen, F = eng.infer_begin(inp["x"], inp["x_lengths"], inp["sid"], inp["tone"], inp["language"], inp["bert"], inp["ja_bert"],
                                  inp["en_bert"], nw, 0.9, 1.0, 0.5)
        w_ceil = eng.debug_read("w_ceil", (B, 1, T))
        flips = int((w_ceil != st["w_ceil"]).sum())
        if flips:
            ylen, F = eng.infer_begin(inp["x"], inp["x_lengths"], inp["sid"], inp["tone"], inp["language"], inp["bert"], inp["ja_bert"],
                                      inp["en_bert"], nw, 0.9, 1.0, 0.5, w_ceil_override=st["w_ceil"][:, 0])
        assert flips <= 2 and ylen.tolist() == st["y_lengths"].tolist()
        o, attn, y_mask, (z, z_p, m_p, logs_p) = eng.infer_finish(B, T, F, nz, 0.6)
        e = rms(o.cpu(), st["o"])
        print(f"B={B} T={T} F={F}: duration flips {flips}, waveform RMS err {e:.3e}, z max err {float((z.cpu() - st['z']).abs().max()):.2e}")
        assert torch.isfinite(o).all() and e < TOL_WAV_TF32
