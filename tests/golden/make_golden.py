#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/<case>.npz holding the reference's outputs for seeded synthetic weights/inputs/noise
(bert_vits2_b200.synth; regenerated from the seeds at test time, so weights are NOT stored).
The reference's two RNG draws (models.py:249 `torch.randn(B,2,T)`, models.py:1071 `torch.randn_like(m_p)`)
are patched to return the seeded noise, in call order.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bert_vits2_b200.spec import ModelConfig  # noqa: E402
from bert_vits2_b200 import synth  # noqa: E402
from oracle.ref_import import build_reference_net  # noqa: E402

CASES = {
    # name: (use_transformer_flow, lengths, languages)
    "tflow_b1": (True, [24], [0]),
    "wnflow_b1": (False, [24], [0]),
    "tflow_b3": (True, [24, 17, 9], [0, 1, 2]),
    "wnflow_b3": (False, [21, 24, 12], [2, 0, 1]),
}
INFER_KW = dict(sdp_ratio=0.5, noise_scale=0.6, noise_scale_w=0.9, length_scale=1.0)
WEIGHT_SEED, INPUT_SEED, NOISE_SEED, F_CAP = 0, 1, 2, 512


class PatchedNoise:
    """Replace torch.randn / torch.randn_like during infer() with the seeded buffers."""

    def __init__(self, noise_w, noise_z):
        self.noise_w, self.noise_z, self.calls = noise_w, noise_z, []

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like
        torch.randn = self.randn
        torch.randn_like = self.randn_like
        return self

    def __exit__(self, *a):
        torch.randn, torch.randn_like = self._randn, self._randn_like

    def randn(self, *size, **kw):
        size = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        assert size == tuple(self.noise_w.shape), (size, self.noise_w.shape)
        self.calls.append(("randn", size))
        return self.noise_w.clone()

    def randn_like(self, t, **kw):
        self.calls.append(("randn_like", tuple(t.shape)))
        assert t.shape[:2] == self.noise_z.shape[:2]
        return self.noise_z[:, :, : t.shape[2]].clone()


def run_reference(net, inp, noise_w, noise_z, **kw):
    cap = {}
    hooks = [
        net.enc_p.register_forward_hook(lambda m, i, o: cap.__setitem__("enc_p", o)),
        net.sdp.register_forward_hook(lambda m, i, o: cap.__setitem__("logw_sdp", o)),
        net.dp.register_forward_hook(lambda m, i, o: cap.__setitem__("logw_dp", o)),
    ]
    with torch.no_grad(), PatchedNoise(noise_w, noise_z) as pn:
        o, attn, y_mask, (z, z_p, m_p, logs_p) = net.infer(
            inp["x"], inp["x_lengths"], inp["sid"], inp["tone"], inp["language"], inp["bert"], inp["ja_bert"],
            inp["en_bert"], **kw)
    for h in hooks:
        h.remove()
    assert [c[0] for c in pn.calls] == ["randn", "randn_like"], pn.calls
    x, m_tok, logs_tok, x_mask = cap["enc_p"]
    return dict(o=o, attn=attn, y_mask=y_mask, z=z, z_p=z_p, m_p=m_p, logs_p=logs_p, x=x, m_p_tok=m_tok,
                logs_p_tok=logs_tok, x_mask=x_mask, logw_sdp=cap["logw_sdp"], logw_dp=cap["logw_dp"],
                w_ceil=attn.sum(2))


def main():
    nets = {}
    for name, (tflow, lengths, langs) in CASES.items():
        if tflow not in nets:
            net, hps = build_reference_net(tflow)
            cfg = ModelConfig.from_hps_model(hps.model, use_transformer_flow=tflow)
            sd = synth.synthetic_state_dict(cfg, WEIGHT_SEED)
            missing, unexpected = net.load_state_dict(sd, strict=False)
            assert not unexpected and all(k.startswith("enc_q.") for k in missing), (missing[:5], unexpected[:5])
            nets[tflow] = (net, cfg)
        net, cfg = nets[tflow]
        inp = synth.synthetic_inputs(cfg, lengths, langs, seed=INPUT_SEED)
        nw, nz = synth.synthetic_noise(cfg, len(lengths), max(lengths), F_CAP, seed=NOISE_SEED)
        out = run_reference(net, inp, nw, nz, **INFER_KW)
        ylen = out["y_mask"].sum((1, 2)).long()
        print(name, "frames", ylen.tolist(), "o", tuple(out["o"].shape), "rms", float(out["o"].pow(2).mean().sqrt()))
        arrays = {k: v.detach().cpu().numpy() for k, v in out.items() if k != "attn"}
        arrays["y_lengths"] = ylen.numpy()
        meta = dict(use_transformer_flow=tflow, lengths=lengths, languages=langs, weight_seed=WEIGHT_SEED,
                    input_seed=INPUT_SEED, noise_seed=NOISE_SEED, f_cap=F_CAP, **INFER_KW)
        arrays["meta"] = np.array(repr(meta))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **arrays)


if __name__ == "__main__":
    main()
