"""Shared helpers for the parity tests (golden loading, synthetic model/case regeneration)."""
import ast
import functools
import os

import numpy as np
import torch

from bert_vits2_b200 import synth
from bert_vits2_b200.spec import ModelConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["tflow_b1", "wnflow_b1", "tflow_b3", "wnflow_b3"]


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = ast.literal_eval(str(d["meta"]))
    arrays = {k: torch.from_numpy(d[k]) for k in d.files if k != "meta"}
    return meta, arrays


@functools.lru_cache(maxsize=4)
def model_for(use_transformer_flow: bool, weight_seed: int = 0):
    cfg = ModelConfig(use_transformer_flow=use_transformer_flow)
    sd = synth.synthetic_state_dict(cfg, weight_seed)
    return cfg, sd


def case_inputs(meta):
    cfg, sd = model_for(meta["use_transformer_flow"], meta["weight_seed"])
    inp = synth.synthetic_inputs(cfg, meta["lengths"], meta["languages"], seed=meta["input_seed"])
    nw, nz = synth.synthetic_noise(cfg, len(meta["lengths"]), max(meta["lengths"]), meta["f_cap"], seed=meta["noise_seed"])
    kw = {k: meta[k] for k in ("sdp_ratio", "noise_scale", "noise_scale_w", "length_scale")}
    return cfg, sd, inp, nw, nz, kw


def rms(a, b):
    return float((a.double() - b.double()).pow(2).mean().sqrt())
