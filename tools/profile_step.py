#!/usr/bin/env python
"""Run N device-resident infer() steps of the bench workload (config 2) for ncu / launch-list captures.
Never use a number printed under a profiler as a bench value."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bert_vits2_b200 import synth  # noqa: E402
from bert_vits2_b200.engine import Engine  # noqa: E402
from bert_vits2_b200.spec import ModelConfig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--T", type=int, default=256)
ap.add_argument("--generator-only", type=int, default=0, help="frames; run bv2_generator only (config 5)")
a = ap.parse_args()
cfg = ModelConfig()
sd = synth.synthetic_state_dict(cfg, 0)
eng = Engine(cfg, sd, "cuda:0", a.precision)
if a.generator_only:
    z, g = synth.synthetic_generator_inputs(cfg, 1, a.generator_only)
    z, g = z.cuda(), g.cuda()
    for _ in range(a.steps):
        o = eng.generator(z, g)
    torch.cuda.synchronize()
    print("generator-only", o.shape, eng.launch_count)
else:
    inp = synth.synthetic_inputs(cfg, [a.T], [0], seed=2)
    nw, nz = synth.synthetic_noise(cfg, 1, a.T, 4096, seed=2)
    d = {k: v.cuda() for k, v in inp.items()}
    nw, nz = nw.cuda(), nz.cuda()
    kw = bench.INFER_KW
    for i in range(a.steps):
        l0 = eng.launch_count
        ylen, F = eng.infer_begin(d["x"], d["x_lengths"], d["sid"], d["tone"], d["language"], d["bert"], d["ja_bert"], d["en_bert"], nw,
                                  kw["noise_scale_w"], kw["length_scale"], kw["sdp_ratio"])
        o, *_ = eng.infer_finish(1, a.T, F, nz, kw["noise_scale"])
        torch.cuda.synchronize()
        print("step", i, "frames", F, "launches", eng.launch_count - l0)
