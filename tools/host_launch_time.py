#!/usr/bin/env python
"""Host-side cost of one infer() step: wall time the C-ABI calls take to ENQUEUE their launches vs the GPU time of the stages.
If the enqueue time of bv2_infer_finish approaches flow + Generator time, the step is launch-bound (a CUDA graph would help)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bert_vits2_b200 import synth
from bert_vits2_b200.engine import Engine
from bert_vits2_b200.spec import ModelConfig
cfg = ModelConfig(); sd = synth.synthetic_state_dict(cfg, 0)
eng = Engine(cfg, sd, "cuda:0", "fp16"); eng.set_profiling(True)
inp, nw, nz = bench.make_case(cfg)
d = {k: v.cuda() for k, v in inp.items()}; nw, nz = nw.cuda(), nz.cuda()
kw = bench.INFER_KW; B, T = inp["x"].shape
eng.reserve(B, T, 2048)
tb = tf = ts = 0.0; n = 20
for i in range(5 + n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ylen, F = eng.infer_begin(d["x"], d["x_lengths"], d["sid"], d["tone"], d["language"], d["bert"], d["ja_bert"], d["en_bert"], nw, kw["noise_scale_w"], kw["length_scale"], kw["sdp_ratio"])
    t1 = time.perf_counter()
    o, *_ = eng.infer_finish(B, T, F, nz, kw["noise_scale"], want_attn=False)
    t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
    if i >= 5: tb += t1 - t0; tf += t2 - t1; ts += t3 - t2
print(f"host ms: infer_begin (incl. its read-back sync) {tb / n * 1e3:.3f} | infer_finish enqueue {tf / n * 1e3:.3f} | wait after enqueue {ts / n * 1e3:.3f} | "
      f"GPU stage ms: enc {eng.stage_ms('encoder_duration'):.3f} flow {eng.stage_ms('flow'):.3f} gen {eng.stage_ms('generator'):.3f}; cpus {os.cpu_count()}")
