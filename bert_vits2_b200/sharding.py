"""Multi-GPU plumbing for the infer path: utterances are independent (no cross-batch op anywhere in reference
models.py:1026-1074), so the path shards embarrassingly -- one process per GPU, full weight replica per rank, no
data-path collective.  The single exchange step is the gather of the finished waveforms; the
reference has no equivalent (inference is single-device, webui.py:31, 397-399).

`deal_buckets` reuses the length-bucketing idea of the reference's training sampler (data_utils.py:305-335):
sort by token count, cut into per-rank batches of similar length so padding (and the padded-tail work the
reference semantics require) is minimal, and deal batches to ranks so that every rank gets the
same number of batches and a balanced token count.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def deal_buckets(lengths: Sequence[int], world_size: int, batch_size: int) -> List[List[List[int]]]:
    """Returns plan[rank] = list of batches, each a list of utterance indices (sorted by length inside a batch).
    Every utterance appears exactly once; ranks differ by at most one batch."""
    order = sorted(range(len(lengths)), key=lambda i: (lengths[i], i))
    batches = [order[i:i + batch_size] for i in range(0, len(order), batch_size)]
    # longest batches first, greedy onto the least-loaded rank (load = padded tokens), ties -> lowest rank
    batches.sort(key=lambda b: -max(lengths[i] for i in b) * len(b))
    plan: List[List[List[int]]] = [[] for _ in range(world_size)]
    load = [0] * world_size
    cap = -(-len(batches) // world_size)
    for b in batches:
        cands = [r for r in range(world_size) if len(plan[r]) < cap]
        r = min(cands, key=lambda q: (load[q], q))
        plan[r].append(b)
        load[r] += max(lengths[i] for i in b) * len(b)
    return plan


def gather_waveforms(wave: torch.Tensor, n_samples: torch.Tensor, dst: int = 0) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """Gather per-rank padded waveform batches [B_r, 1, L_r] (+ valid sample counts [B_r]) to rank `dst`.
    Works on any backend (nccl over NVLink on the GPU box, gloo in CPU tests).  Shapes may differ per rank, so the
    sizes are exchanged first and the payload is padded to the maximum."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return [wave], [n_samples]
    rank = dist.get_rank()
    shape = torch.tensor([wave.shape[0], wave.shape[-1]], dtype=torch.int64, device=wave.device)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape)
    Bm = int(max(s[0] for s in shapes)); Lm = int(max(s[1] for s in shapes))
    pad = torch.zeros(Bm, 1, Lm, dtype=wave.dtype, device=wave.device)
    pad[: wave.shape[0], :, : wave.shape[-1]] = wave
    ns = torch.zeros(Bm, dtype=torch.int64, device=wave.device)
    ns[: n_samples.shape[0]] = n_samples.to(torch.int64)
    outs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    nss = [torch.empty_like(ns) for _ in range(world)] if rank == dst else None
    dist.gather(pad, outs, dst=dst)
    dist.gather(ns, nss, dst=dst)
    if rank != dst:
        return [], []
    waves = [o[: int(s[0]), :, : int(s[1])] for o, s in zip(outs, shapes)]
    counts = [n[: int(s[0])] for n, s in zip(nss, shapes)]
    return waves, counts
