"""N>1 path on CPU: world_size-2 gloo test of the utterance sharding + waveform gather (the only exchange step)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bert_vits2_b200.sharding import deal_buckets, gather_waveforms


def test_deal_buckets_partition_and_balance():
    import random
    rnd = random.Random(0)
    lengths = [rnd.randint(64, 512) for _ in range(256)]  # BASELINE.json config 4
    plan = deal_buckets(lengths, world_size=8, batch_size=32)
    seen = sorted(i for r in plan for b in r for i in b)
    assert seen == list(range(256))
    assert all(len(r) == 1 for r in plan)
    for r in plan:
        for b in r:
            ls = [lengths[i] for i in b]
            assert ls == sorted(ls)
    # length-bucketed: padding waste below 15 %
    waste = sum(max(lengths[i] for i in b) * len(b) - sum(lengths[i] for i in b) for r in plan for b in r)
    assert waste / sum(lengths) < 0.15
    plan3 = deal_buckets(lengths[:70], world_size=3, batch_size=8)
    sizes = [len(r) for r in plan3]
    assert max(sizes) - min(sizes) <= 1 and sorted(i for r in plan3 for b in r for i in b) == list(range(70))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, L = 2 + rank, 512 * (3 + 2 * rank)
    wave = torch.full((B, 1, L), float(rank + 1))
    wave[:, 0, 0] = torch.arange(B, dtype=torch.float32)
    ns = torch.tensor([L - 512 * i for i in range(B)])
    waves, counts = gather_waveforms(wave, ns, dst=0)
    if rank == 0:
        ok = len(waves) == world
        for r in range(world):
            ok &= tuple(waves[r].shape) == (2 + r, 1, 512 * (3 + 2 * r))
            ok &= bool((waves[r][:, 0, 1:] == r + 1).all()) and waves[r][:, 0, 0].tolist() == list(range(2 + r))
            ok &= counts[r].tolist() == [512 * (3 + 2 * r) - 512 * i for i in range(2 + r)]
        q.put(bool(ok))
    else:
        q.put(waves == [] and counts == [])
    dist.barrier()
    dist.destroy_process_group()


def test_gather_waveforms_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(res) and all(p.exitcode == 0 for p in procs)
