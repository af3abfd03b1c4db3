"""CPU, world_size 2 over gloo: the host-side multi-rank logic of bench.py (per-rank synthetic shards, max-over-ranks
timing, whole-job aggregate) — the data path itself has no collective in inference (replicas)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    ms_local = 10.0 + 5.0 * rank                       # rank 1 is slower
    ms = bench.max_over_ranks(ms_local, torch.device("cpu"), world)
    value = bench.aggregate_throughput(units_per_rank_step=64, steps=20, world=world, ms_total=ms)
    x = bench.synthetic_images(2, 1234 + rank * 16)
    sig = torch.tensor([float(x.sum())])
    sigs = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(sigs, sig)
    q.put((rank, ms, value, [float(s) for s in sigs]))
    dist.destroy_process_group()


def test_bench_aggregation_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, value, sigs in res:
        assert ms == 15.0                                # max over ranks, identical on every rank
        assert abs(value - 2 * 64 * 20 / 15e-3) < 1e-6   # whole-job units / max time
        assert sigs[0] != sigs[1]                        # ranks draw different synthetic shards
