"""CPU tier: the N>1 path.  Frames shard across ranks with no data-path collective; only the six
error counters are summed.  world_size-2 gloo, the oracle standing in for the GPU."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import oracle_lib as ol  # noqa: E402
import sc_polar_decoder_hls_b200 as scpd  # noqa: E402


def test_shard_ranges_cover_the_stream_exactly_once():
    for total in (0, 1, 7, 8, 1000, 2 ** 20 + 3):
        for world in (1, 2, 3, 8):
            spans = [bench.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, k = 1024, 512
    flags = scpd.packed_flags("FB_N1024_K512", n)
    lo, hi = bench.shard_range(total, rank, world)
    # each rank starts the channel at its own stream position (xorshift jump-ahead on the device)
    llr = ol.channel(n, hi - lo, ol.sigma(2.5, 0.5), first_frame=lo)
    cnt = ol.count_errors(n, ol.decode(n, 16, 8, 0, 1, flags, llr))
    t = torch.tensor(cnt, dtype=torch.int64)
    summed = bench.sum_counters(t)
    tmax = bench.max_over_ranks(float(rank + 1))
    if rank == 0:
        q.put((summed.tolist(), tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_counters_equal_single_process():
    total, world = 301, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    summed, tmax = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    n = 1024
    flags = scpd.packed_flags("FB_N1024_K512", n)
    llr = ol.channel(n, total, ol.sigma(2.5, 0.5))
    want = ol.count_errors(n, ol.decode(n, 16, 8, 0, 1, flags, llr))
    assert summed == want
    assert tmax == 2.0
