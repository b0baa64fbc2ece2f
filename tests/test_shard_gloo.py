"""N>1 host logic on CPU: world_size-2 gloo processes — stream partition, rank-specific seeds, the
reporting all_gather of per-rank frame counts and the max-over-ranks timing reduction."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from espflix_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from espflix_b200 import shard as sh, synth
    dist = sh.init(backend="gloo")
    first, count = sh.partition(10, world)[rank]
    # each rank "decodes" its own block: here the oracle-free host side only counts pictures
    es, off = synth.generate(synth.SEED0 + sh.stream_seed_index(rank, 0, 4), n_pictures=2, gop=2, noise=0)
    frames_done = count * (len(off) - 1)
    counts = sh.gather_counts(dist, frames_done)
    tmax = sh.max_over_ranks(dist, [1.0 + rank, 5.0 - rank])
    q.put((rank, first, count, counts, tmax, int(es.size)))
    dist.barrier()
    dist.destroy_process_group()


def test_partition():
    assert shard.partition(32768, 8) == [(i * 4096, 4096) for i in range(8)]
    p = shard.partition(10, 3)
    assert p == [(0, 4), (4, 3), (7, 3)] and sum(c for _, c in p) == 10
    assert shard.partition(1, 4) == [(0, 1), (1, 0), (1, 0), (1, 0)]


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, c0, counts0, t0, s0), (r1, f1, c1, counts1, t1, s1) = res
    assert (f0, c0, f1, c1) == (0, 5, 5, 5)
    assert counts0 == counts1 == [10, 10]                 # 5 streams x 2 pictures per rank; sum == input count
    assert t0 == t1 == [2.0, 5.0]
    assert s0 != s1                                       # rank-specific seeds give different streams
