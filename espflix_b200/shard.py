"""espflix_b200/shard.py — multi-GPU plumbing: independent streams shard as one contiguous block per
rank, no data-path collective; a single all_gather of per-rank frame counts for the report
(SURVEY.md 8e). One process per GPU, torch.distributed (nccl on GPUs, gloo in the CPU tests)."""
import os


def env_rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def partition(n_streams, world):
    """Contiguous block of streams per rank: [(first, count)] * world, sizes differ by at most one."""
    base, extra = divmod(n_streams, world)
    out, first = [], 0
    for r in range(world):
        c = base + (1 if r < extra else 0)
        out.append((first, c))
        first += c
    return out


def stream_seed_index(rank, local_index, distinct):
    """Which distinct synthetic stream a (rank, local stream) pair uses: ranks take different windows
    of the seed space so that the whole job is not world copies of the same data."""
    return rank * distinct + (local_index % distinct)


def init(backend=None, device=None):
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world == 1:
        return None
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    kw = {}
    if device is not None:
        kw["device_id"] = device
    dist.init_process_group(backend or "nccl", rank=rank, world_size=world, **kw)
    return dist


def gather_counts(dist, value, device="cpu"):
    """all_gather of one int64 per rank (reporting only) -> list of ints, same on every rank."""
    import torch
    if dist is None:
        return [int(value)]
    world = dist.get_world_size()
    mine = torch.tensor([int(value)], dtype=torch.int64, device=device)
    outs = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(outs, mine)
    return [int(t.item()) for t in outs]


def max_over_ranks(dist, values, device="cpu"):
    """element-wise MAX of a list of floats over ranks (device-timed durations)."""
    import torch
    if dist is None:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]
