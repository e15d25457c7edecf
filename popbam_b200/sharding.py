"""Region sharding for multi-GPU runs: the same plan the `popbam` command line uses (popbam_main.cpp, main()):
a shard is a run of whole windows whose span does not exceed `shard_bp`, shard s goes to device s % world, and rows are
gathered in window order.  Windows are independent, so no data crosses between ranks; reads that
overlap a shard boundary are delivered to both shards (the pileup of a window only uses positions inside the window).
Host logic only -- no compute."""


def plan_shards(win_beg, win_end, shard_bp):
    """[(w0, w1)] runs of windows [w0, w1) with win_end[w1-1] - win_beg[w0] <= shard_bp (at least one window each)."""
    shards, w, nw = [], 0, len(win_beg)
    while w < nw:
        e = w + 1
        while e < nw and int(win_end[e]) - int(win_beg[w]) <= shard_bp:
            e += 1
        shards.append((w, e))
        w = e
    return shards


def shards_of_rank(n_shards, rank, world):
    return list(range(rank, n_shards, world))


def gather_rows(local_rows, rank, world, group=None):
    """local_rows: {shard index: text}.  Returns the rows of all shards in shard (== window) order on rank 0, else None.
    Uses torch.distributed object collectives (gloo or nccl); with world == 1 it is a plain sort."""
    if world == 1:
        return "".join(local_rows[s] for s in sorted(local_rows))
    import torch.distributed as dist
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local_rows, gathered, dst=0, group=group)
    if rank != 0:
        return None
    merged = {}
    for part in gathered:
        merged.update(part)
    return "".join(merged[s] for s in sorted(merged))


def max_over_ranks(value, world, device="cpu"):
    """Slowest rank's time: what a multi-GPU throughput number must be divided by."""
    if world == 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
