"""Stream sharding across ranks: streams never interact, so each rank owns a
contiguous block of streams with all of their state; the only collective of the path is the
all-reduce(SUM) of the [C,4] int64 counters."""


def shard_range(n_streams, rank, world):
    """[lo, hi) of the streams owned by `rank` -- contiguous, sizes differ by at most one."""
    base, rem = divmod(n_streams, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
