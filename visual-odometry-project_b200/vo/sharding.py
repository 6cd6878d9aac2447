"""How independent sequences are spread over ranks (one process per GPU, no data-path collective):
a VO sequence never shards within a frame, so rank r simply owns a contiguous block of sequences."""


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced block [lo, hi) of n_items for `rank` out of `world` ranks."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("shard_range: need 0 <= rank < world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def job_frames_per_second(frames_per_rank, seconds_per_rank):
    """Whole-job throughput: all ranks' frames divided by the slowest rank's time."""
    return sum(frames_per_rank) / max(seconds_per_rank)
