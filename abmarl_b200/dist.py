"""Multi-GPU plumbing: env batches shard across ranks with no data-path collective (envs never interact, SURVEY.md
8(e)); the only collective is a sum of the small episode-statistics vector (NCCL on GPUs, gloo in the CPU tests).

Philox is keyed by the GLOBAL env index (BgwSpec.env_offset), so results do not depend on the number of ranks."""
import torch
import torch.distributed as dist


def shard_envs(total_envs, world_size, rank):
    """Contiguous env range [offset, offset + n) of `rank`; the first `total % world` ranks take one more."""
    base, extra = divmod(int(total_envs), int(world_size))
    n = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, n


def reduce_stats(stats, op=None):
    """Sum (default) a small per-rank vector over all ranks; no-op without an initialised process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=op or dist.ReduceOp.SUM)
    return stats


def global_stats(engine):
    """(agent_steps, episodes, kills, env_steps) summed over every env of every rank, as an int64 tensor."""
    return reduce_stats(engine.stats().clone())
