"""Batch sharding over GPUs: independent ciphertexts partition across ranks in contiguous blocks of ceil(B/G)
(SURVEY.md 8e); there is no data-path collective, torch.distributed only carries the barrier and the reduction of
the per-rank device time (max over ranks)."""


def shard_range(total, rank, world):
    """[begin, end) of the ciphertext indices rank `rank` owns."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank/world is invalid")
    per = -(-total // world)
    begin = min(total, rank * per)
    return begin, min(total, begin + per)


def max_over_ranks(value, dist=None, device=None):
    """max of a python float over all ranks (device time of the slowest rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
