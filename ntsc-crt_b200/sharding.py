"""Multi-GPU host logic: frames (monitors) shard across ranks, no data-path collective.

crt_modulate / crt_demodulate of different `struct CRT` instances never exchange data, so the
N-GPU path is one process per GPU (torchrun), each advancing its own contiguous slice of the batch;
torch.distributed is used only for the barrier, the max-over-ranks timing and -- optionally -- to
all_gather the decoded frames when a caller wants every rank to hold the whole batch (the exchange
BASELINE.json's north_star mentions; it is NVLink-bound).
"""
import os


def rank_info():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) standalone."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) of rank's items; sizes differ by at most one, earlier ranks larger."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(values, device=None, group=None):
    """Element-wise max of a list of floats over all ranks (timings are max-over-ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(x) for x in t]


def allgather_frames(local, group=None):
    """all_gather equally-shaped per-rank frame tensors -> (world * n_local, ...) on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out.view((world * local.shape[0],) + tuple(local.shape[1:]))
