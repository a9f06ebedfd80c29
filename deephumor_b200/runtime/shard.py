"""Data-parallel sharding of a caption job over the GPUs of one box.

The reference has no distributed code; images are independent units (eval-mode BN, per-image beam state), so
the job is cut into contiguous ranges of the GLOBAL image index, every rank runs encoder + decode locally on
replicated weights, and the only collective on the path is ONE all-gather of the `[n_local, max_len]` int64 ids
(+ `[n_local]` lengths) at the end -- NCCL over NVLink on the GPU box, gloo in the CPU tests.  Synthetic inputs
and the injected noise are keyed on the global index (`image_base`), so the gathered result is independent of
the world size.
"""
import torch


def shard_range(total, rank, world):
    """Contiguous range of global image indices owned by `rank`: (first, count).  The first `total % world`
    ranks take one extra image; counts differ by at most one."""
    if world < 1 or not 0 <= rank < world or total < 0:
        raise ValueError(f'bad shard request total={total} rank={rank} world={world}')
    base, extra = divmod(total, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def gather_captions(ids, lengths, total=None, group=None):
    """All-gather of per-rank results in global image order.

    ids int64 [n_local, max_len], lengths int64 [n_local] (on the backend's device) -> (ids [total, max_len],
    lengths [total]) on every rank.  Ragged shards (shard_range with total % world != 0) are padded to the
    largest shard for the collective and trimmed afterwards.  With no initialised process group (or world 1)
    the inputs are returned unchanged."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return ids, lengths
    world = dist.get_world_size(group)
    n_local, max_len = ids.shape
    if total is None:
        counts = [n_local] * world
    else:
        counts = [shard_range(total, r, world)[1] for r in range(world)]
        if counts[dist.get_rank(group)] != n_local:
            raise ValueError(f'rank holds {n_local} rows, shard_range says {counts[dist.get_rank(group)]}')
    width = max(counts)
    # one buffer per rank: ids and the length in an extra column, so the path has exactly ONE collective
    packed = torch.zeros(width, max_len + 1, dtype=torch.int64, device=ids.device)
    packed[:n_local, :max_len] = ids
    packed[:n_local, max_len] = lengths
    out = torch.empty(world * width, max_len + 1, dtype=torch.int64, device=ids.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    out = out.view(world, width, max_len + 1)
    parts = [out[r, :counts[r]] for r in range(world)]
    full = torch.cat(parts, dim=0)
    return full[:, :max_len].contiguous(), full[:, max_len].contiguous()


def generate_sharded(model, total, make_inputs, group=None, **gen_kwargs):
    """Runs `model.generate` on this rank's shard of a `total`-image job and gathers all ids.

    make_inputs(first, count) -> tuple of positional inputs for model.generate (images[, labels]) for global
    indices first .. first+count-1.  Returns (ids [total, max_len], lengths [total]) on every rank."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    first, count = shard_range(total, rank, world)
    out = model.generate(*make_inputs(first, count), image_base=first, **gen_kwargs)
    if isinstance(out, tuple):
        ids, lengths = out
    else:                                   # batch-1 reference-shaped return
        max_len = gen_kwargs.get('max_len', 25)
        ids = torch.zeros(1, max_len, dtype=torch.int64, device=out.device)
        ids[0, :out.numel()] = out
        lengths = torch.tensor([out.numel()], dtype=torch.int64, device=out.device)
    return gather_captions(ids, lengths, total=total, group=group)
