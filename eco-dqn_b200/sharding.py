"""Episode sharding across the GPUs of one box.

Rollout episodes are independent, so rank r simply owns a contiguous block of episodes and of the graphs they use;
nothing crosses NVLink while stepping.  The only exchange is the final gather of per-episode best cuts (and, for
training, the gradient all-reduce in agents/dqn).  One process per GPU, `torch.distributed` (NCCL on GPUs, gloo in
the CPU tests) does the plumbing.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous block of `n_items` owned by `rank`: sizes differ by at most one, earlier ranks get the extras."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %d/%d" % (rank, world))
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_best(local, n_total=None):
    """All-gather the per-episode results of every rank into episode order.  `local` is this rank's 1-D (or [n, ...])
    tensor for its shard_range block; shards may differ in length by one (padded for the collective, trimmed after)."""
    rank, world = world_info()
    if world == 1:
        return local.clone()
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    width = max(sizes)
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    out = torch.cat([p[:s] for p, s in zip(parts, sizes)])
    if n_total is not None and out.shape[0] != n_total:
        raise RuntimeError("gathered %d results, expected %d" % (out.shape[0], n_total))
    return out


def allreduce_mean_(flat):
    """In-place mean over ranks of one flat gradient buffer: the single collective of a data-parallel DQN update."""
    rank, world = world_info()
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
    return flat


def allreduce_mean_grads(params):
    """Average the gradients of `params` over ranks with ONE collective: flatten (58 425 floats for the MPNN), all-reduce,
    scatter back in place.  Returns the flat averaged buffer."""
    params = [p for p in params if p.grad is not None]
    # the gradient kernels (eco_mpnn_grad) leave the gradients as consecutive views of one buffer: reduce it in place
    base, off, aliased = params[0].grad, 0, True
    for p in params:
        g = p.grad
        aliased = aliased and g.is_contiguous() and g.untyped_storage().data_ptr() == base.untyped_storage().data_ptr() \
            and g.storage_offset() == base.storage_offset() + off
        off += g.numel()
    if aliased:
        return allreduce_mean_(torch.as_strided(base, (off,), (1,), base.storage_offset()))
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    allreduce_mean_(flat)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat


def best_per_graph(best_cut, graph_idx, n_graphs):
    """max over the episodes of each graph (the `cut` column of test_network), on whatever device the inputs live."""
    out = torch.full((n_graphs,), torch.iinfo(torch.int32).min, dtype=best_cut.dtype, device=best_cut.device)
    return out.scatter_reduce(0, graph_idx.long(), best_cut, reduce="amax", include_self=True)
