"""Multi-GPU rollout: shard start states over ranks, all-gather the synthetic transitions (NCCL).

The reference is single-process / single-GPU; this is the section 8(e) design:
rows are independent given replicated weights, so each rank rolls a contiguous range of global rows
with Philox draws keyed on the GLOBAL row id (results do not depend on the number of GPUs), and one
exchange at the end assembles the transitions: a count all-gather plus an all-gather of padded
per-rank slabs (an exact all-gather-v), concatenated rank-major.  No collective touches the data path
of the rollout itself.  One process per GPU (torchrun); torch.distributed is plumbing only.
"""
import torch
import torch.distributed as dist

KEYS = ("obss", "actions", "next_obss", "rewards", "terminals", "penalty")   # column order of a packed slab


def shard_range(n, rank, world):
    """Contiguous global-row range [lo, hi) of ``rank``: sizes differ by at most one, ranges tile [0, n)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_transitions(out):
    """dict of [M, w_k] tensors -> one [M, sum w_k] tensor (column order KEYS) and the widths."""
    widths = [out[k].shape[1] for k in KEYS]
    return torch.cat([out[k] for k in KEYS], dim=1), widths


def unpack_transitions(packed, widths):
    res, c = {}, 0
    for k, w in zip(KEYS, widths):
        res[k] = packed[:, c:c + w]
        c += w
    return res


def allgather_transitions(packed, capacity, group=None):
    """Exact all-gather-v of per-rank [M_r, W] row blocks (M_r <= capacity) -> [sum M_r, W], rank-major.

    Two collectives: the counts (world ints) and the zero-padded [capacity, W] slabs.  Works with the
    NCCL backend on CUDA tensors and with gloo on CPU tensors (the CPU tests use the latter)."""
    world = dist.get_world_size(group)
    m, w = packed.shape
    if m > capacity:
        raise ValueError(f"rank holds {m} rows, more than the slab capacity {capacity}")
    cnt = torch.tensor([m], dtype=torch.int64, device=packed.device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    slab = torch.zeros(capacity, w, dtype=packed.dtype, device=packed.device)
    slab[:m] = packed
    slabs = [torch.empty_like(slab) for _ in range(world)]
    dist.all_gather(slabs, slab, group=group)
    counts = [int(c.item()) for c in counts]
    return torch.cat([s[:c] for s, c in zip(slabs, counts)], dim=0), counts


def allgather_slabs(slab, group=None, async_op=False):
    """One collective for an exact all-gather-v: ``slab`` is [capacity + 1, W] with this rank's row count stored
    in-band at slab[capacity, 0].  Returns ([world, capacity + 1, W], counts as a device float tensor[world]);
    no host synchronisation."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(slab.shape), dtype=slab.dtype, device=slab.device)
    if slab.is_cuda:
        work = dist.all_gather_into_tensor(out, slab, group=group, async_op=async_op)
        if async_op:                        # NCCL runs on its own stream; work.wait() orders the caller's stream after it
            return out, out[:, -1, 0], work
    else:                                   # gloo (CPU tests)
        parts = [torch.empty_like(slab) for _ in range(world)]
        dist.all_gather(parts, slab, group=group)
        out = torch.stack(parts, 0)
    return out, out[:, -1, 0]


def sharded_rollout(agent, init_obss, rollout_length, use_trg=True, *, group=None, gather=True, sharded_input=False):
    """MOBODY.rollout over all ranks of ``group``.

    init_obss: the GLOBAL start states [B, S] (every rank passes the same tensor and takes its shard), or with
    ``sharded_input=True`` this rank's own shard (ranks must hold equal-sized shards; weak-scaling benches).
    gather: False -> this rank's transitions only; "padded" -> ([world, cap+1, W] slabs, device counts, widths)
    with no host synchronisation in the exchange; "padded_async" -> same plus the NCCL work handle, so the
    all-gather of step t overlaps the rollout of step t+1 (call .wait() before reading the slabs);
    True -> compact dict of the transitions of all ranks (rank-major).
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    T = int(rollout_length)
    if sharded_input:
        local, lo, cap = init_obss, rank * init_obss.shape[0], init_obss.shape[0] * T
    else:
        lo, hi = shard_range(init_obss.shape[0], rank, world)
        local = init_obss[lo:hi]
        cap = shard_range(init_obss.shape[0], 0, world)[1] * T
    if not gather or T == 0:
        return agent.rollout_device(local, T, use_trg, row0=lo)
    S = local.shape[1]
    probe_w = getattr(agent, "config", {}).get("action_dim")
    W = 2 * S + probe_w + 3 if probe_w is not None else None
    if W is None or not local.is_cuda:       # generic path (CPU tests with a stand-in agent)
        out, info = agent.rollout_device(local, T, use_trg, row0=lo)
        packed, widths = pack_transitions(out)
        allp, counts = allgather_transitions(packed, cap, group)
        stats = torch.tensor([info["num_transitions"], info["reward_mean"] * info["num_transitions"]], dtype=torch.float64,
                             device=packed.device)
        dist.all_reduce(stats, group=group)
        n = int(stats[0].item())
        return unpack_transitions(allp, widths), {"num_transitions": n, "reward_mean": float(stats[1].item()) / max(n, 1),
                                                  "kept": int(sum(counts)), "kept_per_rank": counts}
    slab = torch.empty(cap + 1, W, dtype=torch.float32, device=local.device)
    out, info = agent.rollout_device(local, T, use_trg, row0=lo, out_packed=slab[:cap], sync=False)   # nothing read back
    # in-band header row: kept rows (exact below 2^24), produced transitions, reward sum — written on the stream
    slab[cap, 0:1] = info["kept_dev"].float()
    slab[cap, 1:3] = info["stats_dev"].flip(0).float()
    widths = [S, probe_w, S, 1, 1, 1]
    if gather == "padded_async":
        slabs, counts_dev, work = allgather_slabs(slab, group, async_op=True)
        return (slabs, counts_dev, widths, work), dict(info, world=world, capacity=cap)
    slabs, counts_dev = allgather_slabs(slab, group)
    if gather == "padded":
        return (slabs, counts_dev, widths), dict(info, world=world, capacity=cap)
    hdr = slabs[:, cap, :3].double().cpu()                             # the one host read of the exchange
    counts = [int(c) for c in hdr[:, 0]]
    allp = torch.cat([slabs[r, :counts[r]] for r in range(world)], dim=0)
    n = int(hdr[:, 1].sum())
    return unpack_transitions(allp, widths), {"num_transitions": n, "reward_mean": float(hdr[:, 2].sum()) / max(n, 1),
                                              "kept": int(sum(counts)), "kept_per_rank": counts}
