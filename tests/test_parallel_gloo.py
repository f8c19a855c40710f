"""Host-side logic of the multi-GPU rollout on CPU: 2 ranks over gloo (127.0.0.1)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_tile_the_rows():
    from mobody_b200.parallel import shard_range
    for n in (0, 1, 7, 100000, 1000003):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mobody_b200 import parallel as P
    S, A, T, B = 5, 2, 3, 11
    lo, hi = P.shard_range(B, rank, world)

    class FakeAgent:      # stands in for MOBODY.rollout_device: row i of step t is tagged with its global id
        def rollout_device(self, init, T, use_trg=True, row0=0):
            n = init.shape[0]
            ids = torch.arange(row0, row0 + n, dtype=torch.float32)
            keep = [ids[ids % (t + 2) != 0] for t in range(T)]            # ragged: different survivors per step
            m = torch.cat(keep)
            out = {"obss": m[:, None].repeat(1, S), "actions": m[:, None].repeat(1, A), "next_obss": m[:, None].repeat(1, S) + 0.5,
                   "rewards": m[:, None], "terminals": torch.zeros(len(m), 1), "penalty": torch.ones(len(m), 1)}
            return out, {"num_transitions": n * T, "reward_mean": float(m.mean()) if len(m) else 0.0}
    init = torch.arange(B, dtype=torch.float32)[:, None].repeat(1, S)
    out, info = P.sharded_rollout(FakeAgent(), init, T)
    # reference: concatenate the shards in rank order
    want = []
    for r in range(world):
        l, h = P.shard_range(B, r, world)
        ids = torch.arange(l, h, dtype=torch.float32)
        want.append(torch.cat([ids[ids % (t + 2) != 0] for t in range(T)]))
    want = torch.cat(want)
    ok = torch.equal(out["rewards"][:, 0], want) and out["obss"].shape == (len(want), S) and out["actions"].shape == (len(want), A)
    ok = ok and info["num_transitions"] == B * T and info["kept"] == len(want) and len(info["kept_per_rank"]) == world
    ok = ok and torch.equal(out["next_obss"][:, 0], want + 0.5)
    # exact all-gather-v incl. an empty rank, and the capacity check
    blk = torch.full((rank * 3, 4), float(rank))
    allp, counts = P.allgather_transitions(blk, 8)
    ok = ok and counts == [r * 3 for r in range(world)] and allp.shape == (sum(counts), 4)
    try:
        P.allgather_transitions(torch.zeros(9, 4), 8)
        ok = False
    except ValueError:
        pass
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_rollout_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
