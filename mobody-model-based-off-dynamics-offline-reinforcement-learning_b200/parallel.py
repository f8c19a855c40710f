"""Multi-GPU rollout: shard start states over ranks, all-gather the synthetic transitions (NCCL).

The reference is single-process / single-GPU; this is the section 8(e) design:
rows are independent given replicated weights, so each rank rolls a contiguous range of global rows
with Philox draws keyed on the GLOBAL row id (results do not depend on the number of GPUs), and one
exchange at the end assembles the transitions: a count all-gather plus an all-gather of padded
per-rank slabs (an exact all-gather-v), concatenated rank-major.  No collective touches the data path
of the rollout itself.  One process per GPU (torchrun); torch.distributed is plumbing only.

Two assemblies of the per-rank results are offered:
 * ``gather="p2p"`` (default on CUDA): each rank's rollout packs its kept transitions into slot ``rank`` of its own receive
   buffer, and a narrow kernel on a high-priority side stream streams exactly those rows into the same slot of every peer's
   buffer through peer-mapped pointers (csrc/peer.cu: 128-bit NVLink stores, or one NVSwitch-replicated multicast store,
   then header + flag).  No collective kernel, no padding on the wire, the consumer waits on the device.
 * ``gather="padded"`` / ``"padded_async"``: one NCCL all-gather of zero-padded slabs (round 1; kept as the fallback when
   peer memory cannot be mapped, and for the CPU / gloo tests).
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

KEYS = ("obss", "actions", "next_obss", "rewards", "terminals", "penalty")   # column order of a packed slab


class _StreamMemOps:
    """cuStreamWaitValue32 / cuStreamWriteValue32 (cuda-python): flags awaited and raised IN STREAM ORDER by the front end,
    with no kernel -- a fused step kernel fills every SM completely (228 KB of shared memory), so even a one-warp flag kernel
    has to wait for a tile to finish and then pushes that SM's next tile back."""

    def __init__(self):
        from cuda.bindings import driver as drv
        self.drv = drv
        self.geq = drv.CUstreamWaitValue_flags.CU_STREAM_WAIT_VALUE_GEQ
        self.wdef = drv.CUstreamWriteValue_flags.CU_STREAM_WRITE_VALUE_DEFAULT

    def _ok(self, res, what):
        code = res[0] if isinstance(res, tuple) else res
        if int(code) != 0:
            raise RuntimeError(f"mobody_b200: {what} failed with CUresult {int(code)}")

    def wait_geq(self, stream, addr, value):
        self._ok(self.drv.cuStreamWaitValue32(self.drv.CUstream(int(stream.cuda_stream)), self.drv.CUdeviceptr(int(addr)), int(value), self.geq),
                 "cuStreamWaitValue32")

    def write(self, stream, addr, value):
        self._ok(self.drv.cuStreamWriteValue32(self.drv.CUstream(int(stream.cuda_stream)), self.drv.CUdeviceptr(int(addr)), int(value), self.wdef),
                 "cuStreamWriteValue32")


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
    """One collective for an exact all-gather-v: ``slab`` is [capacity + 1, W] with this rank's header stored in-band in
    the last row as INTEGER words (int32 kept rows, int32 produced transitions, float64 reward sum).  Returns
    ([world, capacity + 1, W], kept rows as a device int32 tensor[world]); no host synchronisation."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(slab.shape), dtype=slab.dtype, device=slab.device)
    if slab.is_cuda:
        work = dist.all_gather_into_tensor(out, slab, group=group, async_op=async_op)
        if async_op:                        # NCCL runs on its own stream; work.wait() orders the caller's stream after it
            return out, out[:, -1, :1].view(torch.int32)[:, 0], work
    else:                                   # gloo (CPU tests)
        parts = [torch.empty_like(slab) for _ in range(world)]
        dist.all_gather(parts, slab, group=group)
        out = torch.stack(parts, 0)
    return out, out[:, -1, :1].view(torch.int32)[:, 0]


class PeerExchange:
    """This rank's receive buffer, mapped by every rank of ``group``, plus the epoch bookkeeping of the push / wait / ack
    protocol (include/mobody_b200.h: mobody_peer_desc).  torch only allocates the memory and carries the handles
    (torch.distributed._symmetric_memory rendezvous, or legacy CUDA IPC through torch.multiprocessing.reductions); every
    byte of the exchange is moved by rollout_pack_push_kernel."""

    def __init__(self, cap_rows, W, device, group=None, mode=None):
        from . import _ffi
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("PeerExchange: at most 8 ranks (one NVSwitch box)")
        self.cap_rows, self.W, self.device = int(cap_rows), int(W), torch.device(device)
        lib = _ffi.lib()
        self.slot_floats = int(lib.mobody_peer_slot_floats(self.cap_rows, self.W))
        n_floats = int(lib.mobody_peer_buffer_bytes(self.world, self.cap_rows, self.W)) // 4
        mode = mode or os.environ.get("MOBODY_PEER_MODE", "symm")
        self.buf, self._keep, ptrs, self.multicast = None, None, None, 0
        if mode == "symm":
            try:
                import torch.distributed._symmetric_memory as symm_mem
                buf = symm_mem.empty(n_floats, dtype=torch.float32, device=self.device)
                buf.zero_()
                hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
                ptrs = [int(q) for q in hdl.buffer_ptrs]
                self.buf, self._keep, self.mode = buf, hdl, "symmetric_memory"
                mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                if mc and os.environ.get("MOBODY_PUSH_MULTICAST", "1") != "0":
                    self.multicast = mc                              # NVLS: one store, replicated to every rank by the switch
            except Exception as e:                                   # noqa: BLE001 -- fall through to legacy IPC
                self._symm_error = repr(e)
                mode = "ipc"
        if mode == "ipc":
            from torch.multiprocessing.reductions import reduce_tensor
            buf = torch.zeros(n_floats, dtype=torch.float32, device=self.device)
            fn, args = reduce_tensor(buf)
            objs = [None] * self.world
            dist.all_gather_object(objs, (fn, args), group=group)
            peers, ptrs = [], []
            for r, (f, a) in enumerate(objs):
                if r == self.rank:
                    peers.append(buf)
                else:
                    a = list(a); a[6] = self.device.index            # map the peer allocation into THIS device's context
                    peers.append(f(*a))
                ptrs.append(peers[-1].data_ptr())
            self.buf, self._keep, self.mode = buf, peers, "cuda_ipc"
        if ptrs is None:
            raise RuntimeError(f"PeerExchange: unknown mode {mode!r}")
        self.ptrs = ptrs
        # every rank's buffer as a tensor in THIS process (device-to-device copies of whole slots: the copy-engine push)
        if self.mode == "symmetric_memory":
            self.peer_views = [self.buf if r == self.rank else self._keep.get_buffer(r, (n_floats,), torch.float32) for r in range(self.world)]
        else:
            self.peer_views = list(self._keep)
        self.flags_off = int(lib.mobody_peer_buffer_bytes(self.world, self.cap_rows, self.W)) - 128     # arrive[8] | ack[8] | ticket | pad
        # push = "dma": whole slots by the copy engines + stream memory operations for the flags (no SM involved);
        #        "sm":  the narrow NVLink-store kernel (only the kept rows travel; one multicast store when the fabric offers it)
        # Measured (100 000 rows per rank, two compute streams): N=2 dma 161.9-164.7 M vs sm 160.7-162.4 M transitions/s (noise);
        # N=4 304.5 vs 309.4; N=8 541.7 vs 590.6 -- N - 1 unicast copies of the slot against ONE multicast store stream.  Default: the copy engines for
        # two ranks or when the fabric offers no multicast mapping, the multicast kernel otherwise.
        self.push = os.environ.get("MOBODY_PUSH") or ("dma" if (self.world <= 2 or not self.multicast) else "sm")
        self.memops = None
        if self.push == "dma":
            try:
                self.memops = _StreamMemOps()
                probe = torch.cuda.Stream(self.device)
                self.memops.write(probe, self.ptrs[self.rank] + self.flags_off + 4 * 20, 0)            # a pad word of the local flag block
                probe.synchronize()
            except Exception as e:                                   # noqa: BLE001 -- no stream memory operations here: kernel push
                self._memops_error, self.memops, self.push = repr(e), None, "sm"
        self.epoch = 0
        self._push_done = [None, None]                               # event of the push kernel that last read workspace slot parity
        # the push kernel is short and bandwidth bound: a high-priority side stream lets its few CTAs take the first SMs that
        # free up instead of queueing behind the next rollout's 700+ step tiles
        prio = int(os.environ.get("MOBODY_PUSH_PRIO", "-1"))
        self._side = torch.cuda.Stream(self.device, priority=prio)
        self.ctas = int(os.environ.get("MOBODY_PUSH_CTAS", "0"))
        self.overlap = os.environ.get("MOBODY_PUSH_OVERLAP", "1") != "0"
        self._descs = {}
        torch.cuda.synchronize(self.device)
        dist.barrier(group)                                          # every buffer is zeroed before anyone pushes into it

    def desc(self, epoch, ctas=0):
        from . import _ffi
        d = _ffi.PeerDesc()
        d.world, d.rank, d.cap_rows, d.W, d.epoch, d.ctas = self.world, self.rank, self.cap_rows, self.W, int(epoch), int(ctas)
        for r, q in enumerate(self.ptrs):
            d.base[r] = q
        d.multicast = self.multicast or None
        return d

    def views(self, epoch):
        """(rows [world, cap_rows, W], header int32 [world, 5]) of the half that holds ``epoch`` in the local buffer."""
        half = self.world * self.slot_floats
        h = self.buf[(epoch & 1) * half:((epoch & 1) + 1) * half].view(self.world, self.slot_floats)
        rows = h[:, :self.cap_rows * self.W].view(self.world, self.cap_rows, self.W)
        hdr = h[:, self.cap_rows * self.W:self.cap_rows * self.W + 5].view(torch.int32)
        return rows, hdr


class GatheredRollout:
    """Result of a p2p-gathered rollout: ``wait()`` makes the current stream wait (on the device) until every rank's rows have
    landed; ``rows`` [world, cap, W] / ``header`` int32 [world, 5] are views of the local receive buffer, valid until the
    second-next sharded_rollout call on this exchange (stream-ordered consumers are covered by the ack protocol)."""

    def __init__(self, ex, epoch, widths, info):
        self.ex, self.epoch, self.widths, self.info = ex, epoch, widths, info
        self.rows, self.header = ex.views(epoch)
        self._waited = False

    def wait(self):
        from . import _ffi
        if os.environ.get("MOBODY_PROBE", "") == "nopush":
            return self
        if not self._waited:
            ex = self.ex
            if ex.push == "dma":
                cur = torch.cuda.current_stream(ex.device)
                for r in range(ex.world):
                    ex.memops.wait_geq(cur, ex.ptrs[ex.rank] + ex.flags_off + 4 * r, self.epoch)
            else:
                d = ex.desc(self.epoch)
                _ffi.check(_ffi.lib().mobody_peer_wait(C.byref(d), _ffi.stream_ptr(ex.device)))
            self._waited = True
        return self

    def counts(self):
        """Host copy of the header: (kept per rank, produced transitions, reward sum).  One synchronising read."""
        self.wait()
        h = self.header.contiguous().cpu()
        kept = [int(v) for v in h[:, 0]]
        produced = int(sum((int(h[r, 1]) & 0xFFFFFFFF) | (int(h[r, 2]) << 32) for r in range(h.shape[0])))
        rsum = float(h[:, 3:5].contiguous().view(torch.float64).sum())
        return kept, produced, rsum

    def to_dict(self):
        kept, produced, rsum = self.counts()
        allp = torch.cat([self.rows[r, :kept[r]] for r in range(len(kept))], dim=0)
        return unpack_transitions(allp, self.widths), {"num_transitions": produced, "reward_mean": rsum / max(produced, 1),
                                                       "kept": int(sum(kept)), "kept_per_rank": kept, "exchange": self.ex.mode}


def p2p_rollout(agent, local, T, use_trg, row0, cap, *, group=None, step0=None, exchange=None, ctas=0, verify_images=True):
    """This rank's shard rolled on the device, packed into its own slot and pushed to every rank (csrc/peer.cu).  Asynchronous: returns a
    GatheredRollout handle; nothing is read back.  The push kernel runs on a side stream so that it overlaps whatever the
    caller enqueues next (e.g. the next rollout); workspaces alternate with the exchange parity.
    Consecutive calls may be enqueued on TWO alternating streams (epoch parity = stream), so that one rollout's partial last
    round of tiles overlaps the next rollout's first: everything keyed on the parity (workspace, receive-buffer half, the
    ack of epoch e - 2) then lives on one stream -- the caller must enqueue ``wait()`` of a handle on the stream it was
    produced on, and verify the weight images once before the fork (``verify_images=False`` here)."""
    from . import _ffi
    lib, dev = _ffi.lib(), local.device
    S, A = local.shape[1], agent.config["action_dim"]
    W = 2 * S + A + 3
    ex = exchange
    if ex is None:
        cache = agent.__dict__.setdefault("_peer_exchanges", {})
        key = (int(cap), W, id(group))
        ex = cache.get(key)
        if ex is None:
            ex = cache[key] = PeerExchange(cap, W, dev, group)
    if cap > ex.cap_rows or W != ex.W:
        raise ValueError("p2p_rollout: exchange is too small for this rollout")
    ex.epoch += 1
    e = ex.epoch
    cur = torch.cuda.current_stream(dev)
    pd = ex.desc(e, ctas or ex.ctas)
    probe = os.environ.get("MOBODY_PROBE", "")    # timing experiments only: 'nopush' = rollout into the symmetric slot, nothing else
    dma, mo = ex.push == "dma", ex.memops
    if e > 2 and probe != "nopush":    # everything enqueued on this stream so far has consumed epoch e - 2: peers may overwrite that half now
        if dma:
            for r in range(ex.world):
                mo.write(cur, ex.ptrs[r] + ex.flags_off + 32 + 4 * ex.rank, e - 2)
        else:
            _ffi.check(lib.mobody_peer_ack(C.byref(pd), e - 2, _ffi.stream_ptr(dev)))
    if ex._push_done[e & 1] is not None:
        cur.wait_event(ex._push_done[e & 1])                         # the push of epoch e - 2 read the slot / counters we are about to reuse
    B = local.shape[0]
    local = _ffi.f32(local, dev)
    ws = agent._rollout_workspace(T, B, S, A, 4 + (e & 1))
    rows, _ = ex.views(e)
    d, keep = agent._rollout_desc(local, T, use_trg, ws, rows[ex.rank], row0=row0, step0=step0, verify_images=verify_images)   # packs into OUR slot of the local buffer
    _ffi.check(lib.mobody_rollout(C.byref(d), _ffi.stream_ptr(dev)))
    kept_dev, stats_dev = ws["counts"][T + 1:T + 2], ws["stats"][:2]
    if probe == "nopush":
        done = torch.cuda.Event(); done.record(cur)
    elif dma:
        # copy-engine push: header row written locally, then the whole slot goes to every peer as ONE device-to-device copy
        # each, bracketed by stream memory operations (peer acked epoch e - 2 -> copy -> arrive flag = e).  No SM involved.
        _ffi.check(lib.mobody_peer_header(C.byref(pd), kept_dev.data_ptr(), stats_dev.data_ptr(), _ffi.stream_ptr(dev)))
        ev = torch.cuda.Event(); ev.record(cur)
        side = ex._side
        side.wait_event(ev)
        off = (e & 1) * ex.world * ex.slot_floats + ex.rank * ex.slot_floats
        src = ex.buf[off:off + ex.slot_floats]
        with torch.cuda.stream(side):
            for k in range(1, ex.world):
                r = (ex.rank + k) % ex.world                          # staggered: at any moment every rank is written by one peer
                if e > 2:
                    mo.wait_geq(side, ex.ptrs[ex.rank] + ex.flags_off + 32 + 4 * r, e - 2)
                ex.peer_views[r][off:off + ex.slot_floats].copy_(src, non_blocking=True)
                mo.write(side, ex.ptrs[r] + ex.flags_off + 4 * ex.rank, e)
            mo.write(side, ex.ptrs[ex.rank] + ex.flags_off + 4 * ex.rank, e)
        done = torch.cuda.Event(); done.record(side)
    elif ex.overlap:
        ev = torch.cuda.Event(); ev.record(cur)
        ex._side.wait_event(ev)
        _ffi.check(lib.mobody_peer_push(C.byref(pd), kept_dev.data_ptr(), stats_dev.data_ptr(), C.c_void_p(ex._side.cuda_stream)))
        done = torch.cuda.Event(); done.record(ex._side)
    else:
        _ffi.check(lib.mobody_peer_push(C.byref(pd), kept_dev.data_ptr(), stats_dev.data_ptr(), _ffi.stream_ptr(dev)))
        done = torch.cuda.Event(); done.record(cur)
    ex._push_done[e & 1] = done
    ex._descs[e & 1] = (d, keep, pd)                                 # keep-alive until the slot is reused
    info = {"kept_dev": kept_dev, "counts_dev": ws["counts"], "stats_dev": stats_dev, "capacity": cap,
            "world": ex.world, "exchange": ex.mode + ("+copy-engine push" if ex.push == "dma" else "+multicast" if ex.multicast else "")}
    return GatheredRollout(ex, e, [S, A, S, 1, 1, 1], info)


def sharded_rollout_host(agent, host_shard, rollout_length, use_trg=True, *, group=None, step0=None, host_output="shard"):
    """MOBODY.rollout's host contract (mobody.py:596-657: start states in, dict of CPU tensors out) for start states sharded
    over the ranks: THIS rank's shard [B_r, S] comes from (pinned) host memory; H2D of the shard, the rollout and the
    peer-memory exchange (after which EVERY rank's device buffer holds the transitions of ALL ranks -- what each training
    replica consumes) happen inside the call, and so does the D2H of the result:
      host_output="shard": this rank's own transitions come back as CPU tensors -- over all ranks the job's gathered output
                           reaches the host exactly once, in parallel over the GPUs' PCIe links;
      host_output="all":   the transitions of all ranks come back on every rank, rank-major (world x the D2H traffic).
    ``info`` always describes the whole job (num_transitions / reward_mean over all ranks, kept_per_rank)."""
    dev = agent.device
    T = int(rollout_length)
    x = host_shard.to(dev, non_blocking=True)
    res = sharded_rollout(agent, x, T, use_trg, group=group, gather="p2p", sharded_input=True, step0=step0)
    kept, produced, rsum = res.counts()                                   # waits for every rank's rows, one small D2H
    W = res.rows.shape[2]
    take = list(range(len(kept))) if host_output == "all" else [res.ex.rank]
    host = agent._host_slab(sum(kept[r] for r in take), W)
    off = 0
    for r in take:
        m = kept[r]
        host[off:off + m].copy_(res.rows[r, :m], non_blocking=True)
        off += m
    torch.cuda.current_stream(dev).synchronize()
    return unpack_transitions(host[:off], res.widths), {"num_transitions": produced, "reward_mean": rsum / max(produced, 1),
                                                        "kept": off, "kept_per_rank": kept}


def _exchange_mode(agent):
    for ex in getattr(agent, "_peer_exchanges", {}).values():
        return ex.mode + ("+copy-engine push" if ex.push == "dma" else "+multicast" if ex.multicast else "")
    return None


def sharded_rollout(agent, init_obss, rollout_length, use_trg=True, *, group=None, gather=True, sharded_input=False, step0=None,
                    exchange=None, verify_images=True):
    """MOBODY.rollout over all ranks of ``group``.

    init_obss: the GLOBAL start states [B, S] (every rank passes the same tensor and takes its shard), or with
    ``sharded_input=True`` this rank's own shard (ranks must hold equal-sized shards; weak-scaling benches).
    gather: False -> this rank's transitions only; "p2p" -> GatheredRollout handle (peer-memory push, asynchronous, see
    p2p_rollout); "p2p_dict" -> the same, waited and compacted to a dict (one host read);
    "padded" -> ([world, cap+1, W] slabs, device counts, widths)
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
        return agent.rollout_device(local, T, use_trg, row0=lo, step0=step0)
    if gather in ("p2p", "p2p_dict"):
        res = p2p_rollout(agent, local, T, use_trg, lo, cap, group=group, step0=step0, exchange=exchange, verify_images=verify_images)
        return res if gather == "p2p" else res.to_dict()
    S = local.shape[1]
    probe_w = getattr(agent, "config", {}).get("action_dim")
    W = 2 * S + probe_w + 3 if probe_w is not None else None
    if W is None or not local.is_cuda:       # generic path (CPU tests with a stand-in agent)
        out, info = agent.rollout_device(local, T, use_trg, row0=lo) if step0 is None else agent.rollout_device(local, T, use_trg, row0=lo, step0=step0)
        packed, widths = pack_transitions(out)
        allp, counts = allgather_transitions(packed, cap, group)
        stats = torch.tensor([info["num_transitions"], info["reward_mean"] * info["num_transitions"]], dtype=torch.float64,
                             device=packed.device)
        dist.all_reduce(stats, group=group)
        n = int(stats[0].item())
        return unpack_transitions(allp, widths), {"num_transitions": n, "reward_mean": float(stats[1].item()) / max(n, 1),
                                                  "kept": int(sum(counts)), "kept_per_rank": counts}
    slab = torch.empty(cap + 1, W, dtype=torch.float32, device=local.device)
    out, info = agent.rollout_device(local, T, use_trg, row0=lo, out_packed=slab[:cap], sync=False, step0=step0)   # nothing read back
    # in-band header row, integer words (exact at any size): int32 kept rows, int32 produced transitions, float64 reward sum
    hdr = slab[cap, :4].view(torch.int32)
    hdr[0:1] = info["kept_dev"]
    hdr[1:2] = info["stats_dev"][1:2].to(torch.int32)
    hdr[2:4] = info["stats_dev"][0:1].clone().view(torch.int32)        # the double's two words (the row need not be 8-byte aligned)
    widths = [S, probe_w, S, 1, 1, 1]
    if gather == "padded_async":
        slabs, counts_dev, work = allgather_slabs(slab, group, async_op=True)
        return (slabs, counts_dev, widths, work), dict(info, world=world, capacity=cap)
    slabs, counts_dev = allgather_slabs(slab, group)
    if gather == "padded":
        return (slabs, counts_dev, widths), dict(info, world=world, capacity=cap)
    hdr = slabs[:, cap, :4].contiguous().cpu()                         # the one host read of the exchange
    hi = hdr.view(torch.int32)
    counts = [int(c) for c in hi[:, 0]]
    allp = torch.cat([slabs[r, :counts[r]] for r in range(world)], dim=0)
    n = int(hi[:, 1].sum())
    rsum = float(hi[:, 2:4].contiguous().view(torch.float64).sum())
    return unpack_transitions(allp, widths), {"num_transitions": n, "reward_mean": rsum / max(n, 1),
                                              "kept": int(sum(counts)), "kept_per_rank": counts}


def self_check(agent, n_rows=10_007, group=None, rounds=5, seed=0):
    """Correctness of the multi-rank assembly on live hardware: the transitions gathered from the ranks' shards must equal,
    bit for bit, what THIS rank computes alone for all ``n_rows`` start states (Philox is keyed on the global row id, so a
    row's result does not depend on its shard).  T = 1: same rows in the same order (rank-major == row order); T = 3: same
    multiset (rank-major differs from the single-GPU step-major order only by a permutation).  The peer-memory exchange is
    run ``rounds`` times back to back on fresh draws (both halves of the receive buffer, the ack hand-shake) and once
    through the NCCL padded all-gather.  Collective: every rank must call it.  Returns {"ok": bool, ...}; never raises on a
    mismatch (the caller reports it)."""
    import numpy as np
    dev = agent.device
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    S = agent.config["state_dim"]
    g = torch.Generator().manual_seed(1234 + seed)
    obs = (0.3 * torch.randn(n_rows, S, generator=g)).to(dev)          # same on every rank
    res = {"world": world, "rows": n_rows}
    ok = True
    key = lambda m: m[np.lexsort(m.cpu().numpy().T[::-1])] if m.numel() else m    # noqa: E731
    base = agent.dynamics._draw
    ex_mode = None
    for i in range(rounds):
        T = 1 if i % 2 == 0 else 3
        s0 = 1000 + 10 * i
        full, fi = agent.rollout_device(obs, T, step0=s0)
        want = fi["packed"][:fi["kept"]]
        try:
            out, info = sharded_rollout(agent, obs, T, group=group, gather="p2p_dict", step0=s0)
            got = pack_transitions(out)[0]
            ex_mode = info.get("exchange", ex_mode)
            same = got.shape == want.shape and info["num_transitions"] == fi["num_transitions"] and info["kept"] == fi["kept"]
            same = same and bool(torch.equal(got, want) if T == 1 else torch.equal(key(got), key(want)))
        except Exception as e:                                            # noqa: BLE001
            same, res["p2p_error"] = False, repr(e)
        ok = ok and same
    res["p2p"] = ok
    full, fi = agent.rollout_device(obs, 1, step0=77)
    (slabs, counts_dev, widths), _ = sharded_rollout(agent, obs, 1, group=group, gather="padded", step0=77)
    cnt = [int(c) for c in counts_dev.cpu()]
    nccl_ok = sum(cnt) == fi["kept"] and bool(torch.equal(torch.cat([slabs[r, :cnt[r]] for r in range(world)], 0), fi["packed"][:fi["kept"]]))
    res["nccl"] = nccl_ok
    agent.dynamics._draw = base
    flag = torch.tensor([1.0 if ok else 0.0, 1.0 if nccl_ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    res["p2p"], res["nccl"] = bool(flag[0].item() == 1.0), bool(flag[1].item() == 1.0)
    res["ok"] = res["p2p"] and res["nccl"]
    return res
