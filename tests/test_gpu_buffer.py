"""GPU parity of the byte/index kernels: bit-exact against reference-generated goldens and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_termination_bit_exact(golden_dir):
    import mobody_b200 as mb
    g = np.load(os.path.join(golden_dir, "termination.npz"))
    task = {"walker2d": "walker2d-medium-v2", "hopper": "hopper-medium-v2", "halfcheetah": "halfcheetah-medium-v2", "ant": "ant-medium-v2"}
    for env in task:
        fn = mb.get_termination_fn(task[env])
        x = g[env + "_x"]
        done = fn(x, x[:, :1], x)                              # numpy in -> numpy bool [B,1], like the reference
        assert done.dtype == np.bool_ and np.array_equal(done, g[env + "_done"]), env
        dev = fn(torch.from_numpy(x).cuda(), torch.from_numpy(x[:, :1]).cuda(), torch.from_numpy(x).cuda())
        assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), g[env + "_done"])
    never = mb.get_termination_fn("pendulum")
    assert not never(np.zeros((5, 3), np.float32), np.zeros((5, 1), np.float32), np.full((5, 3), np.nan, np.float32)).any()


def test_ring_buffer_matches_reference(golden_dir):
    import mobody_b200 as mb
    g = np.load(os.path.join(golden_dir, "buffer.npz"))
    S, A, cap = int(g["S"]), int(g["A"]), int(g["cap"])
    buf = mb.ReplayBuffer(S, A, "cuda", max_size=cap)
    buf.add_batch(None)
    for i in range(int(g["n_batches"])):
        buf.add_batch({k: torch.from_numpy(g[f"b{i}_{k}"]) for k in ("obss", "next_obss", "actions", "rewards", "terminals")})
        assert (buf.ptr, buf.size) == (int(g[f"after{i}_ptr"]), int(g[f"after{i}_size"]))
        for f in ("state", "action", "next_state", "reward", "not_done"):
            assert np.array_equal(getattr(buf, f).cpu().numpy(), g[f"after{i}_{f}"]), (i, f)
    smp = buf.sample(33, ind=g["ind"])
    for f, v in zip(("state", "action", "next_state", "reward", "not_done"), smp):
        assert v.is_cuda and np.array_equal(v.cpu().numpy(), g["sample_" + f])
    with pytest.raises(ValueError):
        buf.add_batch({k: torch.zeros(cap + 1, w) for k, w in (("obss", S), ("next_obss", S), ("actions", A), ("rewards", 1), ("terminals", 1))})
    # direct field pokes used by train_mobody.py:551 and mobody.py:381
    before = buf.reward.clone()
    buf.reward -= 1.0
    assert torch.equal(buf.reward, before - 1.0)
    buf.reward = torch.zeros(cap, 1)
    assert float(buf.reward.abs().sum()) == 0.0
    st_all = buf.sample_all(False)
    assert not st_all[0].is_cuda and st_all[0].shape == (buf.size, S)


def test_philox_indices_bit_exact_and_empty():
    import mobody_b200 as mb
    from oracle.philox import buffer_indices
    buf = mb.ReplayBuffer(3, 2, "cuda", max_size=5000, seed=99)
    with pytest.raises(ValueError):
        buf.sample(4)
    buf.size = 4321
    for draw in range(3):
        got = buf.draw_indices(1000).cpu().numpy()
        assert np.array_equal(got, buffer_indices(99, draw, 1000, 4321))
    assert buf.draw_indices(0).numel() == 0


def test_convert_d4rl_and_sample_shapes():
    import mobody_b200 as mb
    rng = np.random.default_rng(0)
    n, S, A = 777, 17, 6
    ds = dict(observations=rng.standard_normal((n, S)).astype(np.float32), actions=rng.standard_normal((n, A)).astype(np.float32),
              next_observations=rng.standard_normal((n, S)).astype(np.float32), rewards=rng.standard_normal(n).astype(np.float32),
              terminals=(rng.random(n) < 0.1))
    buf = mb.ReplayBuffer(S, A, "cuda")
    buf.convert_D4RL(ds)
    assert buf.size == n
    ind = rng.integers(0, n, 50)
    s, a, ns, r, nd = buf.sample(50, ind=ind)
    assert np.array_equal(s.cpu().numpy(), ds["observations"][ind]) and np.array_equal(a.cpu().numpy(), ds["actions"][ind])
    assert np.array_equal(ns.cpu().numpy(), ds["next_observations"][ind]) and np.array_equal(r.cpu().numpy()[:, 0], ds["rewards"][ind])
    assert np.array_equal(nd.cpu().numpy()[:, 0], 1.0 - ds["terminals"][ind].astype(np.float32))
    assert r.shape == (50, 1) and nd.shape == (50, 1)


@pytest.mark.parametrize("n", [0, 1, 1023, 1024, 1025, 70000, 2_500_000])
def test_compaction_stable_and_exact(n):
    from mobody_b200 import _ffi
    L, st = _ffi.lib(), _ffi.stream_ptr()
    rng = np.random.default_rng(n)
    flags = (rng.random(n) < 0.3).astype(np.uint8)
    vals = rng.standard_normal(n).astype(np.float32)
    if n > 10:
        vals[3] = np.nan; vals[5] = 0.25
    f, v = torch.from_numpy(flags).cuda(), torch.from_numpy(vals).cuda()
    pos = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    scr = torch.empty(int(L.mobody_compact_scratch_ints(n)), dtype=torch.int32, device="cuda")
    with np.errstate(invalid="ignore"):
        cases = [(_ffi.KEEP_U8_ZERO, np.flatnonzero(flags == 0)), (_ffi.KEEP_F32_LE, np.flatnonzero(vals <= 0.25)),
                 (_ffi.KEEP_F32_LT, np.flatnonzero(vals < 0.25)), (_ffi.KEEP_U8_VALID, np.flatnonzero(flags != 255))]
    for kind, want in cases:
        _ffi.check(L.mobody_compact(kind, _ffi.ptr(f), _ffi.ptr(v), 0.25, n, None, _ffi.ptr(scr), _ffi.ptr(pos), _ffi.ptr(cnt), st))
        c = int(cnt.item())
        assert c == len(want) and np.array_equal(pos[:c].cpu().numpy(), want)
    if n >= 1025:   # live count on the device bounds the scan
        live = torch.tensor([1000], dtype=torch.int32, device="cuda")
        _ffi.check(L.mobody_compact(_ffi.KEEP_U8_ZERO, _ffi.ptr(f), None, 0.0, n, _ffi.ptr(live), _ffi.ptr(scr), _ffi.ptr(pos), _ffi.ptr(cnt), st))
        want = np.flatnonzero(flags[:1000] == 0)
        assert int(cnt.item()) == len(want) and np.array_equal(pos[:len(want)].cpu().numpy(), want)


def test_fused_sample_rows_equals_separate_draw_and_gather():
    """mobody_sample_rows (one launch for the three buffer samples of a train step) draws the same Philox indices and
    gathers the same rows as philox_indices + gather_rows per buffer."""
    import mobody_b200 as mb
    from mobody_b200 import _ffi
    S, A = 17, 6
    rng = np.random.default_rng(0)
    bufs, sizes, ns = [], (5000, 700, 1234), (128, 128, 64)
    for n, sd in zip(sizes, (1, 2, 3)):
        b = mb.ReplayBuffer(S, A, "cuda", max_size=n, seed=sd)
        b.add_batch({"obss": rng.standard_normal((n, S)).astype(np.float32), "actions": rng.uniform(-1, 1, (n, A)).astype(np.float32),
                     "next_obss": rng.standard_normal((n, S)).astype(np.float32), "rewards": rng.standard_normal((n, 1)).astype(np.float32),
                     "terminals": np.zeros((n, 1), np.float32)})
        b._draw = 7 + sd
        bufs.append(b)
    RW = bufs[0].RW
    rows = torch.empty(sum(ns), RW, dtype=torch.float32, device="cuda")
    jobs = (_ffi.SampleJob * 3)()
    lo = 0
    for jb, b, n in zip(jobs, bufs, ns):
        jb.rows, jb.n, jb.size, jb.draw, jb.seed, jb.out = _ffi.ptr(b._rows), n, b.size, b._draw, b.seed, rows.data_ptr() + 4 * RW * lo
        lo += n
    _ffi.check(_ffi.lib().mobody_sample_rows(jobs, 3, RW, _ffi.stream_ptr(torch.device("cuda"))))
    want = torch.cat([b.sample_rows(n) for b, n in zip(bufs, ns)], 0)
    assert torch.equal(rows, want)


@pytest.mark.parametrize("S,A,n", [(17, 6, 100_003), (11, 3, 4096), (27, 8, 65_537), (29, 8, 9_999), (64, 32, 5_000), (17, 6, 4_095)])
def test_pack_rows_tile_path_bit_exact(S, A, n):
    """convert_D4RL / add_batch row packing (utils.py:43-92, 173-193) at sizes that take the 128-bit tile kernel (n >= 4096,
    aligned sources), with ragged last tiles, both done conventions, and the unaligned-source fallback: bit-exact vs torch.cat."""
    import mobody_b200 as mb
    from mobody_b200 import _ffi
    g = torch.Generator(device="cuda").manual_seed(S * 7 + A)
    s, a, ns = (torch.randn(n, w, device="cuda", generator=g) for w in (S, A, S))
    r = torch.randn(n, device="cuda", generator=g); d = (torch.rand(n, device="cuda", generator=g) < 0.1).float()
    buf = mb.ReplayBuffer(S, A, "cuda", max_size=8)
    for flip in (True, False):
        got = buf._pack(s, a, ns, r, d, flip)
        want = torch.zeros(n, buf.RW, device="cuda")
        want[:, :S], want[:, S:S + A], want[:, S + A:2 * S + A] = s, a, ns
        want[:, 2 * S + A], want[:, 2 * S + A + 1] = r, (1.0 - d if flip else d)
        assert torch.equal(got, want)
    # a source that is not 16-byte aligned (a view starting one float into an allocation) takes the scalar kernel: same result
    s_un = torch.empty(n * S + 1, device="cuda")[1:].view(n, S).copy_(s)
    assert s_un.data_ptr() % 16 != 0
    out = torch.empty(n, buf.RW, device="cuda")
    _ffi.check(_ffi.lib().mobody_pack_rows(s_un.data_ptr(), _ffi.ptr(a), _ffi.ptr(ns), _ffi.ptr(r), _ffi.ptr(d), n, S, A, 1, _ffi.ptr(out), _ffi.stream_ptr(s.device)))
    assert torch.equal(out, buf._pack(s, a, ns, r, d, True))
