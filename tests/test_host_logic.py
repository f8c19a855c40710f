"""CPU tests of host-side scheduling helpers (no GPU, no library calls)."""
import numpy as np
import pytest

from mobody_b200.mobody import pipe_bounds, _wgrad_splits
from mobody_b200.buffer import ReplayBuffer


def test_pipe_bounds_cover_rows_in_order_with_whole_wave_chunks():
    for sms in (148, 132):                                                   # derived from the device, not hard-coded
        WAVE = sms * 128
        ROWS = 2 * WAVE
        for B in (2 * ROWS, 100_000, 94_720, 94_721, 113_664, 120_000, 200_000, 1_000_003):
            b = pipe_bounds(B, WAVE)
            assert b[0] == 0 and b[-1] == B and all(x < y for x, y in zip(b, b[1:])), (B, b)
            sizes = np.diff(b)
            assert sizes[0] == WAVE                                              # short H2D before the first kernel
            assert all(s % WAVE == 0 for s in sizes[:-1]), (B, sizes)            # only the last chunk may hold a partial wave
            assert sizes.max() <= ROWS + WAVE
            if (B - WAVE) % WAVE and 0 < (B - WAVE) % ROWS % WAVE <= WAVE // 2:
                assert sizes[-1] < WAVE                                          # small remainder -> its own (cheap D2H) chunk


def test_buffer_index_streams_are_distinct():
    """Every buffer draws its indices from its own Philox stream (src / tar / fake samples must not be rank-correlated)."""
    seeds = {ReplayBuffer.stream_seed(0, i) for i in range(1, 200)} | {ReplayBuffer.stream_seed(0, "fake")}
    assert len(seeds) == 200 and all(0 <= s < 2 ** 32 for s in seeds)
    assert ReplayBuffer.stream_seed(1, "fake") != ReplayBuffer.stream_seed(0, "fake")
    assert ReplayBuffer.stream_seed(7, 3) == ReplayBuffer.stream_seed(7, 3)


def test_wgrad_splits_fill_whole_rounds():
    # batch 4096 (10 240 rows) on 148 SMs: 20 critic tiles and 10 actor tiles per split, two CTAs per SM
    n = _wgrad_splits(10_240, (20, 10), None, sm_count=148)
    assert 1 <= n <= 64 and 10 * n <= 296 and 20 * n <= 2 * 296               # actor launch: one round, critic: two
    # small batch: one 32-row chunk per CTA at most, never more splits than chunks
    assert _wgrad_splits(320, (20, 10), None, sm_count=148) == 10
    assert _wgrad_splits(20, (20, 10), None, sm_count=148) == 1
    for rows in (1, 31, 33, 77, 1500, 4096, 50_000):
        n = _wgrad_splits(rows, (20, 10), None, sm_count=148)
        assert 1 <= n <= min(64, (rows + 31) // 32)


def test_dataset_walk_matches_loop_restatement(tmp_path):
    """transitions_from_raw == the reference's N-1 row Python loop (dataset/call_dataset.py:59-110), for [N] and [N,1]
    rewards, float64 sources, with / without a timeouts field, and the degenerate sizes; .npz round trip of read_raw."""
    import numpy as np
    from oracle.ingest_oracle import transitions_loop
    from mobody_b200.dataset import transitions_from_raw, read_raw, tar_dataset_path, call_tar_dataset
    rng = np.random.default_rng(0)
    for n, rshape, with_to in ((257, (257,), False), (64, (64, 1), True), (2, (2,), False), (1, (1,), False)):
        raw = {"observations": rng.standard_normal((n, 11)), "actions": rng.uniform(-1, 1, (n, 3)).astype(np.float32),
               "rewards": rng.standard_normal(rshape), "terminals": rng.random(n) < 0.1}
        if with_to:
            raw["timeouts"] = rng.random(n) < 0.05
        got, want = transitions_from_raw(raw), transitions_loop(raw)
        for k in want:
            w = np.asarray(want[k])
            if w.size == 0:
                assert got[k].shape[0] == 0
                continue
            assert got[k].dtype == w.dtype and got[k].shape == w.shape and np.array_equal(got[k], w), (n, k)
    np.savez(tmp_path / "hopper_kinematic_2.0_medium.npz", **raw)
    back = read_raw(str(tmp_path / "hopper_kinematic_2.0_medium.npz"))
    assert set(back) == set(raw) and np.array_equal(back["observations"], raw["observations"])
    assert tar_dataset_path("/d", "hopper-kinematic", 2.0, "medium") == "/d/mujoco/hopper_kinematic_2.0_medium.hdf5"
    assert tar_dataset_path("/d", "antmaze-small", 1.0) == "/d/antmaze/antmaze_small_1.0.hdf5"
    (tmp_path / "mujoco").mkdir()
    np.savez(tmp_path / "mujoco" / "hopper_kinematic_2.0_medium.npz", **raw)
    assert call_tar_dataset("hopper-kinematic", 2.0, "medium", root=str(tmp_path))["observations"].shape[0] == 0
    with pytest.raises((ImportError, OSError)):
        read_raw(str(tmp_path / "missing.hdf5"))


class _VecEnv:
    """E independent toy episodes with scripted lengths; rewards depend on the action so the policy matters."""
    def __init__(self, lengths, S=5):
        import numpy as np
        self.lengths, self.S, self.E = np.asarray(lengths), S, len(lengths)

    def reset(self):
        import numpy as np
        self.t = 0
        self.rng = np.random.default_rng(1)
        return self.rng.standard_normal((self.E, self.S)).astype(np.float32)

    def step(self, action):
        import numpy as np
        self.t += 1
        ns = self.rng.standard_normal((self.E, self.S)).astype(np.float32)
        rew = np.asarray(action).reshape(self.E, -1).sum(1) + 0.1 * self.t
        return ns, rew, (self.t % self.lengths) == 0, {}          # a finished copy keeps reporting done periodically


class _Pol:
    def select_action(self, state, dist):
        import numpy as np
        return np.tanh(state[:, :2] * dist)


def test_batched_evaluator_matches_loop_restatement(capsys):
    """eval_policy_batch == train_mobody.py:53-98: per-copy return up to its FIRST done, visited transitions in the same order."""
    import numpy as np
    from oracle.ingest_oracle import eval_batch_loop
    from mobody_b200 import evaluate
    lengths = [7, 3, 12, 3, 9]
    want, visited = eval_batch_loop(_Pol(), _VecEnv(lengths), 0.5, len(lengths))
    seen = {}
    orig = evaluate._model_check
    evaluate._model_check = lambda dyn, s, a, ns, r: seen.update(s=np.asarray(s), a=np.asarray(a), ns=np.asarray(ns), r=np.asarray(r))
    try:
        got = evaluate.eval_policy_batch(_Pol(), _VecEnv(lengths), 0.5, len(lengths), eval_cnt=3, dynamics=object(), eval_trg=True)
    finally:
        evaluate._model_check = orig
    assert got == want
    assert np.array_equal(seen["s"], np.asarray(visited[0])) and np.array_equal(seen["a"], np.asarray(visited[1]))
    assert np.array_equal(seen["ns"], np.asarray(visited[2])) and np.array_equal(seen["r"], np.asarray(visited[3]))
    assert len(seen["r"]) == sum(lengths)
    assert "Evaluation on target over 5 episodes" in capsys.readouterr().out
