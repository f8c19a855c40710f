"""CPU tests of host-side scheduling helpers (no GPU, no library calls)."""
import numpy as np

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
