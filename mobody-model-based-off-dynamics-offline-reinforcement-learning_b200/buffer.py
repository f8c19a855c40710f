"""Device-resident ReplayBuffer on packed transition rows.

Mirror of algo/utils.py:13-193 (``ReplayBuffer``): same constructor, ``add`` / ``add_batch`` /
``sample`` / ``sample_all`` / ``convert_D4RL``, and the ``state, action, next_state, reward,
not_done, size, ptr, max_size`` fields the reference pokes directly (train_mobody.py:551,557;
mobody.py:381).  Storage is one CUDA tensor ``rows[max_size, RW]`` with
RW = roundup4(2S+A+2): [state | action | next_state | reward | not_done | pad]; the named fields
are strided views of it.  ``sample`` is a 128-bit row-gather kernel, ``add_batch`` a ring-insert
kernel; indices come from Philox4x32-10 on the device (or are injected for parity tests).
"""
import numpy as np
import torch

from . import _ffi


class ReplayBuffer(object):
    _streams = 0          # buffers created so far in this process: default Philox streams are distinct per buffer

    @staticmethod
    def stream_seed(seed, stream):
        """32-bit Philox key of index stream ``stream`` (an int or a role name) under run seed ``seed``: buffers of one run
        draw from independent streams, like the reference's successive np.random.randint calls (utils.py:128)."""
        if isinstance(stream, str):
            stream = int.from_bytes(stream.encode()[:4].ljust(4, b"\0"), "little")
        x = (int(seed) * 0x9E3779B1 + int(stream) * 0x85EBCA77 + 0x165667B1) & 0xFFFFFFFF
        x ^= x >> 15; x = (x * 0x2C1B3C6D) & 0xFFFFFFFF; x ^= x >> 12; x = (x * 0x297A2D39) & 0xFFFFFFFF; x ^= x >> 15
        return x

    def __init__(self, state_dim, action_dim, device, max_size=int(1e6), seed=None, index_source="philox"):
        self.S, self.A = int(state_dim), int(action_dim)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mobody_b200.ReplayBuffer is device resident: pass a CUDA device")
        self.RW = _ffi.lib().mobody_row_width(self.S, self.A)
        self.max_size, self.ptr, self.size = int(max_size), 0, 0
        self._rows = torch.zeros(self.max_size, self.RW, dtype=torch.float32, device=self.device)
        self.mobile = 0
        if seed is None:                       # a stream of its own: src / tar / fake samples are independent
            ReplayBuffer._streams += 1
            seed = ReplayBuffer.stream_seed(0, ReplayBuffer._streams)
        self.seed, self._draw = int(seed), 0
        self.index_source = index_source       # "philox" (device) | "numpy" (np.random.randint, like the reference)

    # ---- reference field names as views of the packed storage ----
    def _view(self, lo, hi):
        return self._rows[:, lo:hi]

    def _assign(self, lo, hi, value):
        v = torch.as_tensor(value)
        if v.data_ptr() == self._rows[:, lo:hi].data_ptr() and v.shape[0] == self._rows.shape[0]:
            return                                  # `buf.reward -= 1.0` re-assigns the same view
        v = v.to(device=self.device, dtype=torch.float32).reshape(-1, hi - lo)
        if v.shape[0] > self._rows.shape[0]:
            raise ValueError("assigned field has more rows than the buffer")
        self._rows[: v.shape[0], lo:hi].copy_(v)

    state = property(lambda s: s._view(0, s.S), lambda s, v: s._assign(0, s.S, v))
    action = property(lambda s: s._view(s.S, s.S + s.A), lambda s, v: s._assign(s.S, s.S + s.A, v))
    next_state = property(lambda s: s._view(s.S + s.A, 2 * s.S + s.A), lambda s, v: s._assign(s.S + s.A, 2 * s.S + s.A, v))
    reward = property(lambda s: s._view(2 * s.S + s.A, 2 * s.S + s.A + 1), lambda s, v: s._assign(2 * s.S + s.A, 2 * s.S + s.A + 1, v))
    not_done = property(lambda s: s._view(2 * s.S + s.A + 1, 2 * s.S + s.A + 2), lambda s, v: s._assign(2 * s.S + s.A + 1, 2 * s.S + s.A + 2, v))

    def split(self, rows):
        """packed rows [n,RW] -> the reference's 5-tuple (state, action, next_state, reward, not_done) as views."""
        S, A = self.S, self.A
        return (rows[:, :S], rows[:, S:S + A], rows[:, S + A:2 * S + A], rows[:, 2 * S + A:2 * S + A + 1],
                rows[:, 2 * S + A + 1:2 * S + A + 2])

    # ---- inserts ----
    def _pack(self, s, a, ns, r, d, done_is_terminal):
        dev = self.device
        s, a, ns = _ffi.f32(s, dev).reshape(-1, self.S), _ffi.f32(a, dev).reshape(-1, self.A), _ffi.f32(ns, dev).reshape(-1, self.S)
        n = s.shape[0]
        r, d = _ffi.f32(r, dev).reshape(n), _ffi.f32(d, dev).reshape(n)
        out = torch.empty(n, self.RW, dtype=torch.float32, device=dev)
        _ffi.check(_ffi.lib().mobody_pack_rows(_ffi.ptr(s), _ffi.ptr(a), _ffi.ptr(ns), _ffi.ptr(r), _ffi.ptr(d), n,
                                               self.S, self.A, int(done_is_terminal), _ffi.ptr(out), _ffi.stream_ptr(dev)))
        return out

    def add(self, state, action, next_state, reward, done):          # utils.py:32-41
        row = self._pack(np.asarray(state, np.float32)[None], np.asarray(action, np.float32)[None],
                         np.asarray(next_state, np.float32)[None], np.asarray([reward], np.float32),
                         np.asarray([done], np.float32), True)
        self.add_packed(row, 1)

    def add_batch(self, batch):                                       # utils.py:43-92
        if batch is None:
            return
        rows = self._pack(batch["obss"], batch["actions"], batch["next_obss"], batch["rewards"], batch["terminals"], True)
        self.add_packed(rows, rows.shape[0])

    def add_rollout_slab(self, packed, n):
        """add_batch (utils.py:43-92) of the first ``n`` rows of a rollout slab [obs | act | next_obs | reward | terminal |
        penalty] (MOBODY.rollout_device): one kernel converts the row layout (not_done = 1 - terminal) and ring-inserts."""
        n = int(n)
        if n == 0:
            return
        if n > self.max_size:
            raise ValueError(f"add_batch of {n} rows exceeds buffer capacity {self.max_size}")
        if packed.shape[1] != 2 * self.S + self.A + 3 or not packed.is_contiguous():
            raise ValueError("rollout slab must be contiguous [M, 2S+A+3]")
        _ffi.check(_ffi.lib().mobody_ring_insert_transitions(_ffi.ptr(packed), n, None, self.S, self.A, self.ptr, self.max_size,
                                                             _ffi.ptr(self._rows), _ffi.stream_ptr(self.device)))
        self._advance(n)

    def _advance(self, n):
        end = min(self.ptr + n, self.max_size)
        used = end - self.ptr
        self.ptr = end % self.max_size
        self.size = min(self.size + used, self.max_size)
        if self.ptr == 0:
            self.ptr = n - used

    def add_packed(self, rows, n, n_dev=None):
        """Ring insert of ``n`` packed device rows (single wrap, like the reference).  ``n`` is the host
        count used for ptr/size bookkeeping; ``n_dev`` optionally bounds the copy on the device."""
        n = int(n)
        if n == 0:
            return
        if n > self.max_size:
            # the reference raises a shape error here (utils.py:85-90 handles one wrap only)
            raise ValueError(f"add_batch of {n} rows exceeds buffer capacity {self.max_size}")
        _ffi.check(_ffi.lib().mobody_ring_insert(_ffi.ptr(rows), n, _ffi.ptr(n_dev), self.RW, self.ptr, self.max_size,
                                                 _ffi.ptr(self._rows), _ffi.stream_ptr(self.device)))
        self._advance(n)

    # ---- sampling ----
    def draw_indices(self, batch_size, ind=None):
        """int64[batch] on the device: injected, NumPy-global-RNG (reference behaviour) or Philox."""
        if ind is None and self.index_source == "numpy":
            ind = np.random.randint(0, self.size, size=batch_size)   # utils.py:128
        if ind is not None:
            t = torch.as_tensor(np.asarray(ind)) if not torch.is_tensor(ind) else ind
            return t.to(device=self.device, dtype=torch.int64).contiguous()
        if self.size <= 0:
            raise ValueError("sample from an empty ReplayBuffer")      # np.random.randint(0, 0) raises too
        idx = torch.empty(batch_size, dtype=torch.int64, device=self.device)
        _ffi.check(_ffi.lib().mobody_philox_indices(_ffi.ptr(idx), batch_size, self.seed, self._draw, self.size,
                                                    _ffi.stream_ptr(self.device)))
        self._draw += 1
        return idx

    def sample_rows(self, batch_size, ind=None, out=None):
        """Packed rows [batch,RW] gathered on the device (the train step consumes this layout)."""
        idx = self.draw_indices(batch_size, ind)
        if out is None:
            out = torch.empty(batch_size, self.RW, dtype=torch.float32, device=self.device)
        _ffi.check(_ffi.lib().mobody_gather_rows(_ffi.ptr(self._rows), _ffi.ptr(idx), batch_size, self.RW, _ffi.ptr(out),
                                                 _ffi.stream_ptr(self.device)))
        return out

    def sample(self, batch_size, ind=None):                           # utils.py:127-148
        return self.split(self.sample_rows(batch_size, ind))

    def sample_all(self, cuda=True):                                  # utils.py:150-167
        parts = self.split(self._rows[: self.size])
        if cuda:
            return tuple(p.contiguous() for p in parts)
        return tuple(p.cpu() for p in parts)

    def convert_D4RL(self, dataset):                                  # utils.py:173-193
        n = int(np.asarray(dataset["observations"]).shape[0])
        rows = self._pack(dataset["observations"], dataset["actions"], dataset["next_observations"],
                          np.asarray(dataset["rewards"]).reshape(-1), np.asarray(dataset["terminals"]).reshape(-1), True)
        # the dataset becomes the storage: capacity follows it, so a later add() / add_batch() (the online modes,
        # train_mobody.py:693, 743) wraps inside the allocation instead of running past it
        self._rows = rows
        self.size, self.max_size, self.ptr = n, max(n, 1), 0
