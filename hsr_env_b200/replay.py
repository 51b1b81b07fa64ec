"""Device-resident replay buffer fed by ``BatchedHSREnv.step`` outputs.

Same interface and ring semantics as the reference's ``rl_utils.replay_buffer.ReplayBuffer``
(/root/reference/rl_utils/replay_buffer.py:27-82 on top of ``ArrayGroup``, rl_utils/array_group.py:69-101): a ring of
``maxlen`` items whose indices are *relative to the write position* (``buffer[-1]`` is the newest item, ``buffer[-len:0]``
everything), items are nested lists / tuples of arrays, ``append`` stores one item, ``extend`` a batch with a leading
dimension, ``sample(batch_size, seq_len)`` draws uniform relative indices (optionally windows of ``seq_len``).  Here the
arrays are torch tensors that stay on the environment's device, so observations never cross to the host.

Stated intent where the reference is quirky: ``append`` always stores exactly one item (the reference infers "how many"
from the leading dimensions and misreads an item whose arrays all have the same length as a batch, rl_utils/replay_buffer.py:10-24).
"""
from __future__ import annotations

from typing import Optional

import torch


def _map(f, x, *rest):
    if isinstance(x, torch.Tensor):
        return f(x, *rest)
    return [_map(f, xi, *(r[i] for r in rest)) for i, xi in enumerate(x)]


def _as_tensors(x, device):
    if isinstance(x, (list, tuple)):
        return [_as_tensors(xi, device) for xi in x]
    return torch.as_tensor(x, device=device)


class ReplayBuffer:
    def __init__(self, maxlen: int, device=None, generator: Optional[torch.Generator] = None):
        self.maxlen = int(maxlen)
        self.device = torch.device(device) if device is not None else None
        self.generator = generator
        self.buffer = None
        self.full = False
        self.pos = 0

    @property
    def empty(self) -> bool:
        return self.buffer is None

    def __len__(self) -> int:
        return self.maxlen if self.full else self.pos

    def modulate(self, key):
        """relative index / slice -> ring positions (rl_utils/replay_buffer.py:49-53)"""
        if isinstance(key, slice):
            key = torch.arange(key.start or 0, 0 if key.stop is None else key.stop, key.step or 1, device=self.device)
        key = torch.as_tensor(key, device=self.device)
        return (key + self.pos) % self.maxlen

    def __getitem__(self, key):
        assert self.buffer is not None
        idx = self.modulate(key)
        return _map(lambda b: b[idx], self.buffer)

    def __setitem__(self, key, value):
        idx = self.modulate(key)
        value = _as_tensors(value, self.device)

        def put(b, v):
            b[idx] = v.to(b.dtype)
            return b

        _map(put, self.buffer, value)

    def array(self):
        """every stored item, oldest first"""
        if self.buffer is None:
            return torch.empty(0, device=self.device)
        return self[-len(self):0]

    def _allocate(self, item):
        if self.device is None:
            first = item
            while not isinstance(first, torch.Tensor):
                first = first[0]
            self.device = first.device
        self.buffer = _map(lambda t: torch.zeros((self.maxlen,) + tuple(t.shape), dtype=t.dtype, device=self.device), item)

    def _write(self, x, count: int):
        assert count <= self.maxlen, "more items than the buffer holds"
        self[:count] = x
        if self.pos + count >= self.maxlen:
            self.full = True
        self.pos = int((self.pos + count) % self.maxlen)

    def append(self, x):
        """one item (no leading batch dimension)"""
        x = _as_tensors(x, self.device)
        if self.buffer is None:
            self._allocate(x)
        self._write(_map(lambda t: t[None], x), 1)

    def extend(self, x):
        """a batch of items: every array has the same leading dimension"""
        x = _as_tensors(x, self.device)
        first = x
        while not isinstance(first, torch.Tensor):
            first = first[0]
        if self.buffer is None:
            self._allocate(_map(lambda t: t[0], x))
        self._write(x, int(first.shape[0]))

    def sample(self, batch_size: int, seq_len: Optional[int] = None):
        """uniform relative indices in [-len, 0), optionally windows i .. i + seq_len (rl_utils/replay_buffer.py:60-68)"""
        idx = torch.randint(-len(self), 0, (batch_size,), device=self.device, generator=self.generator)
        if seq_len is not None:
            idx = idx[:, None] + torch.arange(seq_len, device=self.device)[None]
        return self[idx]
