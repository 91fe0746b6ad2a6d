"""Host-side logic that needs no GPU: the pinned ring arena of attention_shift (regions are reused only after their
transfer's event, regions still held by the host are stepped over)."""
import types

import torch

from attentionshift_b200 import attention_shift as AS


class _Ring(AS._PinnedRing):
    def __init__(self, n):
        self.buf = torch.empty(n, dtype=torch.uint8)          # pageable stand-in: only the bookkeeping is under test
        self.size, self.head, self.inflight = n, 0, []
        self.waits = 0

    def done(self, token):
        if token is not None:
            token[2] = types.SimpleNamespace(synchronize=lambda: setattr(self, 'waits', self.waits + 1))

    @staticmethod
    def _fresh(n):
        raise AssertionError('ring fell back to a fresh allocation')


def test_ring_reuse_and_held_regions():
    r = _Ring(4096)
    held = None
    for i in range(300):
        v, t = r.take(200 + (37 * i) % 700)
        assert v.numel() % 256 == 0 and v.numel() >= 200
        v[:] = i % 251
        if i == 5:
            held = v                    # never marked done: the ring must step over it on every lap
        else:
            r.done(t)
        if held is not None:
            assert (held == 5).all()
    assert r.waits > 0                  # wrapped around and waited for the events of the regions it took back
    assert len(r.inflight) < 32


def test_ring_oversize_request_bypasses_the_arena():
    class R(_Ring):
        @staticmethod
        def _fresh(n):
            return torch.empty(n, dtype=torch.uint8)
    r = R(1024)
    v, t = r.take(5000)
    assert t is None and v.numel() >= 5000 and r.head == 0


def test_packed_upload_slices_and_shapes(monkeypatch):
    """_upload_i32: several host arrays travel as ONE buffer and come back as int32 views of the right shapes (the fp32
    bit-pattern trick used for coordinates included)."""
    import numpy as np
    ring = _Ring(1 << 16)
    monkeypatch.setattr(AS, '_ring', lambda dev: ring)
    coords = np.array([[1.5, -2.25], [3.0, 4.0]], dtype=np.float32)
    a, b, c, d = AS._upload_i32([[1, 2, 3], np.arange(6).reshape(2, 3), coords.view(np.int32), []], torch.device('cpu'))
    assert a.tolist() == [1, 2, 3] and b.shape == (2, 3) and b.tolist() == [[0, 1, 2], [3, 4, 5]]
    assert torch.equal(c.view(torch.float32), torch.from_numpy(coords)) and d.numel() == 0
    assert all(t.dtype == torch.int32 for t in (a, b, c, d))
    assert len(ring.inflight) == 1 and ring.inflight[0][2] is not None      # one region, marked done after the copy
