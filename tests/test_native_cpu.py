"""The native library on a CPU-only box: ABI mirrors, shm segments, flags, the native loader."""
import ctypes as C

import torch

from dist_mnist_b200 import _native as N
from dist_mnist_b200.parallel.peer_mem import Carver, Segment


def test_library_loads_and_abi_matches():
    lib = N.lib()
    for name, cls in N._MIRRORS.items():
        assert lib.dm_sizeof(name.encode()) == C.sizeof(cls), name
    assert lib.dm_sizeof(b"nope") == -1


def test_shm_segment_roundtrip_and_flags():
    c = Carver()
    c.add("a", 1024)
    c.add("flag", 4)
    seg = Segment.create("shm", c.total, table=c.table(), tag="test")
    try:
        peer = Segment.open(seg.export())
        t = seg.tensor("a", torch.float32)
        t[:] = torch.arange(256, dtype=torch.float32)
        assert torch.equal(peer.tensor("a", torch.float32), torch.arange(256, dtype=torch.float32))
        lib = N.lib()
        lib.dm_store_release_u32(seg.addr("flag"), 7)
        assert lib.dm_load_acquire_u32(peer.addr("flag")) == 7
        assert lib.dm_atomic_add_u32(peer.addr("flag"), 3) == 7
        assert lib.dm_wait_ge_u32(seg.addr("flag"), 10, 1.0) == 0
        assert lib.dm_wait_ge_u32(seg.addr("flag"), 11, 0.05) == 1
        peer.close()
    finally:
        seg.close()


def test_native_loader_matches_next_batch_semantics():
    lib = N.lib()
    n, pix, classes, batch = 10, 6, 3, 4
    x = torch.arange(n * pix, dtype=torch.float32).view(n, pix).contiguous()
    y = torch.eye(classes)[torch.arange(n) % classes].contiguous()
    h = lib.dm_loader_create(x.data_ptr(), y.data_ptr(), n, pix * 4, classes * 4, 8 * 4, classes * 4, batch, 5, 1)
    xb = torch.zeros(batch, 8)
    yb = torch.zeros(batch, classes)
    rows = []
    for _ in range(5):
        lib.dm_loader_next(h, xb.data_ptr(), yb.data_ptr())
        assert float(xb[:, pix:].abs().max()) == 0.0       # row padding untouched
        rows += [int(v) // pix for v in xb[:, 0].tolist()]
        for r in range(batch):
            idx = int(xb[r, 0]) // pix
            assert torch.equal(yb[r], y[idx])
    assert sorted(rows) == sorted(list(range(n)) * 2)       # two full epochs, each sample once per epoch
    assert sorted(rows[:n]) == list(range(n))
    assert lib.dm_loader_epochs(h) == 1
    lib.dm_loader_destroy(h)


def test_native_loader_epoch_feed_wiring(monkeypatch):
    """NativeLoader's epoch feed (fused engine on cuda) with the pinned allocations mocked out: the epoch buffers must
    be handed to the native loader in the right order — buffer 0 holds the rows exactly as next_batch hands them out —
    and a box where pinned memory cannot be allocated falls back to the row-gather path instead of failing."""
    import types

    import torch

    from dist_mnist_b200 import _native as N
    from dist_mnist_b200.parallel.worker import NativeLoader
    from dist_mnist_b200.utils import data

    ds = data.synthetic_mnist(2048, seed=1)
    w = types.SimpleNamespace(cfg=types.SimpleNamespace(backend="cuda"), engine="fused", batch=32, ld_in=784,
                              tdtype=torch.float32, lib=N.lib(), task_index=0)
    real_empty = torch.empty

    def no_pinned(*a, pin_memory=False, **k):
        if pin_memory:
            raise RuntimeError("no pinned memory on this box")
        return real_empty(*a, **k)

    monkeypatch.setattr(torch, "empty", no_pinned)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self, raising=False)
    fallback = NativeLoader(w, ds.images, ds.labels, seed=3)
    assert not fallback.epoch_feed
    monkeypatch.setattr(torch, "empty", lambda *a, pin_memory=False, **k: real_empty(*a, **k))
    fed = NativeLoader(w, ds.images, ds.labels, seed=3)
    assert fed.epoch_feed and N.lib().dm_loader_feed_enabled(fed.handle) == 1
    for step in range(70):      # 64 steps per epoch: the first epoch is a sequence of slices of buffer 0
        x, y = fed.next_batch()
        xf, yf = fallback.next_batch()
        assert torch.equal(x, xf) and torch.equal(y, yf)
        if step < 64:
            assert torch.equal(fed._feed_bufs[0][32 * step: 32 * step + 32], x)
            assert torch.equal(fed._feed_bufs[2][32 * step: 32 * step + 32], y)
    # small datasets, other engines and an explicit opt-out keep the gather path
    assert not NativeLoader(w, ds.images[:512], ds.labels[:512], seed=3).epoch_feed
    assert not NativeLoader(w, ds.images, ds.labels, seed=3, epoch_feed=False).epoch_feed
    w.engine = "graph"
    assert not NativeLoader(w, ds.images, ds.labels, seed=3).epoch_feed
    fed.close()
    fallback.close()
