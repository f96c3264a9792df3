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
