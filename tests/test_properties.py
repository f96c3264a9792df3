"""Property tests (hypothesis): arena layout / item tiling for arbitrary MLP shapes, and the native next_batch
loader against an independent model of TF's `DataSet.next_batch` for arbitrary dataset / batch sizes."""
import torch
from hypothesis import example, given, settings, strategies as st

from dist_mnist_b200 import _native as N
from dist_mnist_b200.models.mlp import MLPSpec
from dist_mnist_b200.parallel import sharding


@settings(max_examples=40, deadline=None)
@given(in_features=st.integers(1, 300), hidden=st.lists(st.integers(1, 300), min_size=1, max_size=3),
       classes=st.integers(2, 16), num_ps=st.integers(1, 4), strategy=st.sampled_from(["round_robin", "byte_balanced"]),
       tile_n=st.sampled_from([32, 64]))
def test_items_tile_any_model_exactly_once(in_features, hidden, classes, num_ps, strategy, tile_n):
    spec = MLPSpec(name="prop", in_features=in_features, hidden=tuple(hidden), num_classes=classes, loss="xent",
                   init="glorot")
    lay = sharding.build_layout(spec, num_ps, strategy, dw_tile_n=tile_n)
    assert set(lay.placement.values()) <= set(range(num_ps))
    seen = 0
    for sh in lay.shards:
        cover = torch.zeros(sh.arena_elems, dtype=torch.int32)
        for it in sh.items:
            assert it.rows >= 1 and 1 <= it.cols <= max(sharding.TILE_M, tile_n)
            for r in range(it.rows):
                cover[it.offset + r * it.ld: it.offset + r * it.ld + it.cols] += 1
        assert int(cover.max()) <= 1                      # no element belongs to two items
        n_var_elems = 0
        for vl in sh.variables:
            assert vl.offset % sharding.ALIGN_ELEMS == 0  # TMA / float4 alignment of every variable
            n_var_elems += vl.spec.numel
            # flag indices of a variable are contiguous and inside the shard's item table
            assert 0 <= vl.item_base and vl.item_base + vl.n_items <= sh.n_items
        assert int(cover.sum()) == n_var_elems            # ... and every parameter belongs to exactly one
        seen += len(sh.variables)
    assert seen == 2 * (len(hidden) + 1)                  # one weight + one bias per layer, each placed once


@settings(max_examples=30, deadline=None)
@given(n=st.integers(1, 40), batch=st.integers(1, 17), n_batches=st.integers(1, 12), seed=st.integers(0, 2 ** 31),
       shuffle=st.booleans())
def test_native_loader_visits_every_sample_once_per_epoch(n, batch, n_batches, seed, shuffle):
    """TF next_batch semantics: consecutive batches walk through a per-epoch permutation; a batch that crosses the
    epoch boundary is completed from the next epoch; without shuffling the order is the dataset order."""
    lib = N.lib()
    pix, classes = 3, 2
    x = torch.arange(n * pix, dtype=torch.float32).view(n, pix).contiguous()
    y = torch.zeros(n, classes)
    y[torch.arange(n), torch.arange(n) % classes] = 1
    h = lib.dm_loader_create(x.data_ptr(), y.data_ptr(), n, pix * 4, classes * 4, pix * 4, classes * 4, batch, seed,
                             int(shuffle))
    xb, yb = torch.zeros(batch, pix), torch.zeros(batch, classes)
    order = []
    for _ in range(n_batches):
        lib.dm_loader_next(h, xb.data_ptr(), yb.data_ptr())
        idx = [int(v) // pix for v in xb[:, 0].tolist()]
        for r, i in enumerate(idx):
            assert torch.equal(xb[r], x[i]) and torch.equal(yb[r], y[i])
        order += idx
    total = n_batches * batch
    for e in range(0, total, n):                          # every complete epoch is a permutation of the dataset
        chunk = order[e: e + n]
        if len(chunk) == n:
            assert sorted(chunk) == list(range(n))
        else:
            assert len(set(chunk)) == len(chunk)          # the partial last epoch has no repeats
    if not shuffle:
        assert order == [i % n for i in range(total)]
    assert lib.dm_loader_epochs(h) == (total - 1) // n     # the wrap happens lazily, when the next sample is needed
    lib.dm_loader_destroy(h)


@settings(max_examples=40, deadline=None)
@given(in_chunks=st.integers(8, 32), in_tail=st.sampled_from([0, 4, 8, 16, 28]), hidden=st.integers(1, 128),
       classes=st.integers(2, 11), num_ps=st.integers(1, 8),
       strategy=st.sampled_from(["round_robin", "byte_balanced", "row_split"]), row_blocks=st.integers(1, 8))
def test_fused_tiling_any_eligible_model(in_chunks, in_tail, hidden, classes, num_ps, strategy, row_blocks):
    """Fused step engine tiling for any eligible 1-hidden-layer model: 8 K-slices (<= 4 chunks each) cover the input
    features exactly, every parameter belongs to exactly one ps item, every item's flag exists on its shard, rows of
    the hidden weight are whole 128-byte lines, and with row_split the slices are dealt round-robin over the shards."""
    in_features = (in_chunks - 1) * 32 + (in_tail if in_tail else 32)
    spec = MLPSpec(name="book", in_features=in_features, hidden=(hidden,), num_classes=classes, loss="book", init="book")
    assert sharding.fused_eligible(spec, 32)
    lay = sharding.build_layout(spec, num_ps, strategy, dw_tile_n=32, engine="fused", ps_row_blocks=row_blocks)
    sl = lay.fused_slices
    assert [s.rank for s in sl] == list(range(8))
    assert sl[0].kc_begin == 0 and all(a.kc_begin + max(a.kc_count, 0) <= b.kc_begin or a.kc_count == 0
                                       for a, b in zip(sl, sl[1:]))
    assert sum(s.kc_count for s in sl) == in_chunks and max(s.kc_count for s in sl) <= 4
    hw = lay.by_name["hid_w"]
    assert hw.ld % 32 == 0 and hw.ld >= in_features
    total = 0
    for sh in lay.shards:
        cover = torch.zeros(sh.arena_elems, dtype=torch.int32)
        for it in sh.items:
            assert 0 <= it.flag < sh.n_flags
            for r in range(it.rows):
                cover[it.offset + r * it.ld: it.offset + r * it.ld + it.cols] += 1
        assert int(cover.max()) <= 1
        total += int(cover.sum())
    assert total == spec.num_params
    if strategy == "row_split":
        assert [s.ps for s in sl] == [r % num_ps for r in range(8)]
    else:
        assert len({s.ps for s in sl}) == 1


@settings(max_examples=150, deadline=None)
@given(n=st.integers(min_value=0, max_value=5000), dst_off=st.integers(min_value=0, max_value=63),
       src_off=st.integers(min_value=0, max_value=63))
@example(n=3136, dst_off=0, src_off=0)      # an MNIST fp32 row, 32-byte aligned: the AVX2 loop
@example(n=3136, dst_off=32, src_off=5)
@example(n=3136, dst_off=16, src_off=3)     # 16-byte aligned only: the SSE2 loop
@example(n=4096, dst_off=0, src_off=1)
@example(n=256, dst_off=0, src_off=0)
@example(n=272, dst_off=48, src_off=7)
@example(n=3140, dst_off=0, src_off=0)      # size not a multiple of 16: memcpy fallback
def test_streaming_row_copy_equals_memcpy_for_any_size_and_alignment(n, dst_off, src_off):
    """loader.h copy_row_streaming (non-temporal stores with AVX2 / SSE2 fast paths and a memcpy fallback) must be a
    plain copy for every size and alignment, and must not touch a byte outside [dst, dst + n)."""
    import numpy as np
    from dist_mnist_b200 import _native as N
    lib = N.lib()
    src = np.random.default_rng(n * 4096 + dst_off * 64 + src_off).integers(0, 256, size=n + 128, dtype=np.uint8)
    dst = np.full(n + 192, 0xA5, dtype=np.uint8)
    base_d = dst.ctypes.data
    pad = (-base_d) % 64                       # start from a 64-byte aligned address, then apply the offset
    lib.dm_copy_row_streaming(base_d + pad + dst_off, src.ctypes.data + src_off, n)
    lo = pad + dst_off
    assert np.array_equal(dst[lo:lo + n], src[src_off:src_off + n])
    assert (dst[:lo] == 0xA5).all() and (dst[lo + n:] == 0xA5).all()
