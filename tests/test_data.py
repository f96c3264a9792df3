import numpy as np
import torch

from dist_mnist_b200.utils import data


def test_synthetic_shapes_and_determinism():
    ds = data.synthetic_mnist(500, seed=1)
    assert ds.images.shape == (500, 784) and ds.labels.shape == (500, 10)
    assert float(ds.images.min()) >= 0 and float(ds.images.max()) <= 1
    assert torch.equal(ds.labels.sum(-1), torch.ones(500))
    assert torch.equal(ds.images, data.synthetic_mnist(500, seed=1).images)


def test_next_batch_epoch_semantics():
    ds = data.synthetic_mnist(10, seed=0)
    it = data.BatchIterator(ds, seed=0)
    seen = []
    for _ in range(5):
        x, y = it.next_batch(4)     # 20 samples = exactly 2 epochs
        assert x.shape == (4, 784) and y.shape == (4, 10)
        seen.append(x)
    allx = torch.cat(seen)
    # every sample appears exactly twice (an epoch boundary inside a batch continues into the next epoch)
    counts = (allx[:, None, :] == ds.images[None, :, :]).all(-1).sum(0)
    assert torch.equal(counts, torch.full((10,), 2))
    assert it.epochs_completed == 1


def test_idx_reader_roundtrip(tmp_path):
    import struct
    imgs = (np.arange(3 * 28 * 28) % 255).astype(np.uint8).reshape(3, 28, 28)
    labs = np.array([1, 2, 3], dtype=np.uint8)
    with open(tmp_path / "train-images-idx3-ubyte", "wb") as f:
        f.write(struct.pack(">BBBB", 0, 0, 8, 3) + struct.pack(">III", 3, 28, 28) + imgs.tobytes())
    with open(tmp_path / "train-labels-idx1-ubyte", "wb") as f:
        f.write(struct.pack(">BBBB", 0, 0, 8, 1) + struct.pack(">I", 3) + labs.tobytes())
    ds = data.load_mnist(str(tmp_path))
    assert ds is not None and not ds.synthetic and len(ds) == 3
    assert ds.labels.argmax(-1).tolist() == [1, 2, 3]
    assert data.load_mnist(str(tmp_path / "missing")) is None
