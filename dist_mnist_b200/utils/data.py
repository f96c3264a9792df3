"""Input data: MNIST-shaped datasets and the `next_batch` iterator.

Reference parity:
  * DS:69   `mnist = read_data_sets('MNIST_data', one_hot=True)` — 55 000 train images of 784 floats in [0, 1]
            and one-hot float labels. There is no network here, so the default is a *synthetic* dataset of the
            same shapes (class-dependent prototypes + noise, so the loss actually decreases); real MNIST idx
            files are used when `--data_dir` contains them.
  * DS:111  `mnist.train.next_batch(32)` — shuffled epochs, sequential batches, an epoch boundary inside a batch
            is completed from the freshly shuffled next epoch. Every worker owns the full set and shuffles
            independently (no sharding across workers).

`BatchIterator` is the Python implementation (CPU backend, tests); the GPU worker uses the native loader in
csrc/executor.cu, which implements the same semantics over a pinned copy of the dataset.
"""
from __future__ import annotations

import gzip
import os
import struct
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

TRAIN_SIZE = 55000  # size of `mnist.train` in the TF loader (60000 minus the 5000 validation images)


@dataclass
class Dataset:
    images: torch.Tensor  # [N, pixels] float32 in [0, 1]
    labels: torch.Tensor  # [N, classes] float32 one-hot
    synthetic: bool = True

    def __len__(self) -> int:
        return self.images.shape[0]

    @property
    def num_pixels(self) -> int:
        return self.images.shape[1]

    @property
    def num_classes(self) -> int:
        return self.labels.shape[1]


def synthetic_mnist(n: int = TRAIN_SIZE, seed: int = 0, pixels: int = 784, classes: int = 10,
                    noise: float = 0.35) -> Dataset:
    """Deterministic MNIST-shaped data: sparse class prototypes ("strokes") plus uniform noise."""
    g = torch.Generator().manual_seed(seed)
    protos = (torch.rand(classes, pixels, generator=g) < 0.18).float() * (0.5 + 0.5 * torch.rand(classes, pixels, generator=g))
    y = torch.randint(0, classes, (n,), generator=g)
    x = protos[y] * (0.6 + 0.4 * torch.rand(n, 1, generator=g)) + noise * torch.rand(n, pixels, generator=g)
    x = x.clamp_(0.0, 1.0).contiguous()
    labels = torch.zeros(n, classes)
    labels[torch.arange(n), y] = 1.0
    return Dataset(x, labels, synthetic=True)


def _read_idx(path: str) -> np.ndarray:
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        _, _, dtype_code, ndim = struct.unpack(">BBBB", f.read(4))
        shape = struct.unpack(">" + "I" * ndim, f.read(4 * ndim))
        assert dtype_code == 0x08, "only ubyte idx files are supported"
        return np.frombuffer(f.read(), dtype=np.uint8).reshape(shape)


def load_mnist(data_dir: Optional[str]) -> Optional[Dataset]:
    """Real MNIST from idx files (`train-images-idx3-ubyte[.gz]`, `train-labels-idx1-ubyte[.gz]`) if present."""
    if not data_dir or not os.path.isdir(data_dir):
        return None
    for suffix in ("", ".gz"):
        img = os.path.join(data_dir, "train-images-idx3-ubyte" + suffix)
        lab = os.path.join(data_dir, "train-labels-idx1-ubyte" + suffix)
        if os.path.exists(img) and os.path.exists(lab):
            x = torch.from_numpy(_read_idx(img).reshape(-1, 784).astype(np.float32) / 255.0)[:TRAIN_SIZE]
            yi = torch.from_numpy(_read_idx(lab).astype(np.int64))[:TRAIN_SIZE]
            labels = torch.zeros(x.shape[0], 10)
            labels[torch.arange(x.shape[0]), yi] = 1.0
            return Dataset(x.contiguous(), labels, synthetic=False)
    return None


def get_dataset(data_dir: Optional[str], n: int = TRAIN_SIZE, seed: int = 0) -> Dataset:
    """`--data_dir` given: the idx files must be there (the reference always trains on real MNIST, DS:69) — a typo
    must not silently turn into training on synthetic data. No `--data_dir`: synthetic 28x28 data (no network)."""
    if data_dir is not None:
        ds = load_mnist(data_dir)
        if ds is None:
            raise FileNotFoundError(
                f"--data_dir {data_dir!r} does not contain the MNIST training idx files "
                f"(train-images-idx3-ubyte[.gz], train-labels-idx1-ubyte[.gz]); omit --data_dir for synthetic data")
        return ds
    return synthetic_mnist(n=n, seed=seed)


class BatchIterator:
    """`DataSet.next_batch` semantics (TF contrib mnist loader): shuffle at the start of every epoch."""

    def __init__(self, dataset: Dataset, seed: int = 0, shuffle: bool = True):
        self.ds = dataset
        self.shuffle = shuffle
        self.rng = np.random.default_rng(seed)
        self.n = len(dataset)
        self.perm = np.arange(self.n)
        if shuffle:
            self.rng.shuffle(self.perm)
        self.cursor = 0
        self.epochs_completed = 0

    def next_batch(self, batch_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
        idx = []
        need = batch_size
        while need > 0:
            if self.cursor == self.n:
                self.cursor = 0
                self.epochs_completed += 1
                if self.shuffle:
                    self.rng.shuffle(self.perm)
            take = min(need, self.n - self.cursor)
            idx.append(self.perm[self.cursor:self.cursor + take].copy())  # copy: perm is reshuffled in place
            self.cursor += take
            need -= take
        sel = torch.from_numpy(np.concatenate(idx))
        return self.ds.images[sel], self.ds.labels[sel]
