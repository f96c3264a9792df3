"""Peer-memory segments: the data plane of the parameter server.

A *segment* is a contiguous block of memory owned by one task and mapped by its peers:

  * `cuda`  — `cudaMalloc` on the owner's GPU, exported as a CUDA IPC handle and opened by peers with
              `cudaIpcOpenMemHandle(..., cudaIpcMemLazyEnablePeerAccess)`. Peer kernels then load/store it
              directly over NVLink/NVSwitch (TMA tensor maps, ld/st.global, red.add, st.release.sys flags).
  * `shm`   — POSIX shared memory (CPU plumbing backend; same protocol, no GPU).

This replaces the reference's gRPC `RecvTensor` transport that `tf.train.Server` provides
(/root/reference/distributed_server-basic.py:80, 108, 112).

`Carver` lays out named, aligned sub-buffers inside a segment; the resulting offset table travels through the
rendezvous store so both sides agree on the layout.
"""
from __future__ import annotations

import ctypes as C
import os
import uuid
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .. import _native as N


def _align(n: int, a: int) -> int:
    return (n + a - 1) // a * a


class Carver:
    """Sequential sub-allocation of named regions (byte offsets) inside one segment."""

    def __init__(self, align: int = 256):
        self.align = align
        self.offsets: Dict[str, Tuple[int, int]] = {}
        self.size = 0

    def add(self, name: str, nbytes: int) -> int:
        off = _align(self.size, self.align)
        self.offsets[name] = (off, nbytes)
        self.size = off + nbytes
        return off

    def table(self) -> Dict[str, Tuple[int, int]]:
        return dict(self.offsets)

    @property
    def total(self) -> int:
        return _align(max(self.size, 1), self.align)


class _CudaArrayView:
    """Minimal __cuda_array_interface__ carrier so torch can wrap a raw device pointer without copying."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None,
        }
        self._owner = owner


@dataclass
class Segment:
    kind: str                 # "cuda" | "shm"
    ptr: int
    nbytes: int
    device: int = -1          # owning CUDA device (cuda) / -1
    name: str = ""            # shm name
    owner: bool = True
    table: Optional[Dict[str, Tuple[int, int]]] = None
    _closed: bool = False

    # ---- creation / export / import ---------------------------------------------------------
    @staticmethod
    def create(kind: str, nbytes: int, device: int = -1, table=None, tag: str = "seg") -> "Segment":
        lib = N.lib()
        out = C.c_void_p()
        if kind == "cuda":
            N.check(lib.dm_cuda_malloc(device, nbytes, C.byref(out)), "cudaMalloc segment")
            return Segment("cuda", out.value, nbytes, device=device, owner=True, table=table)
        if kind == "shm":
            name = f"/dmnist-{tag}-{os.getpid()}-{uuid.uuid4().hex[:8]}"
            N.check(lib.dm_shm_create(name.encode(), nbytes, C.byref(out)), "shm_create")
            return Segment("shm", out.value, nbytes, name=name, owner=True, table=table)
        raise ValueError(kind)

    def export(self) -> dict:
        d = {"kind": self.kind, "nbytes": self.nbytes, "table": self.table, "pid": os.getpid()}
        if self.kind == "cuda":
            h = (C.c_uint8 * 64)()
            N.check(N.lib().dm_ipc_get_handle(self.ptr, C.addressof(h)), "cudaIpcGetMemHandle")
            d["handle"] = bytes(h).hex()
            d["device"] = self.device
            d["ptr"] = self.ptr  # valid only inside the owning process (same-process peers use it directly)
        else:
            d["name"] = self.name
        return d

    @staticmethod
    def open(desc: dict, device: int = -1) -> "Segment":
        """Map a peer's segment. `device` = the *local* CUDA device whose kernels will access it."""
        lib = N.lib()
        table = desc.get("table")
        if table is not None:
            table = {k: tuple(v) for k, v in table.items()}
        if desc["kind"] == "cuda":
            if desc["pid"] == os.getpid():
                # same process (e.g. ps + worker sharing one launcher process): IPC handles cannot be opened by
                # their own creator — use the pointer directly and make sure P2P is enabled for cross-GPU access.
                if device >= 0 and device != desc["device"]:
                    N.check(lib.dm_enable_peer_access(device, desc["device"]), "enable peer access")
                return Segment("cuda", desc["ptr"], desc["nbytes"], device=desc["device"], owner=False, table=table,
                               name="same-process")
            out = C.c_void_p()
            h = (C.c_uint8 * 64).from_buffer_copy(bytes.fromhex(desc["handle"]))
            N.check(lib.dm_ipc_open_handle(device, C.addressof(h), C.byref(out)), "cudaIpcOpenMemHandle")
            return Segment("cuda", out.value, desc["nbytes"], device=desc["device"], owner=False, table=table)
        out = C.c_void_p()
        N.check(lib.dm_shm_open(desc["name"].encode(), desc["nbytes"], C.byref(out)), "shm_open")
        return Segment("shm", out.value, desc["nbytes"], name=desc["name"], owner=False, table=table)

    # ---- addressing -----------------------------------------------------------------------------
    def addr(self, region: str, byte_offset: int = 0) -> int:
        off, _ = self.table[region]
        return self.ptr + off + byte_offset

    def region_bytes(self, region: str) -> int:
        return self.table[region][1]

    # ---- local views (only valid for memory this process may dereference from the host / local device) ----
    def tensor(self, region: str, dtype: torch.dtype, count: Optional[int] = None) -> torch.Tensor:
        """Typed 1-D view of a region. cuda: owner-side (or same-process) view on the owning device;
        shm: host view."""
        off, nbytes = self.table[region]
        if self.kind == "shm":
            buf = (C.c_uint8 * nbytes).from_address(self.ptr + off)
            t = torch.from_numpy(np.frombuffer(buf, dtype=np.uint8)).view(dtype)
        else:
            view = _CudaArrayView(self.ptr + off, nbytes, self)
            t = torch.as_tensor(view, device=f"cuda:{self.device}").view(dtype)
        return t if count is None else t[:count]

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        lib = N.lib()
        if self.kind == "cuda":
            if self.owner:
                lib.dm_cuda_free(self.ptr)
            elif self.name != "same-process":
                lib.dm_ipc_close(self.ptr)
        else:
            lib.dm_shm_unmap(self.ptr, self.nbytes)
            if self.owner:
                lib.dm_shm_unlink(self.name.encode())
