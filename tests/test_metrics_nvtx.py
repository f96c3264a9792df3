"""NVTX tracing hooks: Python ranges are no-ops without CUDA, the native runtime compiles its ranges in."""
import os
import subprocess

from dist_mnist_b200 import _native as N
from dist_mnist_b200.utils import metrics


def test_nvtx_range_and_decorator_are_transparent_without_cuda():
    calls = []

    @metrics.nvtx_annotate("dm.test.fn")
    def fn(a, b=2):
        """doc"""
        calls.append((a, b))
        return a + b

    with metrics.nvtx_range("dm.test.outer"):
        assert fn(1, b=3) == 4
    assert calls == [(1, 3)] and fn.__name__ == "fn" and fn.__doc__ == "doc"


def test_native_library_carries_nvtx_ranges():
    so = str(N.lib_path())
    out = subprocess.run(["strings", so], capture_output=True, text=True).stdout
    # the header-only NVTX v3 shim (looked up lazily through NVTX_INJECTION64_PATH) and our range names
    assert "NVTX_INJECTION64_PATH" in out
    for name in ("dm.fexec.run", "dm.fexec.chunk.launch", "dm.exec.run", "dm.loader.enable_feed"):
        assert name in out, name
    assert os.path.exists(so)
