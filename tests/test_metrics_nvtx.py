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


def test_clock_csv_summary_reports_median_clock_and_throttle_reasons():
    """The `clocks` object of the bench JSON line: median SM clock under load, max clock, and every throttle reason that
    was active in any sample (hw_slowdown / thermal reasons reject a run; sw_power_cap is kept and noted)."""
    csv_text = "\n".join([
        "0, 1965, 1965, 412.31, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
        "0, 1965, 1965, 640.02, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
        "0, 1920, 1965, 998.70, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
        "1, 1965, 1965, 120.00, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
        "0, [N/A], [N/A], [N/A], [N/A], [N/A], [N/A], [N/A], [N/A]",     # a sample taken while the driver was busy
        "garbage line",
    ])
    s = metrics.summarize_clock_csv(csv_text)
    assert s["samples"] == 4 and s["sm_mhz"] == 1965 and s["sm_max_mhz"] == 1965
    assert s["reasons"] == ["sw_power_cap"] and abs(s["power_w_max"] - 998.70) < 1e-6
    hot = metrics.summarize_clock_csv("0, 1200, 1965, 700.0, 0x8, Active, Active, Not Active, Not Active\n")
    assert hot["reasons"] == ["hw_slowdown", "hw_thermal_slowdown"] and hot["sm_mhz"] == 1200
    assert metrics.summarize_clock_csv("")["samples"] == 0
    # without nvidia-smi the sampler is a no-op that still returns the same shape
    sampler = metrics.ClockSampler(interval_ms=10)
    sampler.start()
    out = sampler.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples"}
