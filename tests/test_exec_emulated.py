"""The per-layer graph engine's native executor (csrc/executor.cu) on the emulated CUDA runtime, on CPU.

Same idea as tests/test_fexec_emulated.py: streams are real threads, copies read their source late, stream capture is
emulated (operations issued to capturing streams are recorded and replayed by a graph launch), and the step kernels are
replaced by a function that checksums the batch it finds in the slot's device buffers. The single-step submit / result
API and the native train loop (head singles, groups of U steps fed by gather threads through a ring of pinned buffers,
tail singles, StopAtStep) must give every step exactly its `next_batch` batch; built with ASan + UBSan, and once
with ThreadSanitizer."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dist_mnist_b200", "csrc")


def _build(tmp_path_factory, sanitize):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    cuda_inc = "/usr/local/cuda/include"
    exe = str(tmp_path_factory.mktemp("exec") / "exec_emulated")
    cmd = ["g++", "-O1", "-g", "-std=c++17", *sanitize, "-x", "c++",
           "-I", os.path.join(ROOT, "tests", "native", "fake_cuda"), "-I", CSRC, "-I", cuda_inc,
           os.path.join(ROOT, "tests", "native", "exec_emulated.cpp"), os.path.join(CSRC, "executor.cu"),
           os.path.join(CSRC, "loader_api.cpp"), "-o", exe, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return exe


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    return _build(tmp_path_factory, ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"])


@pytest.fixture(scope="module")
def harness_tsan(tmp_path_factory):
    return _build(tmp_path_factory, ["-fsanitize=thread"])


@pytest.mark.parametrize("nslots,lanes,graph_steps", [(2, 1, 1), (8, 4, 2), (12, 12, 4), (4, 4, 4), (16, 8, 1)])
def test_every_step_sees_its_batch(harness, nslots, lanes, graph_steps):
    env = dict(os.environ, FAKE_CUDA_DELAY_US="60", DM_GATHER_THREADS="3")
    r = subprocess.run([harness, str(nslots), str(lanes), str(graph_steps), "500"], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().startswith("OK"), (r.stdout[-1000:], r.stderr[-4000:])


def test_gather_ring_and_slot_reuse_are_race_free_under_tsan(harness_tsan):
    env = dict(os.environ, FAKE_CUDA_DELAY_US="100", DM_GATHER_THREADS="3", TSAN_OPTIONS="halt_on_error=1 exitcode=66")
    r = subprocess.run([harness_tsan, "8", "4", "2", "400"], capture_output=True, text=True, env=env, timeout=900)
    if "unexpected memory mapping" in r.stderr:      # TSan runtime vs. this kernel's ASLR settings: not our bug
        pytest.skip("ThreadSanitizer cannot run on this kernel (unexpected memory mapping)")
    assert r.returncode == 0 and "WARNING: ThreadSanitizer" not in r.stderr, (r.stdout[-1000:], r.stderr[-4000:])
