"""The three toy scripts of the reference (ps_server-basic.py, basic/1u1m-basic.py, basic/nu1m-basic.py; SURVEY
C17-C19) have equivalents under examples/; their expected outputs are the reference's 7 / 10 / 0.4."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = dict(os.environ, PYTHONPATH=ROOT, PYTHONUNBUFFERED="1", CUDA_VISIBLE_DEVICES="")


def _run(script, *args, timeout=90):
    return subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), *args], capture_output=True,
                          text=True, timeout=timeout, env=ENV, cwd=ROOT)


def test_one_device_placement_example():
    r = _run("one_gpu_basic.py")                      # 1U:11, 24-26: device list, then w+b = 7 and w*b = 10
    assert r.returncode == 0, r.stderr
    assert "7." in r.stdout and "10." in r.stdout and r.stdout.lstrip().startswith("[")


def test_multi_device_op_placement_example():
    r = _run("n_gpu_basic.py")                        # NU:22
    assert r.returncode == 0, r.stderr
    assert "7.0" in r.stdout and "10.0" in r.stdout


@pytest.mark.timeout(120)
def test_toy_parameter_server_round_trip():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = str(s.getsockname()[1])
    ps = subprocess.Popen([sys.executable, os.path.join(ROOT, "examples", "ps_server_basic.py"), "--role", "ps",
                           "--backend", "cpu", "--port", port], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, env=ENV, cwd=ROOT)
    try:
        w = _run("ps_server_basic.py", "--role", "worker", "--iterations", "2", "--backend", "cpu", "--port", port)
        assert w.returncode == 0, w.stderr + w.stdout
        out = w.stdout
        assert "7." in out and "10." in out and "0.4" in out, out     # [w + b, w * b, w / b]  (PSB:54-56, 63)
    finally:
        try:
            ps.wait(timeout=20)
        except subprocess.TimeoutExpired:
            ps.kill()
        ps.communicate()


def test_nvls_probe_example_reports_cleanly_without_gpus():
    # no GPU here: the probe must say why it cannot run and exit non-zero instead of crashing
    r = _run("nvls_multicast_probe.py", "2")
    assert r.returncode == 1 and "PROBE FAILED" in r.stdout and "FAIL:" in r.stdout, (r.stdout, r.stderr)
