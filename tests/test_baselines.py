"""The measured stand-ins of BASELINE.md must stay runnable: `baseline/grpc_ps.py` re-creates the reference's data path
(host-staged gRPC parameter server, full-model pull + full-gradient push per step, ps-side Adam per push,
/root/reference/distributed_server-basic.py:80, 102-103, 112) and runs on CPU; `baseline/nccl_ps.py` needs GPUs."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(180)
def test_grpc_parameter_server_stand_in_trains_and_counts_every_push():
    pytest.importorskip("grpc")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "baseline.grpc_ps", "--workers", "2", "--steps", "30", "--warmup", "5",
                        "--port", str(port)], capture_output=True, text=True, cwd=ROOT, env=env, timeout=170)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert "stand-in" in d["impl"] and "NOT the reference" in d["impl"]      # never to be mistaken for the reference
    assert d["workers"] == 2 and d["global_step"] == 2 * (30 + 5)             # one optimizer step per push
    assert d["value"] > 0 and d["final_loss"] < 0.3
