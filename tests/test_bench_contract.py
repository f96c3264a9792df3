import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_reports_unavailable():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "5", "--warmup", "3"], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and "unavailable" in d and "TensorFlow" in d["unavailable"]


def test_graft_entry_build():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()


def test_bench_defaults_and_config_are_consistent():
    sys.path.insert(0, ROOT)
    import bench
    from dist_mnist_b200.parallel.config import EngineConfig, OptimizerConfig
    a = bench.parse_args([])
    assert a.gpus == 1 and a.steps >= 100 and a.warmup >= 3 and a.optimizer == "adam" and a.batch_size == 32
    # the defaults bench.py derives (nslots = lanes, U = min(lanes, 4), ring = 2 x lanes) must validate
    lanes = a.lanes
    EngineConfig(backend="cuda", lanes=lanes, graph_steps=a.graph_steps or min(lanes, 4), nslots=a.nslots or max(2, lanes),
                 pipeline_slots=max(4, 2 * lanes)).validate(OptimizerConfig("adam", 1e-4))
