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
    from dist_mnist_b200.models import mlp
    from dist_mnist_b200.parallel.config import OptimizerConfig
    a = bench.parse_args([])
    assert a.gpus == 1 and a.steps >= 100 and a.warmup >= 3 and a.optimizer == "adam" and a.batch_size == 32
    # the configuration bench.py derives from its defaults must validate and select the fused engine for the
    # reference's live model (784-100-10, batch 32)
    cfg = bench.engine_config(a)
    cfg.validate(OptimizerConfig("adam", 1e-4))
    assert cfg.nslots >= cfg.lanes
    assert cfg.resolve_engine(mlp.get_model(a.model, a.hidden_units), a.batch_size) == "fused"
