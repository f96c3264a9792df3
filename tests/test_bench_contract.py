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
