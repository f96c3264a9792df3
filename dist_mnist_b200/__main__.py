"""`python -m dist_mnist_b200 --job_name ps|worker --task_index N --ps_hosts ... --worker_hosts ...` — the same command
line as the top-level `distributed_server-basic.py` shim (the reference's entry point,
/root/reference/distributed_server-basic.py:56-116)."""
import sys

from .cli import main

if __name__ == "__main__":
    sys.exit(main())
