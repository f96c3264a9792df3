"""Drop-in entry point with the reference's file name and flags:

    python distributed_server-basic.py --job_name ps     --task_index 0 --ps_hosts 127.0.0.1:9910 --worker_hosts 127.0.0.1:9900,127.0.0.1:9901
    python distributed_server-basic.py --job_name worker --task_index 0 --ps_hosts 127.0.0.1:9910 --worker_hosts 127.0.0.1:9900,127.0.0.1:9901
    python distributed_server-basic.py --job_name worker --task_index 1 --ps_hosts 127.0.0.1:9910 --worker_hosts 127.0.0.1:9900,127.0.0.1:9901

All logic lives in `dist_mnist_b200.cli` (the engine is not TensorFlow).
"""
import sys

from dist_mnist_b200.cli import main

if __name__ == "__main__":
    sys.exit(main())
