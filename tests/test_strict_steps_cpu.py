"""`--strict_steps` on the CPU plumbing backend: with one worker and one step in flight every pull follows the
acknowledgement of the previous push (the reference's `sess.run` order, /root/reference/distributed_server-basic.py:
110-113), so the native train loop is comparable step by step with a plain PyTorch re-implementation fed by the same
`next_batch` sequence. This is also the CPU twin of tests/test_zz_gpu_epoch_feed.py (same reference logic)."""
import pytest
import torch

from dist_mnist_b200.models import mlp
from dist_mnist_b200.parallel.config import EngineConfig, OptimizerConfig
from dist_mnist_b200.session import InProcessCluster
from dist_mnist_b200.utils import data


@pytest.mark.parametrize("okind,lr,tol", [("sgd", 0.05, 1e-6), ("adam", 1e-3, 2e-4)])
def test_strict_native_loop_is_lock_step_with_the_fp32_reference(okind, lr, tol):
    from bench_tools.gpu_e2e import ref_step

    ds = data.synthetic_mnist(1024, seed=9)            # 32 steps per epoch
    spec = mlp.book_model(100)
    params = mlp.init_params(spec, seed=3)
    opt = OptimizerConfig(okind, lr)
    cfg = EngineConfig(backend="cpu", lanes=1, nslots=4, strict_steps=True)
    ref_p = {k: t.clone() for k, t in params.items()}
    ref_m = {k: torch.zeros_like(t) for k, t in params.items()}
    ref_v = {k: torch.zeros_like(t) for k, t in params.items()}
    with InProcessCluster(spec, opt, cfg, batch_size=32, num_ps=1, params=params) as cl:
        w = cl.worker
        loader = w.make_loader(ds.images, ds.labels, seed=7)
        twin = w.make_loader(ds.images, ds.labels, seed=7, epoch_feed=False)   # same sequence
        outs = []
        for n in (5, 20, 50, 45):                      # 120 steps = 3.75 epochs
            outs += list(w.run_steps(n, loader))
        t = 0
        for o in outs:
            x, y = twin.next_batch()
            lref, t = ref_step(spec, ref_p, ref_m, ref_v, t, x, y, opt)
            assert abs(o.loss - lref) <= tol * abs(lref) + 1e-9, (o.seq, o.loss, lref)
            assert o.global_step == o.seq              # every earlier push had been applied when the step ran
        assert loader.epochs == twin.epochs == 3
        got = w.read_variables()
    for k in got:
        assert float((got[k] - ref_p[k]).norm() / (ref_p[k].norm() + 1e-12)) < 1e-3, k


def test_row_split_over_three_ps_tasks_is_lock_step_and_checkpoints(tmp_path):
    """`--sharding row_split` on the CPU plumbing backend (the fused tiling's layout: the hidden weight is split along
    its input features over every ps task, a push goes to three shards): strict lock-step parity with the fp32
    reference, `global_step` == pushes, checkpoint -> fresh cluster -> identical variables."""
    from bench_tools.gpu_e2e import ref_step
    from dist_mnist_b200.utils import ckpt

    ds = data.synthetic_mnist(1024, seed=9)
    spec = mlp.book_model(100)
    params = mlp.init_params(spec, seed=3)
    opt = OptimizerConfig("adam", 1e-3)
    cfg = EngineConfig(backend="cpu", lanes=1, nslots=4, strict_steps=True, sharding="row_split")
    ref_p = {k: t.clone() for k, t in params.items()}
    ref_m = {k: torch.zeros_like(t) for k, t in params.items()}
    ref_v = {k: torch.zeros_like(t) for k, t in params.items()}
    with InProcessCluster(spec, opt, cfg, batch_size=32, num_ps=3, params=params) as cl:
        w = cl.worker
        assert w.engine == "fused" and all(s.n_items > 0 for s in w.layout.shards)   # every shard owns slices
        loader = w.make_loader(ds.images, ds.labels, seed=7)
        twin = w.make_loader(ds.images, ds.labels, seed=7)
        t = 0
        for o in w.run_steps(60, loader):
            x, y = twin.next_batch()
            lref, t = ref_step(spec, ref_p, ref_m, ref_v, t, x, y, opt)
            assert abs(o.loss - lref) <= 2e-4 * abs(lref), (o.seq, o.loss, lref)
        assert w.read_global_step() == 60
        ckpt.save_checkpoint(w, str(tmp_path))
        saved = w.read_variables()
    with InProcessCluster(spec, opt, cfg, batch_size=32, num_ps=3, restore_dir=str(tmp_path)) as cl2:
        got = cl2.worker.read_variables()
        assert cl2.worker.read_global_step() == 60
        for k in saved:
            assert torch.equal(saved[k], got[k]), k
