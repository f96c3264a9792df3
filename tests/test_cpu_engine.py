"""CPU plumbing backend: the full ps/worker protocol (arena, mailbox, flags, acks, global_step, shutdown)
over POSIX shm with the native PS serve loop — BASELINE.json config 1 in-process."""
import tempfile
import time

import pytest
import torch

from dist_mnist_b200.models import mlp
from dist_mnist_b200.parallel.config import EngineConfig, OptimizerConfig
from dist_mnist_b200.session import InProcessCluster, train_loop
from dist_mnist_b200.utils import ckpt, data


def _ref_step(spec, params, m, v, t, x, y, opt):
    loss, grads, _ = mlp.manual_loss_and_grads(spec, params, x, y)
    t += 1
    for k in params:
        g = grads[k]
        if opt.kind == "sgd":
            params[k] = params[k] - opt.lr * g
        else:
            m[k] = opt.beta1 * m[k] + (1 - opt.beta1) * g
            v[k] = opt.beta2 * v[k] + (1 - opt.beta2) * g * g
            lr_t = opt.lr * (1 - opt.beta2 ** t) ** 0.5 / (1 - opt.beta1 ** t)
            params[k] = params[k] - lr_t * m[k] / (v[k].sqrt() + opt.eps)
    return float(loss), t


@pytest.mark.parametrize("okind,nps,model", [("adam", 1, "book"), ("sgd", 1, "book"), ("adam", 2, "book"),
                                             ("adam", 2, "zhihu")])
def test_lockstep_trajectory_matches_reference_math(okind, nps, model):
    spec = mlp.get_model(model)
    opt = OptimizerConfig(okind, 1e-3 if okind == "adam" else 5e-2)
    cfg = EngineConfig(backend="cpu")
    ds = data.synthetic_mnist(512, seed=2)
    it = data.BatchIterator(ds, seed=0)
    params = mlp.init_params(spec, seed=5)
    rp = {k: v.clone() for k, v in params.items()}
    rm = {k: torch.zeros_like(v) for k, v in params.items()}
    rv = {k: torch.zeros_like(v) for k, v in params.items()}
    t = 0
    with InProcessCluster(spec, opt, cfg, batch_size=32, num_ps=nps, params=params) as cl:
        w = cl.worker
        for i in range(8):
            x, y = it.next_batch(32)
            r = w.step(x, y)
            w.wait_applied()
            lref, t = _ref_step(spec, rp, rm, rv, t, x, y, opt)
            assert r.loss == pytest.approx(lref, rel=1e-4)
            assert r.seq == i + 1
        assert w.read_global_step() == 8          # global_step == number of applied pushes (DS:91,103)
        got = w.read_variables()
        for k in got:
            assert torch.allclose(got[k], rp[k], rtol=1e-4, atol=1e-6), k
        if okind == "adam":
            gm = w.read_variables("adam_m")
            for k in gm:
                assert torch.allclose(gm[k], rm[k], rtol=1e-4, atol=1e-7), k


def test_merged_apply_mode_and_nslots1():
    spec = mlp.book_model(32)
    opt = OptimizerConfig("sgd", 1e-2)
    cfg = EngineConfig(backend="cpu", apply_mode="merged", nslots=1)
    ds = data.synthetic_mnist(256, seed=0)
    with InProcessCluster(spec, opt, cfg, batch_size=16) as cl:
        r = train_loop(cl.worker, ds, train_steps=60, log_every=0, chunk=20, print_fn=lambda s: None)
        assert r.last_global_step == 60 and r.steps_run == 60


def test_train_loop_stdout_contract_and_stop(capsys):
    spec = mlp.book_model(100)
    opt = OptimizerConfig("adam", 1e-3)
    cfg = EngineConfig(backend="cpu")
    ds = data.synthetic_mnist(1000, seed=0)
    with InProcessCluster(spec, opt, cfg, batch_size=32) as cl:
        res = train_loop(cl.worker, ds, train_steps=250, chunk=25)
        loss0, _ = None, None
    out = capsys.readouterr().out.splitlines()
    steps = [int(l.split(",")[0].split()[-1]) for l in out if l.startswith("Train step")]
    assert steps == [100, 200]                           # "Train step {}, loss: {}" every 100 global steps (DS:115-116)
    assert all(l.startswith("Train step ") and ", loss: " in l for l in out if "Train" in l)
    assert res.last_global_step == 250 and res.steps_run == 250   # StopAtStepHook on the shared counter (DS:101)
    assert res.last_loss < 0.05


def test_checkpoint_resume_roundtrip():
    spec = mlp.book_model(64)
    opt = OptimizerConfig("adam", 1e-3)
    cfg = EngineConfig(backend="cpu")
    ds = data.synthetic_mnist(512, seed=0)
    d = tempfile.mkdtemp()
    with InProcessCluster(spec, opt, cfg, batch_size=32, num_ps=2) as cl:
        r = train_loop(cl.worker, ds, train_steps=40, log_every=0, checkpoint_dir=d, chunk=10)
        assert r.checkpoint_path and ckpt.latest_checkpoint(d) == r.checkpoint_path
        p1, m1 = cl.worker.read_variables(), cl.worker.read_variables("adam_m")
    with InProcessCluster(spec, opt, cfg, batch_size=32, num_ps=2, restore_dir=d) as cl:
        assert cl.worker.read_global_step() == 40
        p2, m2 = cl.worker.read_variables(), cl.worker.read_variables("adam_m")
        for k in p1:
            assert torch.equal(p1[k], p2[k]) and torch.equal(m1[k], m2[k])
        r = train_loop(cl.worker, ds, train_steps=60, log_every=0, checkpoint_dir=d, chunk=10)
        assert r.steps_run == 20 and r.last_global_step == 60     # resumes from the restored global step
    assert ckpt.latest_checkpoint(tempfile.mkdtemp()) is None


def test_evaluate_and_accuracy_improves():
    spec = mlp.book_model(100)
    opt = OptimizerConfig("adam", 1e-3)
    cfg = EngineConfig(backend="cpu")
    ds = data.synthetic_mnist(1000, seed=4)
    with InProcessCluster(spec, opt, cfg, batch_size=32) as cl:
        l0, a0 = cl.worker.evaluate(ds.images[:300], ds.labels[:300])
        train_loop(cl.worker, ds, train_steps=150, log_every=0, chunk=50)
        cl.worker.wait_applied()
        l1, a1 = cl.worker.evaluate(ds.images[:300], ds.labels[:300])
    assert l1 < l0 and a1 > max(a0, 0.9)


def test_config_validation():
    with pytest.raises(ValueError):
        EngineConfig(push_mode="atomic").validate(OptimizerConfig("adam"))
    with pytest.raises(ValueError):
        EngineConfig(backend="cpu", push_mode="atomic").validate(OptimizerConfig("sgd"))
    with pytest.raises(ValueError):
        EngineConfig(dtype="fp8").validate(OptimizerConfig("adam"))
    EngineConfig(backend="cuda", push_mode="atomic").validate(OptimizerConfig("sgd"))


def test_train_metrics_writer_jsonl(tmp_path, capsys):
    """StepCounterHook / SummarySaverHook analogue: one JSON line per `every_steps` local steps."""
    import json
    from dist_mnist_b200.utils.metrics import TrainMetricsWriter
    spec = mlp.book_model(64)
    ds = data.synthetic_mnist(512, seed=0)
    path = tmp_path / "metrics.jsonl"
    with InProcessCluster(spec, OptimizerConfig("adam", 1e-3), EngineConfig(backend="cpu"), batch_size=32) as cl:
        m = TrainMetricsWriter(str(path), worker_index=0, batch_size=32, every_steps=20, echo=True)
        train_loop(cl.worker, ds, train_steps=100, log_every=0, chunk=10, metrics=m)
        m.close()
    recs = [json.loads(l) for l in path.read_text().splitlines()]
    assert [r["local_steps"] for r in recs] == [20, 40, 60, 80, 100]
    assert recs[-1]["global_step"] == 100 and all(r["steps_per_sec"] > 0 for r in recs)
    assert recs[-1]["loss"] < recs[0]["loss"] and 0.0 <= recs[0]["batch_accuracy"] <= 1.0
    assert recs[-1]["batch_accuracy"] > recs[0]["batch_accuracy"]
    assert sum(l.startswith("INFO global_step/sec:") for l in capsys.readouterr().out.splitlines()) == 5
