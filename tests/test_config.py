"""EngineConfig validation: steps in flight (lanes), steps per graph launch, mailbox slots, executor ring."""
import pytest

from dist_mnist_b200.parallel.config import MAX_SLOTS, EngineConfig, OptimizerConfig

ADAM = OptimizerConfig("adam", 1e-4)
SGD = OptimizerConfig("sgd", 1e-2)


def test_defaults_are_the_reference_like_sequential_worker():
    cfg = EngineConfig(backend="cpu")
    cfg.validate(ADAM)
    assert cfg.lanes == 1 and cfg.graph_steps == 1 and cfg.nslots == 2 and cfg.push_mode == "mailbox"


@pytest.mark.parametrize("lanes,gsteps,nslots,ring", [(2, 1, 2, 4), (4, 2, 4, 8), (8, 4, 8, 16), (16, 4, 16, 32), (4, 4, 4, 4)])
def test_valid_pipelines(lanes, gsteps, nslots, ring):
    EngineConfig(lanes=lanes, graph_steps=gsteps, nslots=nslots, pipeline_slots=ring).validate(ADAM)


@pytest.mark.parametrize("kw,msg", [
    (dict(lanes=4, nslots=2, pipeline_slots=8), "nslots"),                    # more steps in flight than mailbox slots
    (dict(lanes=4, graph_steps=3, nslots=4, pipeline_slots=12), "graph_steps"),
    (dict(lanes=4, graph_steps=2, nslots=4, pipeline_slots=6), "pipeline_slots"),   # 3 groups over 2 run streams
    (dict(lanes=8, graph_steps=4, nslots=8, pipeline_slots=4), "pipeline_slots"),   # ring shallower than lanes
    (dict(lanes=2, nslots=2, pipeline_slots=2 * MAX_SLOTS), "pipeline_slots"),
    (dict(lanes=0), "lanes"),
    (dict(nslots=0), "nslots"),
])
def test_invalid_pipelines(kw, msg):
    with pytest.raises(ValueError, match=msg):
        EngineConfig(**kw).validate(ADAM)


def test_atomic_push_needs_sgd_fp32_cuda_and_has_no_mailbox_limit():
    with pytest.raises(ValueError, match="sgd"):
        EngineConfig(push_mode="atomic").validate(ADAM)
    with pytest.raises(ValueError, match="fp32"):
        EngineConfig(push_mode="atomic", dtype="bf16").validate(SGD)
    with pytest.raises(ValueError, match="cuda"):
        EngineConfig(push_mode="atomic", backend="cpu").validate(SGD)
    EngineConfig(push_mode="atomic", lanes=8, graph_steps=4, nslots=2, pipeline_slots=16).validate(SGD)
