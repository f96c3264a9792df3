"""Model definition parity: init statistics, the loss quirk, hand-derived gradients, optimizer rule."""
import math

import pytest
import torch

from dist_mnist_b200.models import mlp


def test_book_init_statistics():
    p = mlp.init_params(mlp.book_model(100), seed=0)
    assert p["hid_w"].shape == (100, 784) and p["sm_w"].shape == (10, 100)   # [out, in] layout
    assert float(p["hid_b"].abs().max()) == 0.0 and float(p["sm_b"].abs().max()) == 0.0   # DS:43, DS:47
    # truncated normal: |z| <= 2 sigma, std ~= 0.88 sigma
    for name, sigma in (("hid_w", 1 / 28), ("sm_w", 1 / math.sqrt(100))):   # DS:41-42, DS:45-46
        w = p[name]
        assert float(w.abs().max()) <= 2 * sigma + 1e-7
        assert float(w.std()) == pytest.approx(0.8796 * sigma, rel=0.08)
    assert torch.equal(p["hid_w"], mlp.init_params(mlp.book_model(100), seed=0)["hid_w"])


def test_book_loss_is_cross_entropy_over_all_elements():
    torch.manual_seed(0)
    logits = torch.randn(32, 10)
    y = torch.zeros(32, 10)
    y[torch.arange(32), torch.randint(0, 10, (32,))] = 1
    book = mlp.loss_from_logits(mlp.book_model(), logits, y)
    ce = torch.nn.functional.cross_entropy(logits, y.argmax(-1))
    assert float(book) == pytest.approx(float(ce) / 10, rel=1e-5)    # mean over B x 10 (DS:53)
    xent = mlp.loss_from_logits(mlp.zhihu_model(), logits, y)
    assert float(xent) == pytest.approx(float(ce), rel=1e-5)         # DS:35


def test_clip_passes_no_gradient_below_1e_10():
    logits = torch.tensor([[60.0, 0.0, 0.0]])
    y = torch.tensor([[0.0, 1.0, 0.0]])
    spec = mlp.MLPSpec(hidden=(4,), num_classes=3)
    l = mlp.loss_from_logits(spec, logits, y)
    assert float(l) == pytest.approx(-math.log(1e-10) / 3, rel=1e-4)


@pytest.mark.parametrize("spec", [mlp.book_model(100), mlp.book_model(33), mlp.zhihu_model(), mlp.wide_model()])
def test_manual_backward_matches_autograd(spec):
    torch.manual_seed(1)
    p = mlp.init_params(spec, seed=3)
    x = torch.rand(16, 784)
    y = torch.zeros(16, 10)
    y[torch.arange(16), torch.randint(0, 10, (16,))] = 1
    l1, g1, _ = mlp.loss_and_grads(spec, p, x, y)
    l2, g2, _ = mlp.manual_loss_and_grads(spec, p, x, y)
    assert float(l1) == pytest.approx(float(l2), rel=1e-6)
    for k in g1:
        assert torch.allclose(g1[k], g2[k], rtol=1e-4, atol=1e-7), k


def test_model_catalogue():
    assert mlp.book_model().variable_names() == [("hid_w", "hid_b"), ("sm_w", "sm_b")]
    assert mlp.zhihu_model().layer_sizes == [(784, 500), (500, 500), (500, 10)]
    assert mlp.zhihu_model().num_params == 648010
    assert mlp.wide_model().num_params == 1863690
    with pytest.raises(ValueError):
        mlp.get_model("resnet")
