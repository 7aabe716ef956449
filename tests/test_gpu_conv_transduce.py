"""GPU parity of ConvTransduce1D (criterions/transducer.py:370-556) against fixtures produced
by the reference's own module on the oracle shim (tests/golden/make_golden.py gen_conv):
outputs, input gradients and kernel-weight gradients in forward-score and viterbi modes, and
the reference's shape tests (tests/transducer_test.py:57-96)."""
import numpy as np
import pytest
import torch

import _golden as G

pytestmark = pytest.mark.gpu


def _layer(z, name):
    from gtn_applications_b200.criterions.transducer import ConvTransduce1D
    ks, stride, blank, opt, learn, vit, spike = (int(v) for v in z[name + "_config"])
    lexicon = G.unpack(z["lexicon"], z["lexicon_offsets"])
    layer = ConvTransduce1D(lexicon, ks, stride, blank, blank_optional=bool(opt), learn_params=bool(learn),
                            scale=str(z[name + "_scale"]), normalize=str(z[name + "_normalize"]),
                            viterbi=bool(vit), spike=bool(spike))
    if learn:
        layer.kernel_params.data = torch.tensor(z[name + "_params"])
    return layer.cuda()


@pytest.mark.parametrize("name", ["fwd", "learn", "viterbi", "forced_spike"])
def test_fixtures_from_reference(name):
    z = G.load("conv")
    layer = _layer(z, name)
    x = torch.tensor(z[name + "_inputs"], device="cuda", requires_grad=True)
    y = layer(x)
    y.backward(torch.tensor(z[name + "_grad_outputs"], device="cuda"))
    np.testing.assert_allclose(y.detach().cpu().numpy(), z[name + "_outputs"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), z[name + "_grad_inputs"], rtol=1e-3, atol=1e-5)
    if layer.kernel_params is not None:
        np.testing.assert_allclose(layer.kernel_params.grad.cpu().numpy(), z[name + "_grad_params"],
                                   rtol=1e-3, atol=1e-5)


def test_shapes_and_errors_like_the_reference():
    """tests/transducer_test.py:57-96"""
    from gtn_applications_b200.criterions.transducer import ConvTransduce1D
    lexicon = [(0, 0), (0, 1), (1, 0), (1, 1)]
    conv = ConvTransduce1D(lexicon, 5, 3, 2)
    B, C = 2, 3
    with pytest.raises(ValueError):
        conv(torch.randn(B, 0, C, device="cuda"))
    for tin in (1, 2, 3, 4):
        conv(torch.randn(B, tin, C, device="cuda"))
    for ti, to in zip((1, 3, 4, 6, 7, 8), (1, 1, 2, 2, 3, 3)):
        x = torch.randn(B, ti, C, device="cuda", requires_grad=True)
        y = conv(x)
        assert y.shape == (B, to, len(lexicon))
        y.backward(torch.ones_like(y))
        assert x.grad.shape == x.shape and torch.isfinite(x.grad).all()
    with pytest.raises(ValueError):
        ConvTransduce1D([(0, 0, 0)], 3, 1, 2)          # kernel too small for the repeats
    with pytest.raises(ValueError):
        ConvTransduce1D(lexicon, 5, 3, 2, scale="cubic")
