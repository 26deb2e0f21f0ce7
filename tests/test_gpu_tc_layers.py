"""Each tensor-core generator layer in isolation against the FP64 oracle of that layer
(UpSampling3D + Conv3D 'same' + PixelNorm + LeakyReLU, gan_train_cwgangp_pixelnorm.py:330-343).

Layer 2 at ndomain 16 runs the resident-plane kernel (gen_tc_planes.cu): sample PAIRS per CTA, so odd batches
(ragged last pair), a single sample and more pairs than SMs are the edge cases.  Inputs are pre-rounded to
fp16 so that the only error sources are the fp16 folded weights and the fp16 output rounding:
tolerance 4e-3 of the output scale (PixelNorm bounds |y| <= sqrt(C)), stated here.
"""
import numpy as np
import pytest
import torch

import rdg_oracle as O
from rdg_b200 import weights as W

pytestmark = pytest.mark.gpu

CIN = (256, 256, 128)


@pytest.fixture(scope="module")
def gen(ctx16):
    from rdg_b200.engine import Generator
    gw = W.randomize_biases(W.init_generator_weights(3))
    return Generator(gw, ctx=ctx16), gw


def _oracle_layer(gw, layer, x):
    k = torch.as_tensor(gw[2 + 2 * layer]).double()
    b = torch.as_tensor(gw[3 + 2 * layer]).double()
    y = O._conv3d_keras(O.upsample2(torch.as_tensor(x).double()), k, b, 1, "same")
    return O.lrelu(O.pixel_norm(y)).numpy()


@pytest.mark.parametrize("layer,B", [(0, 5), (1, 9), (2, 1), (2, 2), (2, 7), (2, 40), (2, 301)])
def test_layer_matches_oracle(gen, layer, B):
    g, gw = gen
    f = 1 << layer
    rng = np.random.default_rng(100 + layer * 10 + B)
    x = rng.standard_normal((B, 3 * f, 2 * f, 2 * f, CIN[layer])).astype(np.float16).astype(np.float32)
    y = g.tc_layer_device(layer, g.ctx.dev(x), mode="fp16").cpu().numpy()
    ref = _oracle_layer(gw, layer, x)
    assert y.shape == ref.shape
    err = np.abs(y - ref)
    scale = np.sqrt(np.mean(ref ** 2))
    assert np.max(err) <= 4e-3 * max(scale, 1.0) * 4, (np.max(err), scale)
    assert np.sqrt(np.mean(err ** 2)) <= 1.5e-3 * scale


def test_layer2_zero_halo_exact(gen):
    """A delta input at a corner position exercises every halo view of the resident planes: the response must be
    confined to the 3^3 upsampled neighbourhood and match the oracle (no leakage across samples of a pair or
    across line / plane boundaries of the shared-halo layout)."""
    g, gw = gen
    B = 4
    x = np.zeros((B, 12, 8, 8, 128), np.float32)
    x[0, 0, 0, 0, :] = 1.0
    x[1, 11, 7, 7, :] = -0.5
    x[2, 5, 0, 7, 3] = 2.0
    x[3, 6, 7, 0, 100] = 1.0
    zero_b = [np.zeros_like(w) if i % 2 else w for i, w in enumerate(gw)]
    from rdg_b200.engine import Generator
    g0 = Generator(zero_b, ctx=g.ctx)
    try:
        y = g0.tc_layer_device(2, g.ctx.dev(x), mode="fp16").cpu().numpy()
        ref = _oracle_layer(zero_b, 2, x)
        # with zero bias, positions outside the receptive field are exactly 0 / sqrt(1e-8) * 0 = 0
        assert np.all(y[ref == 0.0] == 0.0)
        assert np.max(np.abs(y - ref)) <= 2e-2
        nz = np.abs(ref) > 0.05
        assert np.max(np.abs(y[nz] - ref[nz]) / np.abs(ref[nz])) <= 2e-2
    finally:
        g.set_weights(gw)
