"""Additional-input variants: the day-of-year and longitude scripts only change the number of
condition channels -- n_channel = 3 (revision1/additional_inputs/gan_train_cwgangp_pixelnorm_doy.py:135) or 2
(..._lon.py:136): the generator Dense grows to 100 + nd*nd*n_channel inputs and the critic's first conv to
n_channel + 1 input channels (..._doy.py:303-314).  Same kernels, `Context(nd, ncond)`; parity against the oracle.
"""
import numpy as np
import pytest
import torch

import rdg_oracle as O
from rdg_b200 import weights as W

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


def _rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


@pytest.fixture(scope="module", params=[2, 3])
def nets(request):
    from rdg_b200.engine import Context, Critic, Generator
    ncond = request.param
    ctx = Context(16, ncond, max_chunk=256)
    gw = W.randomize_biases(W.init_generator_weights(10 + ncond, 16, ncond))
    cw = W.randomize_biases(W.init_critic_weights(20 + ncond, 16, ncond))
    yield Generator(gw, ctx=ctx), Critic(cw, ctx=ctx), gw, cw, ncond
    ctx.close()


def _inputs(B, ncond, seed):
    rng = np.random.default_rng(seed)
    cond = np.empty((B, 16, 16, ncond), np.float32)
    cond[..., 0] = np.clip(rng.gamma(0.8, 12.0, size=(B, 16, 16)), 0, 200) / 127.4       # daily sum, normalised
    cond[..., 1:] = rng.random((B, 1, 1, ncond - 1))                                       # doy / lon fields: constant maps
    z = rng.standard_normal((B, 100)).astype(np.float32)
    return z, cond


def test_weight_shapes(nets):
    gen, crit, gw, cw, ncond = nets
    assert gw[0].shape == (100 + 256 * ncond, 3072) and cw[0].shape == (3, 3, 3, ncond + 1, 64)


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("fp16", 1e-2)])
def test_generator_matches_oracle(nets, mode, tol):
    gen, _, gw, _, ncond = nets
    z, cond = _inputs(9, ncond, 77)
    ref = O.generator_forward(gw, z, cond, torch.float64)
    out = gen.predict([z, cond], mode=mode)
    assert out.shape == (9, 24, 16, 16, 1)
    assert _rel(out.astype(np.float64), ref) <= tol
    assert np.max(np.abs(out.sum(axis=1) - 1.0)) <= 1e-5


def test_mm_output_uses_the_precipitation_channel(nets):
    gen, _, gw, _, ncond = nets
    z, cond = _inputs(6, ncond, 5)
    out = gen.forward_device(gen.ctx.dev(z), gen.ctx.dev(cond), scen_per_cond=1, mode="fp16", out_mm=True).cpu().numpy()
    want = cond[..., 0] * np.float32(127.4)
    assert np.max(np.abs(out.sum(axis=1) - want) / want) <= 1e-5


def test_critic_forward_matches_oracle(nets):
    _, crit, _, cw, ncond = nets
    rng = np.random.default_rng(3)
    z, cond = _inputs(7, ncond, 9)
    x = rng.random((7, 24, 16, 16, 1)).astype(np.float32)
    x /= x.sum(axis=1, keepdims=True)
    ref = O.critic_forward(cw, x, cond, None, torch.float64)
    out = crit.predict([x, cond])
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-6)
    masks = [(rng.random(s) < 0.75).astype(np.float32) for s in O.critic_mask_shapes(16, 7)]
    ref_m = O.critic_forward(cw, x, cond, masks, torch.float64)
    np.testing.assert_allclose(crit.predict([x, cond], masks), ref_m, rtol=2e-5, atol=2e-6)


def test_training_steps_match_oracle(nets):
    """Critic step (3 critic passes + gradient penalty) and generator step with extra condition channels."""
    from rdg_b200.engine import GanTrainer
    gen, crit, gw, cw, ncond = nets
    B = 2
    rng = np.random.default_rng(41)
    z, cond = _inputs(B, ncond, 13)
    x = rng.random((B, 24, 16, 16, 1)).astype(np.float32)
    x /= x.sum(axis=1, keepdims=True)
    alpha = rng.random((B, 1, 1, 1, 1)).astype(np.float32)
    masks3 = [[(rng.random(s) < 0.75).astype(np.float32) for s in O.critic_mask_shapes(16, B)] for _ in range(3)]
    ref_losses, ref_grads, _ = O.critic_step(gw, cw, x, cond, z, alpha, masks3, torch.float64)
    tr = GanTrainer(gen, crit, gen_mode="fp32")
    losses = tr.critic_grads(x, cond, z, alpha.reshape(-1), masks3).cpu().numpy()
    np.testing.assert_allclose(losses, ref_losses, rtol=5e-5, atol=2e-6)
    g = tr.grad_tensor(1).cpu().numpy()
    off = 0
    for i, (rg, shp) in enumerate(zip(ref_grads, W.critic_shapes(16, ncond))):
        n = int(np.prod(shp))
        # loose bound: a LeakyReLU sign flip on a near-zero pre-activation moves gradients by ~1e-3 (see test_gpu_critic_train.py)
        assert _rel_l2(g[off:off + n].reshape(shp), rg) <= 1e-2, f"critic grad tensor {i}"
        off += (n + 3) // 4 * 4
    ref_loss, ref_gg = O.generator_step(gw, cw, z, cond, masks3[0], torch.float64)
    loss = float(tr.generator_grads(z, cond, masks3[0]).item())
    assert abs(loss - ref_loss) <= 5e-5 * max(1.0, abs(ref_loss))
    gg = tr.grad_tensor(0).cpu().numpy()
    off = 0
    for i, (rg, shp) in enumerate(zip(ref_gg, W.generator_shapes(16, ncond))):
        n = int(np.prod(shp))
        if i < 9:
            assert _rel_l2(gg[off:off + n].reshape(shp), rg) <= 1e-2, f"generator grad tensor {i}"
        off += (n + 3) // 4 * 4


@pytest.mark.parametrize("extra,nch", [("doy", 3), ("lon", 2)])
def test_dropin_training_module_variants(extra, nch):
    """revision1/additional_inputs/gan_train_cwgangp_pixelnorm_{doy,lon}.py through the drop-in module: the extra condition
    channels (sin / cos of the day of year, :175-185; normalised x index, ..._lon.py:176-185) are appended to the daily-sum
    condition, on the host path and on the device-sampler path alike, and both training steps run."""
    import gan_train_cwgangp_pixelnorm as m
    try:
        m.setup(seed=2, extra=extra, device_sampler=False)
        assert m.n_channel == nch and m.generator.ncond == nch
        np.random.seed(5)
        rb, rc = next(m.generate_real_samples(4))
        rl, rcc = m.generate_latent_points(3)
        assert rc.shape == (4, 16, 16, nch) and rcc.shape == (3, 16, 16, nch)
        # literal restatement of the reference's channel construction for the first sample
        np.random.seed(5)
        ixs = np.random.randint(m.n_samples, size=4)
        t, y, x = m.indices_all[ixs[0]]
        if extra == "doy":
            d = m.timelist_all[t]
            assert np.allclose(rc[0, :, :, 1], np.sin(2 * np.pi * d / 365)) and np.allclose(rc[0, :, :, 2], np.cos(2 * np.pi * d / 365))
        else:
            assert np.allclose(rc[0, :, :, 1], (x - m.min_lonidx) / m.max_lonidx)
        m.setup(seed=2, extra=extra, device_sampler=True)
        np.random.seed(5)
        db, dc = next(m.generate_real_samples(4))
        dl, dcc = m.generate_latent_points(3)
        assert np.array_equal(db.cpu().numpy(), rb)
        np.testing.assert_allclose(dc.cpu().numpy(), rc.astype(np.float32), rtol=0, atol=0)
        np.testing.assert_allclose(dcc.cpu().numpy(), rcc.astype(np.float32), rtol=0, atol=0)
        losses = m.critic_model.train_on_batch([db, dc, np.random.normal(size=(4, 100))])
        g_loss = m.generator_model.train_on_batch([dl, dcc], None)
        assert len(losses) == 4 and np.isfinite(losses).all() and np.isfinite(g_loss)
    finally:
        m.setup(seed=0, extra=None, device_sampler=False)
