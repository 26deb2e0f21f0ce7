"""Large-domain variant (alternative_domains/gan_train_cwgangp_pixelnorm_largedomain.py: ndomain=64,
Dense 4196->49152, Reshape (3,8,8,256), critic Flatten 8192) -- SURVEY 8a row a18 / BASELINE config #5."""
import numpy as np
import pytest
import torch

import rdg_oracle as O
from rdg_b200 import weights as W

pytestmark = pytest.mark.gpu
ND = 64


@pytest.fixture(scope="module")
def big():
    from rdg_b200.engine import Context, Critic, Generator
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ctx = Context(ND, 1, max_chunk=8)
    gw = W.randomize_biases(W.init_generator_weights(0, ND))
    cw = W.randomize_biases(W.init_critic_weights(1, ND), seed=9)
    yield Generator(gw, ctx=ctx), Critic(cw, ctx=ctx), gw, cw
    ctx.close()


def _inputs(B, seed=64):
    rng = np.random.default_rng(seed)
    cond = (np.clip(rng.gamma(0.8, 12.0, size=(B, ND, ND, 1)), 0, 200) / 127.4).astype(np.float32)
    z = rng.standard_normal((B, 100)).astype(np.float32)
    return z, cond, rng


def test_generator_forward_all_modes(big):
    gen, _, gw, _ = big
    z, cond, _ = _inputs(3)
    ref = O.generator_forward(gw, z, cond, torch.float64)
    assert ref.shape == (3, 24, ND, ND, 1)
    for mode, tol in (("fp32", 1e-5), ("fp16", 1e-2), ("bf16", 4e-2)):
        out = gen.predict([z, cond], mode=mode)
        assert out.shape == ref.shape
        assert float(np.max(np.abs(out - ref) / np.abs(ref))) <= tol, mode
        assert float(np.max(np.abs(out.sum(axis=1) - 1.0))) <= 1e-5


def test_critic_forward_and_step(big):
    from rdg_b200.engine import GanTrainer
    gen, crit, gw, cw = big
    B = 2
    z, cond, rng = _inputs(B, seed=5)
    x = rng.standard_normal((B, 24, ND, ND, 1)) * 2
    x = np.exp(x - x.max(axis=1, keepdims=True)); x = (x / x.sum(axis=1, keepdims=True)).astype(np.float32)
    ref = O.critic_forward(cw, x, cond, None, torch.float64)
    out = crit.predict([x, cond])
    assert np.max(np.abs(out - ref)) <= 1e-5 * max(1.0, float(np.abs(ref).max()))
    alpha = rng.random((B, 1, 1, 1, 1)).astype(np.float32)
    ref_losses, ref_grads, _ = O.critic_step(gw, cw, x, cond, z, alpha, None, torch.float64)
    tr = GanTrainer(gen, crit, gen_mode="fp32")
    losses = tr.critic_grads(x, cond, z, alpha.reshape(-1), None).cpu().numpy()
    np.testing.assert_allclose(losses, ref_losses, rtol=5e-5, atol=1e-6)
    g = tr.grad_tensor(1).cpu().numpy()
    off = 0
    for i, (rg, shp) in enumerate(zip(ref_grads, W.critic_shapes(ND, 1))):
        n = int(np.prod(shp)); mine = g[off:off + n].reshape(shp); off += (n + 3) // 4 * 4
        rel = np.linalg.norm(mine - rg) / (np.linalg.norm(rg) + 1e-30)
        assert rel <= 1e-3, f"critic grad {i}: {rel:.2e}"      # loose: LeakyReLU sign flips, see test_gpu_critic_train.py


def test_dropin_training_module_large_domain(tmp_path, monkeypatch):
    """alternative_domains/gan_train_cwgangp_pixelnorm_largedomain.py through the drop-in module: setup(domain=64), one
    iteration of train() (5 critic steps + 1 generator step, :463-491) at batch 2 on synthetic radar-shaped data."""
    import gan_train_cwgangp_pixelnorm as m
    monkeypatch.chdir(tmp_path)
    try:
        m.setup(seed=1, domain=64, device_sampler=True)
        assert m.ndomain == 64 and m.generator.get_weights()[0].shape == (100 + 64 * 64, 49152)
        m.hist['d_loss'].clear(); m.hist['g_loss'].clear()
        m.train(1, 2, bat_per_epo=1, save=False)
        assert len(m.hist['g_loss']) == 1 and np.isfinite(m.hist['g_loss'][0]) and np.isfinite(m.hist['d_loss'][0])
        fake, cond = m.generate_fake_samples(2)
        assert fake.shape == (2, 24, 64, 64, 1) and np.max(np.abs(fake.sum(axis=1) - 1)) <= 1e-5
    finally:
        m.ndomain = 16
        m.setup(seed=0, extra=None, device_sampler=False)


def test_tensor_core_training_mode_large_domain(big):
    """train_mode="tf32" at ndomain = 64 (Dense 4196 -> 49152, convs on 16^2 / 32^2 / 64^2 grids, critic Flatten 8192): losses and every
    gradient tensor of a critic step and a generator step against the library's own FP32 SIMT mode on the same inputs (the FP32
    mode is pinned against the oracle above and in test_gpu_critic_train.py); bound 1e-2 relative L2 per tensor."""
    from rdg_b200.engine import GanTrainer
    gen, crit, gw, cw = big
    gen.set_weights(gw); crit.set_weights(cw)
    B = 2
    z, cond, rng = _inputs(B, seed=71)
    x = rng.random((B, 24, ND, ND, 1)).astype(np.float32); x /= x.sum(axis=1, keepdims=True)
    alpha = rng.random(B).astype(np.float32)
    masks3 = [[(rng.random(s) < 0.75).astype(np.float32) for s in O.critic_mask_shapes(ND, B)] for _ in range(3)]

    def split(flat, shapes):
        out, off = [], 0
        for shp in shapes:
            n = int(np.prod(shp)); out.append(flat[off:off + n].reshape(shp)); off += (n + 3) // 4 * 4
        return out

    res = {}
    for mode in ("fp32", "tf32"):
        tr = GanTrainer(gen, crit, gen_mode="fp32", train_mode=mode)
        lc = tr.critic_grads(x, cond, z, alpha, masks3).cpu().numpy()
        gc = split(tr.grad_tensor(1).cpu().numpy().copy(), W.critic_shapes(ND, 1))
        lg = float(tr.generator_grads(z, cond, masks3[0]).item())
        gg = split(tr.grad_tensor(0).cpu().numpy().copy(), W.generator_shapes(ND, 1))
        res[mode] = (lc, gc, lg, gg)
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b) + 1e-30))
    np.testing.assert_allclose(res["tf32"][0], res["fp32"][0], rtol=2e-3, atol=1e-4)
    assert abs(res["tf32"][2] - res["fp32"][2]) <= 2e-3 * max(1.0, abs(res["fp32"][2]))
    for i, (a, b) in enumerate(zip(res["tf32"][1], res["fp32"][1])):
        assert rel(a, b) <= 1e-2, f"critic tensor {i}: {rel(a, b):.2e}"
    for i, (a, b) in enumerate(zip(res["tf32"][3][:9], res["fp32"][3][:9])):
        assert rel(a, b) <= 1e-2, f"generator tensor {i}: {rel(a, b):.2e}"
