"""Known-answer tests pinning the CPU oracle's Keras/TF semantics (the reference has no tests,
golden vectors or runnable TensorFlow here, so the oracle is pinned by hand-computable cases)."""
import os

import numpy as np
import pytest
import torch

import rdg_oracle as O
from rdg_b200 import weights as W

GOLD = os.path.join(os.path.dirname(__file__), "golden", "generator_critic_nd16.npz")


def test_pixelnorm_known_vector():
    x = torch.tensor([[3.0, 4.0, 0.0, 0.0]])
    y = O.pixel_norm(x).numpy()
    np.testing.assert_allclose(y, [[3 / 2.5, 4 / 2.5, 0, 0]], rtol=1e-6)     # sqrt(mean(9,16,0,0)) = 2.5
    z = O.pixel_norm(torch.zeros(1, 8)).numpy()
    assert np.all(z == 0) and np.isfinite(z).all()                            # epsilon inside the sqrt


def test_tf_same_pad_table():
    """stride-2 k=3 'same' pads asymmetrically for even inputs."""
    table = {11: (6, 1, 1), 7: (4, 1, 1), 6: (3, 0, 1), 4: (2, 0, 1), 3: (2, 1, 1), 2: (1, 0, 1),
             31: (16, 1, 1), 16: (8, 0, 1), 8: (4, 0, 1)}
    for i, want in table.items():
        assert O.tf_same_pad(i, 3, 2) == want
        assert W.tf_same_pad(i, 3, 2) == want
    for i in (4, 8, 16, 24):
        assert O.tf_same_pad(i, 3, 1) == (i, 1, 1)
    assert [W.tf_valid_out(i, 3, 2) for i in (24, 16, 64)] == [11, 7, 31]


def test_critic_geometry_chain():
    g16 = W.critic_geometry(16)
    assert [g[1] for g in g16] == [(11, 7, 7), (6, 4, 4), (3, 2, 2), (2, 1, 1)]
    assert [g[2] for g in g16] == [(0, 0, 0), (1, 1, 1), (0, 0, 0), (1, 0, 0)]
    g64 = W.critic_geometry(64)
    assert [g[1] for g in g64] == [(11, 31, 31), (6, 16, 16), (3, 8, 8), (2, 4, 4)]
    assert W.critic_shapes(16)[8] == (512, 1) and W.critic_shapes(64)[8] == (8192, 1)


def test_parameter_counts():
    """3 974 273 generator / 2 880 065 critic parameters; 209 168 513 for the 64x64 generator."""
    assert sum(int(np.prod(s)) for s in W.generator_shapes(16)) == 3_974_273
    assert sum(int(np.prod(s)) for s in W.critic_shapes(16)) == 2_880_065
    assert sum(int(np.prod(s)) for s in W.generator_shapes(64)) == 209_168_513


def test_asymmetric_same_padding_matters():
    """TF 'same' (0,1) differs from the symmetric PyTorch padding=1 on an even input."""
    rng = np.random.default_rng(0)
    x = torch.as_tensor(rng.standard_normal((1, 4, 4, 4, 2)))
    k = torch.as_tensor(rng.standard_normal((3, 3, 3, 2, 3)))
    y = O._conv3d_keras(x, k, None, 2, "same")
    assert y.shape == (1, 2, 2, 2, 3)
    # hand computation of output (0,0,0): window starts at input index 0 (no leading pad)
    want = torch.einsum("thwc,thwco->o", x[0, :3, :3, :3], k)
    np.testing.assert_allclose(y[0, 0, 0, 0].numpy(), want.numpy(), rtol=1e-12)
    # last output along each axis sees one trailing zero
    want_last = torch.einsum("thwc,thwco->o", x[0, 2:, 2:, 2:], k[:2, :2, :2])
    np.testing.assert_allclose(y[0, 1, 1, 1].numpy(), want_last.numpy(), rtol=1e-12)


def test_dense_reshape_index_map():
    """Dense input order [latent(100), cond row-major]; Dense output j -> (t,h,w,c) = unravel(j,(3,2,2,256)) (A1)."""
    gw = [np.zeros(s, np.float32) for s in W.generator_shapes(16)]
    gw[0][100 + 5 * 16 + 7, ((1 * 2 + 0) * 2 + 1) * 256 + 9] = 1.0      # cond[y=5,x=7] -> (t=1,h=0,w=1,c=9)
    cond = np.zeros((1, 16, 16, 1), np.float32); cond[0, 5, 7, 0] = 2.0
    w = [torch.as_tensor(a, dtype=torch.float64) for a in gw]
    x = torch.cat([torch.zeros(1, 100, dtype=torch.float64), torch.as_tensor(cond, dtype=torch.float64).reshape(1, -1)], 1)
    d = O.lrelu(x @ w[0] + w[1]).reshape(1, 3, 2, 2, 256)
    assert float(d[0, 1, 0, 1, 9]) == 2.0 and float(d.abs().sum()) == 2.0


def test_softmax_over_hours_and_conservation():
    gw = W.randomize_biases(W.init_generator_weights(0))
    rng = np.random.default_rng(1)
    cond = rng.gamma(0.8, 12.0, size=(2, 16, 16, 1)).astype(np.float32) / 127.4
    z = rng.standard_normal((2, 100)).astype(np.float32)
    out = O.generator_forward(gw, z, cond, torch.float32)
    assert out.shape == (2, 24, 16, 16, 1)
    assert np.max(np.abs(out.sum(axis=1) - 1)) <= 1e-6         # the invariant visible in the reference's shipped CSVs
    assert out.min() > 0


def test_fold_equivalence():
    """Upsample(2) + 3^3 'same' conv == 8 phase-specific 2^3 convs on the low-res grid (A5)."""
    gw = W.randomize_biases(W.init_generator_weights(2))
    rng = np.random.default_rng(2)
    cond = rng.gamma(0.8, 12.0, size=(2, 16, 16, 1)).astype(np.float32) / 127.4
    z = rng.standard_normal((2, 100)).astype(np.float32)
    a = O.generator_forward(gw, z, cond, torch.float64)
    b = O.generator_forward_folded(gw, z, cond, torch.float64)
    assert np.max(np.abs(a - b) / np.abs(a)) <= 1e-12


def test_generate_scenarios_contract():
    gw = W.init_generator_weights(0)
    cond = 10 * np.ones((16, 16, 1))
    np.random.seed(354)
    out = O.generate_scenarios(gw, cond, 3)
    assert out.shape == (3, 24, 16, 16) and out.dtype == np.float64
    np.testing.assert_allclose(out.sum(axis=1), 10.0, rtol=1e-5)
    np.random.seed(354)
    assert O.generate_scenarios(gw, cond, 1).shape == (24, 16, 16)       # squeeze quirk, reference :62
    cond[0, 0, 0] = 0
    np.random.seed(1)
    assert np.all(O.generate_scenarios(gw, cond, 2)[:, :, 0, 0] == 0)


def test_gradient_penalty_matches_finite_difference():
    cw = W.randomize_biases(W.init_critic_weights(1))
    rng = np.random.default_rng(3)
    x = rng.random((1, 24, 16, 16, 1)); cond = rng.random((1, 16, 16, 1))
    w = [torch.as_tensor(a, dtype=torch.float64) for a in cw]
    xt = torch.as_tensor(x).requires_grad_(True)
    d = O._critic_graph(w, xt, torch.as_tensor(cond))
    (g,) = torch.autograd.grad(d.sum(), xt)
    v = rng.standard_normal(x.shape); v /= np.linalg.norm(v)
    eps = 1e-6
    f = lambda xx: float(O._critic_graph(w, torch.as_tensor(xx), torch.as_tensor(cond)).sum())
    fd = (f(x + eps * v) - f(x - eps * v)) / (2 * eps)
    assert abs(fd - float((g.numpy() * v).sum())) <= 1e-6 * max(1, abs(fd))


def test_adam_first_steps_by_hand():
    """Keras Adam, beta1=0: m = g; v1 = 0.1 g^2; lr_1 = lr*sqrt(0.1); theta -= lr_1 * g/(sqrt(v1)+1e-7)."""
    p, g = [np.array([1.0], np.float32)], [np.array([0.5], np.float32)]
    p1, v1, _ = O.adam_update(p, g, [np.zeros(1, np.float32)], 1)
    want = 1.0 - 1e-4 * np.sqrt(0.1) * 0.5 / (np.sqrt(0.1 * 0.25) + 1e-7)
    assert abs(float(p1[0][0]) - want) < 1e-7 and abs(float(v1[0][0]) - 0.025) < 1e-9
    p2, v2, _ = O.adam_update(p1, g, v1, 2)
    v_want = 0.9 * 0.025 + 0.1 * 0.25
    want2 = want - 1e-4 * np.sqrt(1 - 0.81) * 0.5 / (np.sqrt(v_want) + 1e-7)
    assert abs(float(p2[0][0]) - want2) < 1e-7


def test_golden_vectors_reproduce():
    g = np.load(GOLD)
    gw = W.randomize_biases(W.init_generator_weights(0))
    cw = W.randomize_biases(W.init_critic_weights(1), seed=9)
    frac = O.generator_forward(gw, g["latent"], g["cond"], torch.float64)
    np.testing.assert_allclose(frac, g["fractions"], rtol=1e-10)
    losses, grads, _ = O.critic_step(gw, cw, g["x_real"], g["cond"], g["latent"], g["alpha"], None, torch.float64)
    np.testing.assert_allclose(losses, g["critic_losses"], rtol=1e-9)
    np.testing.assert_allclose([np.linalg.norm(x) for x in grads], g["critic_grad_norms"], rtol=1e-8)


def test_fp32_oracle_close_to_fp64():
    gw = W.randomize_biases(W.init_generator_weights(0))
    g = np.load(GOLD)
    a = O.generator_forward(gw, g["latent"], g["cond"], torch.float32)
    assert np.max(np.abs(a - g["fractions"]) / g["fractions"]) <= 2e-5


def test_independent_numpy_restatement_agrees():
    """oracle/rdg_oracle_np.py (plain numpy, per-tap loops, hand-written backward; shares no code with rdg_oracle.py, which uses
    torch conv3d / autograd) gives the same generator fields, critic score and gradient-penalty input gradient to <= 1e-12 in
    float64 -- with dropout masks, non-zero biases and a second conditioning channel (ncond = 2) in one of the cases."""
    import rdg_oracle_np as ON
    from rdg_b200 import weights as W
    for nd, ncond, seed in ((16, 1, 11), (16, 2, 12)):
        rng = np.random.default_rng(seed)
        gw = [w.astype(np.float64) for w in W.randomize_biases(W.init_generator_weights(seed, nd, ncond))]
        cw = [w.astype(np.float64) for w in W.randomize_biases(W.init_critic_weights(seed + 1, nd, ncond))]
        z = rng.standard_normal((1, 100))
        cond = rng.gamma(0.8, 12.0, size=(1, nd, nd, ncond)) / 127.4
        f_t = O.generator_forward(gw, z, cond, torch.float64)
        f_n = ON.generator_one(gw, z[0], cond[0])
        assert f_t.shape == (1, 24, nd, nd, 1)
        assert np.abs(f_t[0] - f_n).max() <= 1e-12
        np.testing.assert_allclose(f_n.sum(axis=0), 1.0, rtol=0, atol=1e-12)
        masks = [(rng.random(s) < 0.75).astype(np.float64) for s in O.critic_mask_shapes(nd, 1)]
        s_t = O.critic_forward(cw, f_t, cond, masks, torch.float64)
        s_n, g_n = ON.critic_one(cw, f_n, cond[0], [m[0] for m in masks], want_input_grad=True)
        assert abs(float(s_t[0, 0]) - s_n) <= 1e-12 * max(1.0, abs(s_n))
        # gradient penalty input gradient: torch autograd (through critic_step's extras) vs the hand-written backward
        x_real = rng.random((1, 24, nd, nd, 1)); x_real /= x_real.sum(axis=1, keepdims=True)
        alpha = rng.random((1, 1, 1, 1, 1))
        m3 = [[np.ones_like(m) for m in masks], [np.ones_like(m) for m in masks], masks]
        losses, _, ex = O.critic_step(gw, cw, x_real, cond, z, alpha, m3, torch.float64)
        xhat = alpha[0] * x_real[0] + (1 - alpha[0]) * f_n
        _, g_hat = ON.critic_one(cw, xhat, cond[0], [m[0] for m in masks], want_input_grad=True)
        assert np.abs(ex["xhat_grad"][0] - g_hat).max() <= 1e-12 * max(1.0, np.abs(g_hat).max())
        gp_n = ON.gradient_penalty_term(cw, xhat, cond[0], [m[0] for m in masks])
        assert abs(losses[3] - gp_n) <= 1e-12 * max(1.0, gp_n)
    # the public call: mm/h fields reproduce the daily sum
    cond_mm = np.random.default_rng(3).gamma(0.8, 12.0, size=(16, 16, 1))
    out = ON.generate_scenarios_one(gw[:0] + [w.astype(np.float64) for w in W.init_generator_weights(4)], cond_mm, np.random.default_rng(5).standard_normal(100))
    np.testing.assert_allclose(out.sum(axis=0), cond_mm[..., 0], rtol=1e-12, atol=1e-12)
