"""Tensor-core training mode (train_mode="tf32": tcgen05 kind::tf32 implicit GEMMs, csrc/tcg_gemm.cu + train.cu) against the FP64
oracle on the cases of test_gpu_critic_train.py, the device-resident step calls, and the CUDA-graph replay of a whole iteration.
Bound: gradients <= 1e-2 relative L2 per tensor (tf32 operands, 10-bit mantissa, FP32 accumulation; measured
~1e-3), losses <= 2e-3."""
import numpy as np
import pytest
import torch

import rdg_oracle as O
from rdg_b200 import weights as W

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-2


def _rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.fixture(scope="module")
def nets(ctx16):
    from rdg_b200.engine import Critic, Generator
    gw = W.randomize_biases(W.init_generator_weights(0))
    cw = W.randomize_biases(W.init_critic_weights(1), seed=9)
    return Generator(gw, ctx=ctx16), Critic(cw, ctx=ctx16), gw, cw


def _batch(B, seed=3, nd=16):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, 24, nd, nd, 1)) * 2
    x = np.exp(x - x.max(axis=1, keepdims=True)); x = (x / x.sum(axis=1, keepdims=True)).astype(np.float32)
    cond = (np.clip(rng.gamma(0.8, 12.0, size=(B, nd, nd, 1)), 0, 200) / 127.4).astype(np.float32)
    z = rng.standard_normal((B, 100)).astype(np.float32)
    alpha = rng.random((B, 1, 1, 1, 1)).astype(np.float32)
    return x, cond, z, alpha, rng


def _masks(rng, B):
    return [(rng.random(s) < 0.75).astype(np.float32) for s in O.critic_mask_shapes(16, B)]


def _split(flat, shapes):
    out, off = [], 0
    for shp in shapes:
        n = int(np.prod(shp))
        out.append(flat[off:off + n].reshape(shp))
        off += (n + 3) // 4 * 4
    return out


@pytest.mark.parametrize("use_masks", [False, True])
def test_critic_step_tc(nets, use_masks):
    """Config #3 in the tensor-core mode: 4 losses and all 10 critic gradient tensors vs autograd with create_graph."""
    from rdg_b200.engine import GanTrainer
    gen, crit, gw, cw = nets
    gen.set_weights(gw); crit.set_weights(cw)
    B = 32
    x, cond, z, alpha, rng = _batch(B, seed=11)
    masks3 = [_masks(rng, B) for _ in range(3)] if use_masks else None
    ref_losses, ref_grads, _ = O.critic_step(gw, cw, x, cond, z, alpha, masks3, torch.float64)
    tr = GanTrainer(gen, crit, gen_mode="fp32", train_mode="tf32")
    losses = tr.critic_grads(x, cond, z, alpha.reshape(-1), masks3).cpu().numpy()
    np.testing.assert_allclose(losses, ref_losses, rtol=2e-3, atol=1e-4)
    grads = _split(tr.grad_tensor(1).cpu().numpy(), W.critic_shapes(16, 1))
    for i, (mine, rg) in enumerate(zip(grads, ref_grads)):
        assert _rel_l2(mine, rg) <= GRAD_TOL, f"critic grad tensor {i} {mine.shape}: rel l2 {_rel_l2(mine, rg):.3e}"


@pytest.mark.parametrize("B", [1, 8, 32])
def test_generator_step_tc(nets, B):
    from rdg_b200.engine import GanTrainer
    gen, crit, gw, cw = nets
    gen.set_weights(gw); crit.set_weights(cw)
    rng = np.random.default_rng(40 + B)
    cond = (np.clip(rng.gamma(0.8, 12.0, size=(B, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32)
    z = rng.standard_normal((B, 100)).astype(np.float32)
    masks = _masks(rng, B)
    ref_loss, ref_grads = O.generator_step(gw, cw, z, cond, masks, torch.float64)
    tr = GanTrainer(gen, crit, train_mode="tf32")
    loss = float(tr.generator_grads(z, cond, masks).item())
    grads = _split(tr.grad_tensor(0).cpu().numpy(), W.generator_shapes(16, 1))
    assert abs(loss - ref_loss) <= 2e-3 * max(1.0, abs(ref_loss))
    for i in range(9):
        assert _rel_l2(grads[i], ref_grads[i]) <= GRAD_TOL, f"generator grad tensor {i}: rel l2 {_rel_l2(grads[i], ref_grads[i]):.3e}"
    assert abs(float(grads[9][0])) <= 1e-6      # output-conv bias: softmax over hours is shift invariant


def test_tc_mode_close_to_fp32_mode(nets):
    """Same inputs through both training modes of the library itself (no oracle): the tensor-core gradients stay within the
    tf32 bound of the FP32 SIMT ones, tensor by tensor."""
    from rdg_b200.engine import GanTrainer
    gen, crit, gw, cw = nets
    gen.set_weights(gw); crit.set_weights(cw)
    B = 16
    x, cond, z, alpha, rng = _batch(B, seed=23)
    masks3 = [_masks(rng, B) for _ in range(3)]
    out = {}
    for mode in ("fp32", "tf32"):
        tr = GanTrainer(gen, crit, gen_mode="fp32", train_mode=mode)
        l = tr.critic_grads(x, cond, z, alpha.reshape(-1), masks3).cpu().numpy()
        out[mode] = (l, _split(tr.grad_tensor(1).cpu().numpy().copy(), W.critic_shapes(16, 1)))
    np.testing.assert_allclose(out["tf32"][0], out["fp32"][0], rtol=2e-3, atol=1e-4)
    for a, b in zip(out["tf32"][1], out["fp32"][1]):
        assert _rel_l2(a, b) <= GRAD_TOL


@pytest.mark.parametrize("segmented", [False, True])
def test_device_steps_and_graph_replay(ctx16, segmented):
    """rdg_critic_step_dev / rdg_generator_step_dev / rdg_adam_apply_dev: noise, alpha and dropout masks from the device Philox
    state; a captured iteration replays to the same trajectory as the same calls issued eagerly (identical seeds and counters;
    the only difference allowed is the summation order of the filter-gradient atomics)."""
    from rdg_b200.engine import Adam, Critic, GanTrainer, Generator
    B = 8
    x, cond, _, _, rng = _batch(5 * B, seed=31)
    xr = torch.as_tensor(x.reshape(5, B, 24, 16, 16, 1), device="cuda")
    cd = torch.as_tensor(cond.reshape(5, B, 16, 16, 1), device="cuda")

    def fresh():
        g = Generator(W.init_generator_weights(7), ctx=ctx16, mode="fp16")
        c = Critic(W.init_critic_weights(8), ctx=ctx16)
        ctx16.lib.rdg_adam_reset(ctx16.handle, 0); ctx16.lib.rdg_adam_reset(ctx16.handle, 1)
        return GanTrainer(g, c, Adam(1e-4, 0.0, 0.9), gen_mode="fp16", train_mode="tf32", seed=5)

    n_iter = 3
    # eager reference trajectory
    tr = fresh()
    dl = torch.zeros((5, 4), device="cuda"); gl = torch.zeros(1, device="cuda")
    eager_losses = []
    for _ in range(n_iter):
        tr.iteration_device(xr, cd, cd[0], dl, gl)
        tr.finish()
        eager_losses.append((dl.cpu().numpy().copy(), float(gl.item())))
        if len(eager_losses) == 1:
            w_eager = [w.copy() for w in tr.generator.get_weights()] + [w.copy() for w in tr.critic.get_weights()]
    assert tr.optimizer.iterations == 6 * n_iter == tr._pull_counters()[0] and tr._pull_counters()[1] == (5 * n_iter, n_iter)
    assert all(np.isfinite(l[0]).all() and np.isfinite(l[1]) for l in eager_losses)
    # the steps really used fresh randomness each time: same data, different losses
    assert abs(eager_losses[0][0][0, 0] - eager_losses[0][0][1, 0]) > 0

    # captured (capture_iteration runs one eager pass on its own constant buffers before capturing)
    tr3 = fresh()
    ig3 = tr3.capture_iteration(B, segmented=segmented)     # segmented: the data-parallel form (NCCL between per-phase graphs)
    # rewind: fresh weights / moments / counters, then replay n_iter times on the real data
    tr3.generator.set_weights(W.init_generator_weights(7)); tr3.critic.set_weights(W.init_critic_weights(8))
    ctx16.lib.rdg_adam_reset(ctx16.handle, 0); ctx16.lib.rdg_adam_reset(ctx16.handle, 1)
    tr3.optimizer.iterations = 0
    tr3._push_counters(0, (0, 0))
    ig3.x_real.copy_(xr); ig3.cond.copy_(cd); ig3.cond_gen.copy_(cd[0])
    for it in range(n_iter):
        ig3.replay()
        torch.cuda.synchronize()
        # iteration 0 starts from identical weights: only the order of the gradient atomics differs.  From the second iteration on
        # the weights already differ where Adam stepped by lr * sign(rounding noise) (see below), so the losses drift apart
        rt, at = (2e-3, 2e-4) if it == 0 else (2e-2, 2e-3)
        np.testing.assert_allclose(ig3.d_losses.cpu().numpy(), eager_losses[it][0], rtol=rt, atol=at, err_msg=f"iteration {it}")
        assert abs(float(ig3.g_loss.item()) - eager_losses[it][1]) <= rt * max(1.0, abs(eager_losses[it][1])), f"iteration {it}"
        if it == 0:
            w_graph = tr3.generator.get_weights() + tr3.critic.get_weights()
    assert tr3.optimizer.iterations == 6 * n_iter == tr3._pull_counters()[0]
    # Weights after the FIRST iteration.  The filter gradients accumulate with FP32 atomics (summation order varies from run to
    # run, ~1e-7 relative) and beta_1 = 0 Adam steps by lr * sign(g) at first, so elements whose gradient is rounding noise step
    # the other way: two eager runs of the same iteration already differ in ~2 % of the generator's weights (tools/diag_replay.py);
    # the trajectories then separate like any chaotic training run.  Hence: one iteration, and a distribution bound.
    for i, (a, b) in enumerate(zip(w_graph, w_eager)):
        d = np.abs(a - b)
        frac = float(np.mean(d <= 1e-4 * np.abs(b) + 5e-6))
        assert frac >= 0.9 or a.size < 100, f"tensor {i} {a.shape}: only {frac:.4f} of the elements agree, max diff {d.max():.3e}"
        assert d.max() <= 1e-3, f"tensor {i} {a.shape}: max diff {d.max():.3e}"
