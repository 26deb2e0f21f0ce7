"""GPU parity of the on-device ensemble statistics against the oracle: area means
(generate_and_evaluate.py:533-535) and properscoring.crps_ensemble + area mean (generate_and_evaluate_crps.py:189-191).
FP32 sums on the device vs float64 on the host: tolerance 2e-5 relative (+1e-6 absolute for dry pixels)."""
import numpy as np
import pytest
import torch

import rdg_oracle as O
from rdg_b200 import weights as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gen(ctx16):
    from rdg_b200.engine import Generator
    return Generator(W.randomize_biases(W.init_generator_weights(0)), ctx=ctx16)


@pytest.mark.parametrize("n_cond,spc", [(3, 10), (2, 100), (1, 1), (2, 160), (2, 333), (1, 1000), (1, 1100)])
def test_stats_kernels_match_oracle(gen, n_cond, spc):
    rng = np.random.default_rng(spc)
    fields = rng.gamma(0.8, 2.0, size=(n_cond * spc, 24, 16, 16)).astype(np.float32)
    fields[:, :, :2, :2] = 0.0                                   # dry pixels: ties at zero
    obs = rng.gamma(0.8, 2.0, size=(n_cond, 24, 16, 16)).astype(np.float32)
    obs[:, :, :2, :2] = 0.0
    out = gen.ensemble_stats_device(gen.ctx.dev(fields), spc, gen.ctx.dev(obs), want_crps_field=True)
    am = out["area_mean"].cpu().numpy()
    np.testing.assert_allclose(am, O.area_mean(fields), rtol=2e-6)
    ref = np.stack([O.crps_ensemble_energy(obs[i], fields[i * spc:(i + 1) * spc]) for i in range(n_cond)])
    np.testing.assert_allclose(out["crps"].cpu().numpy(), ref, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(out["crps_area_mean"].cpu().numpy(), ref.mean(axis=(2, 3)), rtol=2e-5, atol=1e-6)
    if spc <= 10:   # the literal sorted-CDF algorithm of properscoring (pure-Python loops: small case only)
        lit = np.stack([O.crps_ensemble(obs[i], fields[i * spc:(i + 1) * spc]) for i in range(n_cond)])
        np.testing.assert_allclose(out["crps"].cpu().numpy(), lit, rtol=2e-5, atol=1e-6)


def test_generate_stats_host_equals_generate_then_reduce(gen):
    """One call = the reference's CRPS loop body for every condition (generate_and_evaluate_crps.py:177-191)."""
    n_cond, spc = 5, 40
    rng = np.random.default_rng(8)
    cond_mm = np.clip(rng.gamma(0.8, 12.0, size=(n_cond, 16, 16, 1)), 0, 200).astype(np.float32)
    cond = cond_mm / np.float32(127.4)
    z = rng.standard_normal((n_cond * spc, 100)).astype(np.float32)
    frac = rng.random((n_cond, 24, 16, 16)).astype(np.float32)
    obs = frac / frac.sum(axis=1, keepdims=True) * cond_mm[..., 0][:, None]
    fields = gen.generate_ensemble_host(z, cond, spc, mode="fp16", out_mm=True)
    am, cr = gen.generate_ensemble_stats_host(z, cond, spc, obs, mode="fp16", out_mm=True)
    np.testing.assert_allclose(am, O.area_mean(fields), rtol=2e-6)
    ref = np.stack([O.crps_ensemble_energy(obs[i], fields[i * spc:(i + 1) * spc]).mean(axis=(1, 2)) for i in range(n_cond)])
    np.testing.assert_allclose(cr, ref, rtol=2e-5, atol=1e-6)
    # conservation carries over: the 24 area means of a scenario add up to the area mean of its daily sum
    np.testing.assert_allclose(am.sum(axis=1), np.repeat(cond_mm[..., 0].mean(axis=(1, 2)), spc), rtol=1e-5)
    am2, cr2 = gen.generate_ensemble_stats_host(z, cond, spc, None, mode="fp16")
    assert cr2 is None and np.array_equal(am, am2)
    with pytest.raises(Exception):
        gen.generate_ensemble_stats_host(z[:-1], cond, spc, obs)
