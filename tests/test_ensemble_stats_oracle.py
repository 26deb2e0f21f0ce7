"""CPU checks of the oracle's ensemble statistics.  properscoring (the reference's CRPS dependency,
generate_and_evaluate_crps.py:189) is not installed here: the restated algorithm is pinned by hand-computed cases and by
the algebraic identity between its CDF-integral form and the energy form."""
import numpy as np

import rdg_oracle as O


def test_crps_hand_cases():
    # one member: |x - y|
    assert O.crps_ensemble(np.array([2.0]), np.array([[5.0]]))[0] == 3.0
    # obs 0, members {1, 3}: int_0^1 1 dx + int_1^3 (1/2)^2 dx = 1.5
    assert abs(O.crps_ensemble(np.array([0.0]), np.array([[3.0], [1.0]]))[0] - 1.5) < 1e-15
    # obs inside the ensemble: members {0, 2}, obs 1 -> int_0^1 (1/2)^2 + int_1^2 (1/2)^2 = 0.5
    assert abs(O.crps_ensemble(np.array([1.0]), np.array([[0.0], [2.0]]))[0] - 0.5) < 1e-15
    # obs above all members: members {0, 1}, obs 3 -> int_0^1 (1/2)^2 + int_1^3 1 = 2.25
    assert abs(O.crps_ensemble(np.array([3.0]), np.array([[0.0], [1.0]]))[0] - 2.25) < 1e-15
    # ties and zeros (dry pixels): all members equal the observation
    assert O.crps_ensemble(np.zeros(4), np.zeros((7, 4))).max() == 0.0


def test_crps_cdf_form_equals_energy_form():
    rng = np.random.default_rng(0)
    f = rng.gamma(0.8, 2.0, size=(37, 5, 6))
    f[:, 0, 0] = 0.0
    o = rng.gamma(0.8, 2.0, size=(5, 6))
    a, b = O.crps_ensemble(o, f), O.crps_ensemble_energy(o, f)
    assert np.max(np.abs(a - b)) <= 1e-12


def test_area_mean():
    x = np.arange(2 * 24 * 4 * 4, dtype=np.float64).reshape(2, 24, 4, 4)
    np.testing.assert_allclose(O.area_mean(x)[1, 3], x[1, 3].mean())


def test_sample_windows_hand_case():
    """Oracle restatement of generate_real_samples' preprocessing (gan_train_cwgangp_pixelnorm.py:148-163) on a 2x2 window."""
    data = np.zeros((2, 24, 3, 3), np.float32)
    data[1, :, 1:, 1:] = np.arange(1, 25, dtype=np.float32)[:, None, None]        # hour h carries h mm at every pixel
    batch, cond = O.sample_windows(data, np.array([[1, 1, 1]]), 2, norm_scale=127.4)
    assert batch.shape == (1, 24, 2, 2, 1) and cond.shape == (1, 2, 2, 1)
    assert batch.dtype == np.float32 and cond.dtype == np.float32
    np.testing.assert_allclose(cond[0, :, :, 0], np.float32(300.0) / np.float32(127.4))       # 1 + 2 + ... + 24 = 300
    np.testing.assert_allclose(batch[0, :, 0, 0, 0], np.arange(1, 25, dtype=np.float32) / np.float32(300.0))
    assert abs(float(batch[0].sum(axis=0).max()) - 1) < 1e-6
    none, cond2 = O.sample_windows(data, np.array([[1, 1, 1]]), 2, with_batch=False)
    assert none is None and np.array_equal(cond, cond2)
