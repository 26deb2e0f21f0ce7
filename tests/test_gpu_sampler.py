"""Device-side batch sampler against the numpy statements of generate_real_samples /
generate_latent_points (gan_train_cwgangp_pixelnorm.py:143-193): bit-exact (FP32 gather, ordered hour sum, IEEE division)."""
import numpy as np
import pytest

import rdg_oracle as O

pytestmark = pytest.mark.gpu


def _radar(days=9, ny=40, nx=52, seed=0):
    rng = np.random.default_rng(seed)
    base = rng.gamma(0.8, 12.0, size=(days, 1, ny, nx)).astype(np.float32)
    prof = rng.standard_normal((days, 24, ny, nx)).astype(np.float32) * 2
    prof = np.exp(prof - prof.max(axis=1, keepdims=True)); prof /= prof.sum(axis=1, keepdims=True)
    return (base * prof).astype(np.float32)


@pytest.mark.parametrize("nd", [16, 32])
def test_gather_bit_exact(ctx16, nd):
    from rdg_b200.sampler import DeviceSampler
    data = _radar()
    idx = np.array([(t, y, x) for t in range(9) for y in range(0, 40 - nd + 1, 8) for x in range(0, 52 - nd + 1, 4)])
    s = DeviceSampler(ctx16, data, idx, nd)
    ixs = np.random.default_rng(3).integers(0, len(idx), size=37)
    batch, cond = s.gather(ixs)
    rb, rc = O.sample_windows(data, idx[ixs], nd)
    assert batch.shape == (37, 24, nd, nd, 1) and cond.shape == (37, nd, nd, 1)
    assert np.array_equal(batch.cpu().numpy(), rb) and np.array_equal(cond.cpu().numpy(), rc)
    assert abs(float(batch.sum(dim=1).max()) - 1) <= 1e-6
    _, cond2 = s.gather(ixs, with_batch=False)
    assert np.array_equal(cond2.cpu().numpy(), rc)
    assert s.gather(ixs[:0])[1].shape == (0, nd, nd, 1)                    # empty batch


def test_dry_pixel_and_bad_index_are_reported(ctx16):
    from rdg_b200.sampler import DeviceSampler
    data = _radar()
    data[2, :, 5, 7] = 0.0                                                 # daily sum 0 -> 0/0, the reference's assert fires (:167)
    idx = np.array([(2, 0, 0), (1, 0, 0), (8, 30, 40)])
    s = DeviceSampler(ctx16, data, idx, 16)
    with pytest.raises(AssertionError):
        s.gather(np.array([0]))
    s.gather(np.array([1]))
    with pytest.raises(IndexError):
        s.gather(np.array([2]))                                            # y + 16 > 40


def test_same_random_stream_as_the_numpy_sampler(ctx16):
    """np.random.seed reproduces the reference's batches: same draws in the same order (:147, :179-181)."""
    import gan_train_cwgangp_pixelnorm as m
    m.setup(seed=4, device_sampler=False)
    np.random.seed(11)
    rb, rc = next(m.generate_real_samples(5))
    rl, rcc = m.generate_latent_points(6)
    m.setup(seed=4, device_sampler=True)
    np.random.seed(11)
    db, dc = next(m.generate_real_samples(5))
    dl, dcc = m.generate_latent_points(6)
    assert np.array_equal(db.cpu().numpy(), rb) and np.array_equal(dc.cpu().numpy(), rc)
    assert np.array_equal(dl, rl) and np.array_equal(dcc.cpu().numpy(), rcc)
    # and the training step takes the device batches as they are
    loss = m.critic_model.train_on_batch([db, dc, np.random.normal(size=(5, 100))])
    assert len(loss) == 4 and np.isfinite(loss).all()
    fake, c = m.generate_fake_samples(3)
    assert fake.shape == (3, 24, 16, 16, 1)
    m.setup(seed=4, device_sampler=False)
