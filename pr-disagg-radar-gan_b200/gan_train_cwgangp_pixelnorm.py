"""Drop-in for the model / training-step surface of the reference's `gan_train_cwgangp_pixelnorm.py`.

The reference is a flat script that loads cluster data at import time and trains immediately
(:81-140, :524-529).  This module keeps its names, constants, signatures and step semantics but
builds nothing at import: call `setup(...)` (optionally with your own `(days,24,ny,nx)` array and
`valid_indices`; otherwise synthetic radar-shaped data, BASELINE config #4) and then `train(...)`,
or drive `critic_model.train_on_batch` / `generator_model.train_on_batch` yourself.

Reference symbol                                   -> here
  create_generator() / create_discriminator() :272-357 -> same names, return rdg_b200 Generator / Critic
  PixelNormalization / RandomWeightedAverage / GradientPenalty / wasserstein_loss :215-270
                                                    -> thin classes/functions (the fused CUDA steps implement them)
  critic_model.train_on_batch([X_real, cond_real, latent],[valid,fake,dummy]) :472
  generator_model.train_on_batch([latent, cond], valid) :482
  generate_real_samples / generate_latent_points / generate_fake_samples / generate :143-211
  train(n_epochs, _batch_size, start_epoch=0) :431-521 (plots omitted; hist.csv + .h5 checkpoints kept)
"""
import os

import numpy as np

from rdg_b200 import hdf5 as _hdf5
from rdg_b200 import weights as _W

# ---- constants, reference :48-78
startdate, enddate = '20090101', '20161231'
ndomain = 16
stride = 16
tres = 1
nhours = 24 // tres
tp_thresh_daily = 5
n_thresh = 20
norm_scale = 127.4
n_disc = 5
GRADIENT_PENALTY_WEIGHT = 10
latent_dim = 100
batch_size = 32
n_epoch_and_batch_size_list = ((50, 32),)
n_channel = 1
name = 'wgancp_pixelnorm'
params = f'{startdate}-{enddate}-tp_thresh_daily{tp_thresh_daily}_n_thresh{n_thresh}_ndomain{ndomain}_stride{stride}'
outdir = os.environ.get('RDG_OUTDIR', 'trained_models')

data = None          # (days, 24, ny, nx) float32 mm/h, reference :120-138
indices_all = None   # (n,3) int (tidx, yidx, xidx), reference :117-119
n_samples = 0
generator = critic = critic_model = generator_model = optimizer = None
hist = {'d_loss': [], 'g_loss': []}


def wasserstein_loss(y_true, y_pred):
    """reference :215-216"""
    return np.mean(np.asarray(y_true) * np.asarray(y_pred))


class RandomWeightedAverage:
    """alpha*real + (1-alpha)*fake with alpha~U[0,1) per sample (reference :219-227); the CUDA critic
    step fuses this, the class is kept for API parity / host-side use."""

    def call(self, inputs, alpha=None, **kwargs):
        real, fake = np.asarray(inputs[0]), np.asarray(inputs[1])
        if alpha is None:
            alpha = np.random.uniform(size=(batch_size, 1, 1, 1, 1))
        return alpha * real + (1 - alpha) * fake

    __call__ = call


class GradientPenalty:
    """sqrt(sum(grad^2)) - 1 per sample (reference :230-244).  K.gradients has no host equivalent: the
    gradient w.r.t. the interpolated sample is produced inside rdg_critic_step_grads."""

    def call(self, grad):
        g = np.asarray(grad)
        return np.sqrt(np.sum(g.reshape(g.shape[0], -1) ** 2, axis=1, keepdims=True)) - 1

    __call__ = call


def _PixelNormalization():
    from raindisagg_gan_pretrained import PixelNormalization as P   # same class, reference :249-270
    return P


_CTXS = {}


def _ctx():
    """One context per (ndomain, n_channel): the additional-input and large-domain variants only change these two."""
    key = (ndomain, n_channel)
    if key not in _CTXS:
        from rdg_b200.engine import Context
        _CTXS[key] = Context(ndomain, n_channel)
    return _CTXS[key]


def create_discriminator():
    """reference :272-309 (glorot_uniform kernels, zero biases)"""
    from rdg_b200.engine import Critic
    return Critic(_W.init_critic_weights(int(np.random.randint(2 ** 31)), ndomain, n_channel), ctx=_ctx())


def create_generator():
    """reference :312-357 (RandomNormal(0.02) kernels, zero biases)"""
    from rdg_b200.engine import Generator
    return Generator(_W.init_generator_weights(int(np.random.randint(2 ** 31)), ndomain, n_channel), ctx=_ctx(),
                     mode=os.environ.get('RDG_MODE', 'fp16'))


class _CriticModel:
    """critic_model (reference :387-392): outputs [valid, fake, disc_gp], losses [wasserstein, wasserstein, mse],
    loss_weights [1, 1, 10], Adam shared with generator_model."""

    def __init__(self, trainer):
        self._t = trainer

    def train_on_batch(self, inputs, targets=None, alpha=None, masks3="draw"):
        return self._t.critic_train_on_batch(inputs, targets, alpha, masks3)

    def predict(self, inputs, alpha=None):
        """[valid, fake, disc_gp] without updating (reference :461 calls it once before training): critic scores of the real
        and the generated samples and the GradientPenalty output sqrt(sum(grad^2)) - 1 per sample (:238-241) at the
        RandomWeightedAverage of both (:221-224; alpha ~ U[0,1) per sample unless given).  predict() runs without dropout."""
        import torch
        x_real, cond, latent = inputs
        x_real = np.asarray(x_real, np.float32); cond = np.asarray(cond, np.float32)
        fake_img = generator.predict([latent, cond])
        n = x_real.shape[0]
        if alpha is None:
            alpha = np.random.uniform(size=(n, 1, 1, 1, 1))
        alpha = np.asarray(alpha, np.float32).reshape(n, 1, 1, 1, 1)
        xhat = alpha * x_real + (1 - alpha) * fake_img
        t = self._t
        # gradient of sum_b D(xhat_b) w.r.t. xhat: the generator-step backward of the critic with cotangent 1 per sample
        g = t.critic_input_gradient(xhat, cond)
        gp = np.sqrt(np.sum(g.reshape(n, -1) ** 2, axis=1, keepdims=True)) - 1
        return [critic.predict([x_real, cond]), critic.predict([fake_img, cond]), gp.astype(np.float32)]


class _GeneratorModel:
    """generator_model (reference :395-408)."""

    def __init__(self, trainer):
        self._t = trainer

    def train_on_batch(self, inputs, target=None, masks="draw"):
        return self._t.generator_train_on_batch(inputs, target, masks)


def synthetic_radar(days=64, ny=64, nx=64, seed=0):
    """Radar-shaped stand-in for the SMHI data (reference :120-138): (days,24,ny,nx) float32 mm/h with a
    skewed intensity distribution and smooth hourly structure; every day is wet enough to pass the
    reference's validity filter (compute_valid_indices.py:46-48)."""
    rng = np.random.default_rng(seed)
    base = rng.gamma(0.8, 12.0, size=(days, 1, ny, nx)).astype(np.float32)
    prof = rng.standard_normal((days, 24, ny, nx)).astype(np.float32) * 2
    prof = np.exp(prof - prof.max(axis=1, keepdims=True))
    prof /= prof.sum(axis=1, keepdims=True)
    return (base * prof).astype(np.float32)


_GRAPHS = {}         # captured training iterations of train() (tensor-core mode)
sampler = None       # rdg_b200.sampler.DeviceSampler when the batches are drawn on the GPU (setup(device_sampler=True))
extra_inputs = None  # None | 'doy' | 'lon': the additional-input variants of revision1/additional_inputs/
timelist_all = None  # 'doy': day of year of every day of `data` (..._doy.py:128)
min_lonidx = max_lonidx = 0   # 'lon': normalisation of the x index (..._lon.py:127-129)


def _extra_channels(idcs_batch):
    """Extra condition channels of the additional-input variants, (n, ndomain, ndomain, k) or None:
    'doy' -> sin / cos of the day of year (gan_train_cwgangp_pixelnorm_doy.py:175-185),
    'lon' -> x index normalised with the min / max over the valid indices (..._lon.py:176-185)."""
    if extra_inputs is None:
        return None
    if extra_inputs == 'doy':
        batch_doy = timelist_all[idcs_batch[:, 0]]
        batch_doy = np.tile(batch_doy, (1, ndomain, ndomain, 1)).T
        return np.concatenate([np.sin(2 * np.pi * batch_doy / 365), np.cos(2 * np.pi * batch_doy / 365)], axis=-1)
    batch_lon = (idcs_batch[:, 2] - min_lonidx) / max_lonidx
    return np.tile(batch_lon, (1, ndomain, ndomain, 1)).T


def _with_extra(batch_cond, idcs_batch):
    extra = _extra_channels(idcs_batch)
    if extra is None:
        return batch_cond
    if isinstance(batch_cond, np.ndarray):
        return np.concatenate([batch_cond, extra], axis=-1)
    import torch
    return torch.cat([batch_cond, torch.as_tensor(extra, dtype=batch_cond.dtype, device=batch_cond.device)], dim=-1)


def setup(data_array=None, valid_indices=None, seed=0, gen_mode=None, device_sampler=None, extra=None, timelist=None,
          domain=None):
    """Build what the reference builds at import (:117-140, :360-408).  device_sampler (default: env RDG_DEVICE_SAMPLER=1):
    keep the radar array in HBM and gather / normalise the batches there (replaces the GeneratorEnqueuer workers, :440-449).
    extra = 'doy' (with timelist = day of year per day) or 'lon': the additional-input variants (n_channel = 3 / 2);
    domain = 64: the large-domain variant (alternative_domains/gan_train_cwgangp_pixelnorm_largedomain.py:59)."""
    global data, indices_all, n_samples, generator, critic, critic_model, generator_model, optimizer, trainer, sampler
    global extra_inputs, timelist_all, min_lonidx, max_lonidx, n_channel, ndomain
    from rdg_b200.engine import Adam, GanTrainer
    if extra not in (None, 'doy', 'lon'):
        raise ValueError("extra must be None, 'doy' or 'lon'")
    extra_inputs = extra
    _GRAPHS.clear()
    n_channel = {None: 1, 'doy': 3, 'lon': 2}[extra]
    if domain is not None:
        ndomain = int(domain)
    global params
    params = f'{startdate}-{enddate}-tp_thresh_daily{tp_thresh_daily}_n_thresh{n_thresh}_ndomain{ndomain}_stride{stride}'
    np.random.seed(seed)
    data = synthetic_radar(seed=seed) if data_array is None else np.asarray(data_array, np.float32)
    assert data.ndim == 4 and data.shape[1] == nhours
    if valid_indices is None:
        ny, nx = data.shape[2:]
        valid_indices = [(t, y, x) for t in range(data.shape[0]) for y in range(0, ny - ndomain + 1, stride)
                         for x in range(0, nx - ndomain + 1, stride)]
    indices_all = np.array(valid_indices)
    n_samples = len(indices_all)
    if extra == 'doy':
        timelist_all = np.arange(data.shape[0]) % 365 + 1 if timelist is None else np.asarray(timelist)
        assert len(timelist_all) == data.shape[0]
    if extra == 'lon':
        min_lonidx, max_lonidx = np.min(indices_all[:, 2]), np.max(indices_all[:, 2])
    generator = create_generator()        # generator first, like the reference (:361-362)
    critic = create_discriminator()
    optimizer = Adam(lr=0.0001, beta_1=0, beta_2=0.9)     # reference :385
    # Under torch.distributed (data-parallel replicas) GanTrainer broadcasts rank 0's weights / Adam state and gives every rank its
    # own device random stream; the host-side draws (batch indices, latent noise: np.random below) get a per-rank seed as well,
    # AFTER the weight initialisation above, which therefore is the same on every rank.
    # precision of the FROZEN generator forward inside the critic step: FP32 in the FP32 parity mode, the fp16 tensor-core forward
    # (<= 1e-2 of the oracle) in the tensor-core training mode unless RDG_TRAIN_GEN_MODE says otherwise
    default_gen_mode = 'fp16' if os.environ.get('RDG_TRAIN_MODE', 'fp32') == 'tf32' else 'fp32'
    trainer = GanTrainer(generator, critic, optimizer, gen_mode=gen_mode or os.environ.get('RDG_TRAIN_GEN_MODE', default_gen_mode),
                         seed=seed)
    if trainer.rank:
        np.random.seed(seed + trainer.rank)
    critic_model = _CriticModel(trainer)
    generator_model = _GeneratorModel(trainer)
    if device_sampler is None:
        device_sampler = os.environ.get('RDG_DEVICE_SAMPLER', '0') == '1'
    sampler = None
    if device_sampler:
        from rdg_b200.sampler import DeviceSampler
        sampler = DeviceSampler(_ctx(), data, indices_all, ndomain, norm_scale)
    return trainer


def _windows(ixs):
    idcs = indices_all[ixs]
    batch = np.stack([data[t, :, y:y + ndomain, x:x + ndomain] for t, y, x in idcs])   # view_as_windows, :151-152
    return np.expand_dims(batch, -1)


def generate_real_samples(n_batch):
    """reference :143-174"""
    while sampler is not None:
        ixs = np.random.randint(n_samples, size=n_batch)                       # same draw as below (:147)
        batch, batch_cond = sampler.gather(ixs)
        yield [batch, _with_extra(batch_cond, indices_all[ixs])]
    while True:
        ixs = np.random.randint(n_samples, size=n_batch)
        batch = _windows(ixs)
        batch_cond = np.sum(batch, axis=1)
        for i in range(n_batch):
            batch[i] = batch[i] / batch_cond[i]
        batch_cond = batch_cond / norm_scale
        batch_cond = _with_extra(batch_cond, indices_all[ixs])
        assert batch.shape == (n_batch, nhours, ndomain, ndomain, 1)
        assert batch_cond.shape == (n_batch, ndomain, ndomain, n_channel)
        assert ~np.any(np.isnan(batch)) and ~np.any(np.isnan(batch_cond))
        assert np.max(batch) <= 1 and np.min(batch) >= 0
        yield [batch, batch_cond]


def generate_latent_points(n_batch):
    """reference :177-193"""
    latent = np.random.normal(size=(n_batch, latent_dim))
    ixs = np.random.randint(0, n_samples, size=n_batch)
    if sampler is not None:
        _, batch_cond = sampler.gather(ixs, with_batch=False)
        return [latent, _with_extra(batch_cond, indices_all[ixs])]
    batch_cond = np.sum(_windows(ixs), axis=1) / norm_scale
    batch_cond = _with_extra(batch_cond, indices_all[ixs])
    assert batch_cond.shape == (n_batch, ndomain, ndomain, n_channel)
    assert ~np.any(np.isnan(batch_cond))
    return [latent, batch_cond]


def generate_latent_points_as_generator(n_batch):
    while True:
        yield generate_latent_points(n_batch)


def generate_fake_samples(n_batch):
    """reference :201-206"""
    latent, cond = generate_latent_points(n_batch)
    if not isinstance(cond, np.ndarray):
        cond = cond.cpu().numpy()
    return [generator.predict([latent, cond]), cond]


def generate(cond):
    """reference :209-212"""
    latent = np.random.normal(size=(1, latent_dim))
    return generator.predict([latent, np.expand_dims(cond, 0)])


def train(n_epochs, _batch_size, start_epoch=0, bat_per_epo=None, save=True):
    """reference :431-521.  Same schedule (n_disc critic steps, then one generator step), same NaN guard
    (:487-488), same per-epoch checkpoints (:520-521) written as Keras-layout HDF5; the matplotlib sample /
    loss figures (:494-518) are outside the accelerated path and omitted, hist.csv is kept.
    In the tensor-core training mode (RDG_TRAIN_MODE=tf32; RDG_TRAIN_GRAPH=0 to opt out) an iteration is not issued as six
    train_on_batch calls but replayed from one CUDA graph (GanTrainer.capture_iteration): the n_disc real batches and the
    generator step's conditions are copied into the graph's input buffers, the latent noise / interpolation weights / dropout
    masks come from the device random streams (same distributions as :470, :223, :289-301), and the losses of the LAST critic
    step and of the generator step are read once per iteration for the reference's log line and NaN guard."""
    global batch_size
    batch_size = _batch_size
    graph = None
    if trainer.train_mode == 'tf32' and os.environ.get('RDG_TRAIN_GRAPH', '1') == '1':
        key = (id(trainer), batch_size, n_disc)          # captured once per trainer and batch size (a capture costs ~0.4 s)
        if key not in _GRAPHS:
            _GRAPHS.clear()
            _GRAPHS[key] = trainer.capture_iteration(batch_size, n_critic=n_disc, preserve_state=True)
        graph = _GRAPHS[key]
    sample_gen = generate_real_samples(batch_size)                 # GeneratorEnqueuer worker processes in the reference
    gan_sample_gen = generate_latent_points_as_generator(batch_size)
    valid = -np.ones((batch_size, 1)); fake = np.ones((batch_size, 1)); dummy = np.zeros((batch_size, 1))
    if bat_per_epo is None:
        bat_per_epo = int(n_samples / batch_size)
    staged = None

    def _draw_iteration():
        return [next(sample_gen) for _ in range(n_disc)], next(gan_sample_gen)[1]

    for i in range(n_epochs):
        epoch = 1 + i + start_epoch
        for j in range(bat_per_epo):
            if graph is not None:
                if staged is None:
                    staged = _draw_iteration()
                reals, cond = staged
                for k, (X_real, cond_real) in enumerate(reals):
                    graph.x_real[k].copy_(_ctx().dev(X_real).reshape(graph.x_real[k].shape), non_blocking=True)
                    graph.cond[k].copy_(_ctx().dev(cond_real).reshape(graph.cond[k].shape), non_blocking=True)
                graph.cond_gen.copy_(_ctx().dev(cond).reshape(graph.cond_gen.shape), non_blocking=True)
                graph.replay()
                # the next iteration's batches are drawn on the host while the GPU runs this one (same np.random order)
                staged = _draw_iteration() if (j + 1 < bat_per_epo or i + 1 < n_epochs) else None
                d_last = graph.d_losses[-1].cpu().numpy()              # [total, l_valid, l_fake, l_gp] of the last critic step
                d_loss = np.mean([d_last[1], d_last[2]])
                g_loss = float(graph.g_loss.item())
            else:
                for _ in range(n_disc):
                    X_real, cond_real = next(sample_gen)
                    latent = np.random.normal(size=(batch_size, latent_dim))
                    d_loss = critic_model.train_on_batch([X_real, cond_real, latent], [valid, fake, dummy])
                    d_loss = np.mean([d_loss[1], d_loss[2]])
                latent, cond = next(gan_sample_gen)
                g_loss = generator_model.train_on_batch([latent, cond], valid)
            print(f'{epoch}, {j + 1}/{bat_per_epo}, d_loss {d_loss} g:{g_loss} ')
            if np.isnan(g_loss) or np.isnan(d_loss):
                raise ValueError('encountered nan in g_loss and/or d_loss')
            hist['d_loss'].append(d_loss)
            hist['g_loss'].append(g_loss)
        trainer.check_exchange()          # data-parallel runs: no barrier of the gradient exchange timed out during the epoch
        with open('hist.csv', 'w') as f:
            f.write(',d_loss,g_loss\n')
            for k, (d, g) in enumerate(zip(hist['d_loss'], hist['g_loss'])):
                f.write(f'{k},{d},{g}\n')
        if save:
            # Weight files in the Keras HDF5 layout (model_weights tree + layer_names / weight_names attributes), WITHOUT the
            # model_config attribute tf.keras.models.load_model needs (it holds the marshalled Lambda of :349-350, which only
            # TensorFlow can write): load them with create_generator().load_weights(path) on the reference side, or with
            # rdg_b200.hdf5.load_keras_weights here.
            os.makedirs(outdir, exist_ok=True)
            _hdf5.save_keras_weights(f'{outdir}/gen_{params}_{epoch:04d}.h5', generator.get_weights(), 'generator')
            _hdf5.save_keras_weights(f'{outdir}/disc_{params}_{epoch:04d}.h5', critic.get_weights(), 'critic')
