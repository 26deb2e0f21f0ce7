"""Drop-in for the reference's `raindisagg_gan_pretrained.py` (93 lines; the only importable module
of sipposip/pr-disagg-radar-gan) with the TensorFlow call boundary replaced by the B200 library.

Same names and behaviour: `norm_scale`, `generator_file`, `PixelNormalization`, `gen`, `latent_dim`,
`generate_scenarios(cond, n_scenarios)`, `plot_scenarios(scenarios)` (reference :13-14, :18-45, :47,
:52-65, :68-90).  Differences, all at the boundary:
  * the Keras HDF5 file is read by rdg_b200.hdf5 (no h5py / TF); `RDG_GENERATOR_FILE` may override
    the path (the shipped blob is not part of the public repository snapshot);
  * `gen.predict` runs on cuda:0 through librdg_b200.so; arithmetic mode from `RDG_MODE`
    (fp16 [default] | bf16 tensor-core operands, fp32 SIMT);
  * `gen.compile(...)` (:49, a TF 2.1 workaround) has no equivalent and is not needed.
"""
import os

import numpy as np

from rdg_b200 import hdf5 as _hdf5
from rdg_b200 import weights as _W

norm_scale = 127.4
generator_file = os.environ.get(
    "RDG_GENERATOR_FILE",
    'trained_models/gen_20090101-20161231-tp_thresh_daily5_n_thresh20_ndomain16_stride16_0020.h5')


class PixelNormalization:
    """x / sqrt(mean_c(x^2) + 1e-8) over the channel axis (reference :18-39), on the GPU."""

    def __init__(self, **kwargs):
        pass

    def call(self, inputs):
        import ctypes as C
        import torch
        from rdg_b200 import _lib
        x = torch.as_tensor(np.ascontiguousarray(inputs, dtype=np.float32), device="cuda")
        y = torch.empty_like(x)
        c = x.shape[-1]
        _lib.check(_lib.load().rdg_pixelnorm(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), x.numel() // c, c, 0, None))
        torch.cuda.synchronize()
        return y.cpu().numpy()

    __call__ = call

    def compute_output_shape(self, input_shape):
        return input_shape


def _load_generator(path):
    if not os.path.exists(path):
        raise OSError(f"Unable to open file (unable to open file: name = '{path}'); set RDG_GENERATOR_FILE or "
                      f"place the reference's trained generator there")
    from rdg_b200.engine import Generator
    ws = _hdf5.load_keras_weights(path)
    nd = int(round(np.sqrt(ws[0].shape[0] - _W.LATENT_DIM)))   # Dense input = latent + nd*nd*1
    return Generator(ws, nd=nd, ncond=1)


# load the trained generator network (reference :43-45)
gen = _load_generator(generator_file)
latent_dim = gen.latent_dim            # gen.inputs[0].shape[1] (reference :47)


def generate_scenarios(cond, n_scenarios):
    """Reference :52-65, statement for statement (same numpy-global-RNG draw, same dtype quirks)."""
    # the generator takes normalized daily sums, so we have to divide by norm_scale
    cond = cond / norm_scale
    # for each cond, make several predictions with different latent noise
    latent = np.random.normal(size=(n_scenarios, latent_dim))
    cond_batch = np.repeat(cond[np.newaxis], repeats=n_scenarios, axis=0)
    generated = gen.predict([latent, cond_batch])
    # remove empty channel dimension
    generated = generated.squeeze()
    # this now contains daily fractions. convert to mm/h
    generated_precip = generated * cond.squeeze() * norm_scale
    return generated_precip


def plot_scenarios(scenarios):
    """Reference :68-90 (matplotlib; not part of the accelerated path)."""
    from matplotlib import pyplot as plt
    from matplotlib.colors import LogNorm
    nrows = len(scenarios)
    fig = plt.figure(figsize=(24, nrows))
    plt.axis('off')
    for iplot in range(nrows):
        for jplot in range(24):
            ax = plt.subplot(nrows, 24, iplot * 24 + jplot + 1)
            if iplot == 0:
                ax.annotate(f'{jplot:02d}:00', xy=(0.5, 1), xytext=(0, 5), xycoords='axes fraction',
                            textcoords='offset points', size='large', ha='center', va='baseline')
            im = plt.imshow(scenarios[iplot, jplot - 1, :, :], cmap=plt.cm.gist_earth_r, norm=LogNorm(vmin=0.01, vmax=50))
            plt.axis('off')
    fig.subplots_adjust(right=0.93)
    cbar = fig.colorbar(im, cax=fig.add_axes([0.93, 0.15, 0.007, 0.7]))
    cbar.set_label('fraction of daily precipitation', fontsize=16)
    return fig
