"""Device-side batch sampler: generate_real_samples / generate_latent_points of
gan_train_cwgangp_pixelnorm.py:143-198 with the radar array resident in HBM.  The random draws are the reference's own
numpy calls in the reference's order (so `np.random.seed` reproduces the same batches); gather, daily sum, fraction and
condition normalisation run in one CUDA kernel (`rdg_sample_windows`) and the batch never exists on the host -- this
replaces the GeneratorEnqueuer worker processes (:440-449)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import weights as W


class DeviceSampler:
    def __init__(self, ctx, data, indices_all, ndomain=16, norm_scale=W.NORM_SCALE):
        self.ctx, self.nd, self.norm_scale = ctx, int(ndomain), float(norm_scale)
        data = data if isinstance(data, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(data, dtype=np.float32))
        if data.dim() != 4 or data.shape[1] != W.NHOURS or data.dtype != torch.float32:
            raise ValueError("data must be float32 (days, 24, ny, nx)")                       # reference asserts :131-138
        self.data = data.to(f"cuda:{ctx.device}").contiguous()
        self.indices_all = np.ascontiguousarray(np.asarray(indices_all), dtype=np.int32)
        if self.indices_all.ndim != 2 or self.indices_all.shape[1] != 3:
            raise ValueError("indices_all must have shape (n_samples, 3)")
        self.n_samples = len(self.indices_all)
        self._idx_dev = torch.as_tensor(self.indices_all, device=self.data.device)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.data.device)

    def gather(self, ixs, with_batch=True, check=True):
        """ixs: positions in indices_all (numpy int array) -> (batch [n,24,nd,nd,1] or None, cond [n,nd,nd,1]) cuda f32."""
        n, nd, dev = len(ixs), self.nd, self.data.device
        rows = self._idx_dev[torch.as_tensor(np.asarray(ixs, dtype=np.int64), device=dev)].contiguous()
        batch = torch.empty((n, W.NHOURS, nd, nd, 1), device=dev) if with_batch else None
        cond = torch.empty((n, nd, nd, 1), device=dev)
        if check:
            self._flag.zero_()
        _, _, ny, nx = self.data.shape
        _lib.check(self.ctx.lib.rdg_sample_windows(
            C.c_void_p(self.data.data_ptr()), int(self.data.shape[0]), int(ny), int(nx), C.c_void_p(rows.data_ptr()), n, nd,
            self.norm_scale, C.c_void_p(batch.data_ptr()) if with_batch else None, C.c_void_p(cond.data_ptr()),
            C.c_void_p(self._flag.data_ptr()) if check else None, self.ctx._stream()))
        if check:
            f = int(self._flag.item())
            if f & 2:
                raise IndexError("window outside the data array")
            if f & 1:
                raise AssertionError("fractions outside [0, 1] or NaN (a window with a dry pixel?)")   # reference :167-170
        return batch, cond

    def real_samples(self, n_batch):
        """generate_real_samples (:143-174) as a generator of [batch, batch_cond] device tensors."""
        while True:
            ixs = np.random.randint(self.n_samples, size=n_batch)                              # :147
            batch, cond = self.gather(ixs)
            yield [batch, cond]

    def latent_points(self, n_batch, latent_dim=W.LATENT_DIM):
        """generate_latent_points (:177-193): [latent (host draw, as the reference), batch_cond (device)]."""
        latent = np.random.normal(size=(n_batch, latent_dim))                                  # :179
        ixs = np.random.randint(0, self.n_samples, size=n_batch)                               # :181
        _, cond = self.gather(ixs, with_batch=False)
        return [latent, cond]
