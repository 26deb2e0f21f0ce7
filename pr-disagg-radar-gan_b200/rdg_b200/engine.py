"""Host-side mirror of the Keras model objects the reference uses on its hot path.

`Generator` / `Critic` stand in for the `tf.keras.Model`s built by create_generator() /
create_discriminator() (gan_train_cwgangp_pixelnorm.py:272-357): `.predict([...])`,
`.get_weights()`, `.set_weights()`, `.inputs`-like metadata.  All arithmetic happens in
librdg_b200.so on a B200; torch is only used for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np
import torch

from . import _lib
from . import weights as W

_DEFAULT_MODE = os.environ.get("RDG_MODE", "fp16")


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("rdg_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


class Context:
    """One rdg_ctx per (device, nd, ncond).  Holds both nets' weights and the workspaces."""

    def __init__(self, nd=16, ncond=1, device=None, max_chunk=0):
        _require_cuda()
        self.lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        self.nd, self.ncond = int(nd), int(ncond)
        h = C.c_void_p()
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device="cuda")  # make sure the primary context exists
            _lib.check(self.lib.rdg_ctx_create(C.byref(h), self.device, self.nd, self.ncond, int(max_chunk)))
        self.handle = h
        mc, sm = C.c_int(), C.c_int()
        self.lib.rdg_ctx_info(self.handle, None, None, C.byref(mc), C.byref(sm))
        self.max_chunk, self.sm_count = mc.value, sm.value
        self._lock = threading.Lock()

    def close(self):
        if getattr(self, "handle", None):
            self.lib.rdg_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def dev(self, a, dtype=torch.float32):
        """numpy/torch -> contiguous float32 CUDA tensor on this context's device."""
        if isinstance(a, torch.Tensor):
            return a.to(device=f"cuda:{self.device}", dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=f"cuda:{self.device}")


class Generator:
    """Stand-in for the Keras generator model (create_generator, gan_train...py:312-357)."""

    def __init__(self, weights=None, nd=16, ncond=1, ctx=None, mode=None, seed=0):
        self.ctx = ctx or Context(nd, ncond)
        self.nd, self.ncond = self.ctx.nd, self.ctx.ncond
        self.latent_dim = W.LATENT_DIM
        self.mode = mode or _DEFAULT_MODE
        # Keras-like metadata used by the reference: gen.inputs[0].shape[1] (raindisagg_gan_pretrained.py:47)
        self.input_shapes = [(None, self.latent_dim), (None, self.nd, self.nd, self.ncond)]
        self.output_shape = (None, W.NHOURS, self.nd, self.nd, 1)
        self.set_weights(weights if weights is not None else W.init_generator_weights(seed, self.nd, self.ncond))

    def set_weights(self, weights):
        ws = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
        W.check_shapes(ws, W.generator_shapes(self.nd, self.ncond), "generator")
        ptrs, sizes = _lib.float_ptr_array(ws)
        _lib.check(self.ctx.lib.rdg_generator_set_weights(self.ctx.handle, ptrs, sizes, len(ws)))

    def get_weights(self):
        shapes = W.generator_shapes(self.nd, self.ncond)
        ws = [np.empty(s, np.float32) for s in shapes]
        ptrs, sizes = _lib.float_ptr_array(ws)
        _lib.check(self.ctx.lib.rdg_generator_get_weights(self.ctx.handle, ptrs, sizes, len(ws)))
        return ws

    # ---- device-resident forward (throughput path)
    def forward_device(self, latent, cond, scen_per_cond=1, mode=None, out_mm=False, norm_scale=W.NORM_SCALE,
                       out=None, check=True):
        """latent [B,100] cuda f32; cond [ceil(B/spc),nd,nd,ncond] cuda f32 -> out [B,24,nd,nd] cuda f32."""
        ctx = self.ctx
        B = int(latent.shape[0])
        if out is None:
            out = torch.empty((B, W.NHOURS, self.nd, self.nd), device=latent.device, dtype=torch.float32)
        flag = torch.zeros(1, device=latent.device, dtype=torch.int32) if check else None
        m = _lib.MODES[mode or self.mode]
        _lib.check(ctx.lib.rdg_generator_forward(
            ctx.handle, C.c_void_p(latent.data_ptr()), C.c_void_p(cond.data_ptr()), int(scen_per_cond),
            C.c_void_p(out.data_ptr()), B, m, _lib.OUT_MM if out_mm else _lib.OUT_FRACTION, float(norm_scale),
            C.c_void_p(flag.data_ptr()) if check else None, ctx._stream()))
        if check and int(flag.item()) != 0:
            raise _lib.NonFiniteError("found nan in output of per_gridpoint_softmax")
        return out

    # ---- Keras surface
    def predict(self, inputs, batch_size=None, mode=None):
        """gen.predict([latent, cond_batch]) -> (B,24,nd,nd,1) float32 fractions
        (raindisagg_gan_pretrained.py:60, generate_and_evaluate.py:223,406,527)."""
        latent, cond = inputs
        latent = np.ascontiguousarray(latent, dtype=np.float32)
        cond = np.ascontiguousarray(cond, dtype=np.float32)
        B = latent.shape[0]
        if latent.shape != (B, self.latent_dim) or cond.shape != (B, self.nd, self.nd, self.ncond):
            raise ValueError(f"predict: expected latent (B,{self.latent_dim}) and cond (B,{self.nd},{self.nd},"
                             f"{self.ncond}), got {latent.shape} and {cond.shape}")
        out = np.empty((B, W.NHOURS, self.nd, self.nd, 1), np.float32)
        if B == 0:
            return out
        m = _lib.MODES[mode or self.mode]
        _lib.check(self.ctx.lib.rdg_generate_host(
            self.ctx.handle, latent.ctypes.data_as(C.c_void_p), cond.ctypes.data_as(C.c_void_p), 1,
            out.ctypes.data_as(C.c_void_p), B, m, _lib.OUT_FRACTION, float(W.NORM_SCALE)))
        return out

    def generate_ensemble_host(self, latent, cond_norm, scen_per_cond, out=None, mode=None, out_mm=True,
                               norm_scale=W.NORM_SCALE):
        """Host-buffer ensemble generation: latent [B,100], cond_norm [ceil(B/spc),nd,nd,ncond] (already
        divided by norm_scale) -> [B,24,nd,nd] mm/h (or fractions).  numpy arrays or pinned CPU torch tensors."""
        def ptr(a):
            return C.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor) else a.ctypes.data_as(C.c_void_p)
        B = int(latent.shape[0])
        if out is None:
            out = np.empty((B, W.NHOURS, self.nd, self.nd), np.float32)
        m = _lib.MODES[mode or self.mode]
        _lib.check(self.ctx.lib.rdg_generate_host(
            self.ctx.handle, ptr(latent), ptr(cond_norm), int(scen_per_cond), ptr(out), B, m,
            _lib.OUT_MM if out_mm else _lib.OUT_FRACTION, float(norm_scale)))
        return out


class Critic:
    """Stand-in for the Keras critic model (create_discriminator, gan_train...py:272-309)."""

    def __init__(self, weights=None, nd=16, ncond=1, ctx=None, seed=1):
        self.ctx = ctx or Context(nd, ncond)
        self.nd, self.ncond = self.ctx.nd, self.ctx.ncond
        self.set_weights(weights if weights is not None else W.init_critic_weights(seed, self.nd, self.ncond))

    def set_weights(self, weights):
        ws = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
        W.check_shapes(ws, W.critic_shapes(self.nd, self.ncond), "critic")
        ptrs, sizes = _lib.float_ptr_array(ws)
        _lib.check(self.ctx.lib.rdg_critic_set_weights(self.ctx.handle, ptrs, sizes, len(ws)))

    def get_weights(self):
        ws = [np.empty(s, np.float32) for s in W.critic_shapes(self.nd, self.ncond)]
        ptrs, sizes = _lib.float_ptr_array(ws)
        _lib.check(self.ctx.lib.rdg_critic_get_weights(self.ctx.handle, ptrs, sizes, len(ws)))
        return ws

    def forward_device(self, sample, cond, masks=None):
        """sample [B,24,nd,nd(,1)] cuda f32, cond [B,nd,nd,ncond] cuda f32 -> scores [B] cuda f32."""
        ctx = self.ctx
        B = int(sample.shape[0])
        score = torch.empty((B,), device=sample.device, dtype=torch.float32)
        mp = None
        if masks is not None:
            mp = (C.c_void_p * 4)(*[C.c_void_p(m.data_ptr()) for m in masks])
        _lib.check(ctx.lib.rdg_critic_forward(ctx.handle, C.c_void_p(sample.data_ptr()), C.c_void_p(cond.data_ptr()),
                                              mp, C.c_void_p(score.data_ptr()), B, ctx._stream()))
        return score

    def predict(self, inputs, masks=None):
        """critic.predict([sample, cond]) -> (B,1) float32."""
        sample, cond = inputs
        s = self.ctx.dev(sample)
        c = self.ctx.dev(cond)
        m = None if masks is None else [self.ctx.dev(x) for x in masks]
        out = self.forward_device(s, c, m)
        torch.cuda.synchronize(self.ctx.device)
        return out.cpu().numpy().reshape(-1, 1)
