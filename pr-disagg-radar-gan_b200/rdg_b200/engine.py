"""Host-side mirror of the Keras model objects the reference uses on its hot path.

`Generator` / `Critic` stand in for the `tf.keras.Model`s built by create_generator() /
create_discriminator() (gan_train_cwgangp_pixelnorm.py:272-357): `.predict([...])`,
`.get_weights()`, `.set_weights()`, `.inputs`-like metadata.  All arithmetic happens in
librdg_b200.so on a B200; torch is only used for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

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
        self.max_chunk, self.sm_count = mc.value, sm.value       # one context = one caller at a time (not thread-safe, like a Keras model)

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
        self.mode = mode or os.environ.get("RDG_MODE", _DEFAULT_MODE)
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

    def tc_layer_device(self, layer, x, mode="fp16"):
        """One tensor-core layer alone (tests): x [B,T,H,W,Cin] cuda f32 (rounded to 16 bit inside) ->
        LeakyReLU(PixelNorm(conv3(upsample2(x)) + b)) [B,2T,2H,2W,Cout] cuda f32 (gan_train...py:330-343)."""
        cout = (256, 128, 64)[layer]
        B, T, H, Wd, _ = x.shape
        y = torch.empty((B, 2 * T, 2 * H, 2 * Wd, cout), device=x.device, dtype=torch.float32)
        _lib.check(self.ctx.lib.rdg_tc_layer(self.ctx.handle, int(layer), _lib.MODES[mode], C.c_void_p(x.data_ptr()),
                                             C.c_void_p(y.data_ptr()), int(B), self.ctx._stream()))
        return y

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


    # ---- ensemble statistics on the device
    def ensemble_stats_device(self, fields, scen_per_cond, obs=None, want_crps_field=False):
        """fields [n_cond*spc,24,nd,nd] cuda f32 (members of a condition contiguous), obs [n_cond,24,nd,nd] cuda f32 or None
        -> dict(area_mean [n_cond*spc,24], crps_area_mean [n_cond,24], crps [n_cond,24,nd,nd]) as cuda tensors:
        np.mean(generated, (2,3)) (generate_and_evaluate.py:533-535) and properscoring.crps_ensemble(obs, generated, axis=0)
        with its area mean (generate_and_evaluate_crps.py:189-191)."""
        B = int(fields.shape[0])
        spc = int(scen_per_cond)
        if B % spc:
            raise ValueError("ensemble_stats: the batch must hold whole ensembles")
        n_cond = B // spc
        dev = fields.device
        out = {"area_mean": torch.empty((B, W.NHOURS), device=dev, dtype=torch.float32)}
        crps = cam = None
        if obs is not None:
            cam = out["crps_area_mean"] = torch.empty((n_cond, W.NHOURS), device=dev, dtype=torch.float32)
            if want_crps_field:
                crps = out["crps"] = torch.empty((n_cond, W.NHOURS, self.nd, self.nd), device=dev, dtype=torch.float32)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        _lib.check(self.ctx.lib.rdg_ensemble_stats(self.ctx.handle, p(fields), n_cond, spc, p(obs), p(out["area_mean"]),
                                                   p(crps), p(cam), self.ctx._stream()))
        return out

    def generate_ensemble_stats_host(self, latent, cond_norm, scen_per_cond, obs_mm=None, mode=None, out_mm=True,
                                     norm_scale=W.NORM_SCALE):
        """The CRPS loop of generate_and_evaluate_crps.py:177-191 in one call: generate scen_per_cond scenarios for every
        condition and reduce them on the GPU; only the statistics come back.  latent [n_cond*spc,100], cond_norm
        [n_cond,nd,nd,ncond], obs_mm [n_cond,24,nd,nd] (real_precip) or None -> (area_mean [n_cond*spc,24],
        crps_area_mean [n_cond,24] or None).  numpy arrays or pinned CPU torch tensors."""
        def ptr(a):
            if a is None:
                return None
            return C.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor) else a.ctypes.data_as(C.c_void_p)
        B = int(latent.shape[0])
        n_cond = B // int(scen_per_cond)
        am = np.empty((B, W.NHOURS), np.float32)
        cr = np.empty((n_cond, W.NHOURS), np.float32) if obs_mm is not None else None
        m = _lib.MODES[mode or self.mode]
        _lib.check(self.ctx.lib.rdg_generate_stats_host(
            self.ctx.handle, ptr(latent), ptr(cond_norm), int(scen_per_cond), ptr(obs_mm), B, m,
            _lib.OUT_MM if out_mm else _lib.OUT_FRACTION, float(norm_scale), ptr(am), ptr(cr)))
        return am, cr


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

    def forward_device(self, sample, cond, masks=None, mode="fp32"):
        """sample [B,24,nd,nd(,1)] cuda f32, cond [B,nd,nd,ncond] cuda f32 -> scores [B] cuda f32.
        mode "fp16" / "bf16": tensor-core scoring (inference only, no dropout masks)."""
        ctx = self.ctx
        B = int(sample.shape[0])
        score = torch.empty((B,), device=sample.device, dtype=torch.float32)
        if mode != "fp32":
            if masks is not None:
                raise ValueError("the tensor-core critic mode is inference only (no dropout masks)")
            _lib.check(ctx.lib.rdg_critic_forward_tc(ctx.handle, C.c_void_p(sample.data_ptr()), C.c_void_p(cond.data_ptr()),
                                                     C.c_void_p(score.data_ptr()), B, _lib.MODES[mode], ctx._stream()))
            return score
        mp = None
        if masks is not None:
            mp = (C.c_void_p * 4)(*[C.c_void_p(m.data_ptr()) for m in masks])
        _lib.check(ctx.lib.rdg_critic_forward(ctx.handle, C.c_void_p(sample.data_ptr()), C.c_void_p(cond.data_ptr()),
                                              mp, C.c_void_p(score.data_ptr()), B, ctx._stream()))
        return score

    def predict(self, inputs, masks=None, mode="fp32"):
        """critic.predict([sample, cond]) -> (B,1) float32."""
        sample, cond = inputs
        s = self.ctx.dev(sample)
        c = self.ctx.dev(cond)
        m = None if masks is None else [self.ctx.dev(x) for x in masks]
        out = self.forward_device(s, c, m, mode=mode)
        torch.cuda.synchronize(self.ctx.device)
        return out.cpu().numpy().reshape(-1, 1)


# --------------------------------------------------------------------------------------------
# training (gan_train_cwgangp_pixelnorm.py:360-408, 431-491)
# --------------------------------------------------------------------------------------------
class _DevBuf:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class Adam:
    """tf.optimizers.Adam(lr, beta_1, beta_2) in its Keras OptimizerV2 form.  ONE object is
    shared by the critic and the generator models (gan_train_cwgangp_pixelnorm.py:385, 391, 408), so
    the step counter `iterations` is shared as well."""

    def __init__(self, lr=0.0001, beta_1=0.0, beta_2=0.9, epsilon=1e-7):
        self.lr, self.beta_1, self.beta_2, self.epsilon = lr, beta_1, beta_2, epsilon
        self.iterations = 0


class GanTrainer:
    """The two compiled Keras models of the reference (`critic_model`, `generator_model`) as one object.

    Data-parallel: if torch.distributed is initialised, gradients are summed over ranks with one NCCL
    all-reduce of the flat FP32 gradient buffer per optimizer step and scaled by 1/world (the losses are
    batch means, :215-216), so N ranks with batch B each reproduce one rank with batch N*B.
    """

    WHICH_GEN, WHICH_CRITIC = 0, 1

    def __init__(self, generator, critic, optimizer=None, gen_mode="fp32", seed=0, process_group=None, train_mode=None,
                 dropout=True):
        """train_mode "fp32": FP32 SIMT gradients (<= 2e-5 parity mode); "tf32": every wide contraction on tcgen05 kind::tf32
        (default: env RDG_TRAIN_MODE or "fp32").  Under torch.distributed the replicas are made identical here (weights and
        Adam moments broadcast from rank 0) and every rank gets its own random stream (seed + rank) for noise / alpha / masks."""
        assert generator.ctx is critic.ctx, "generator and critic must share one Context"
        self.ctx, self.generator, self.critic = generator.ctx, generator, critic
        self.optimizer = optimizer or Adam()
        self.gen_mode = gen_mode
        self.pg = process_group
        self.train_mode = train_mode or os.environ.get("RDG_TRAIN_MODE", "fp32")
        if self.train_mode not in ("fp32", "tf32"):
            raise ValueError("train_mode must be 'fp32' or 'tf32'")
        self.dropout = bool(dropout)
        self.rank = self._rank()
        self.seed = int(seed)
        self._gen = torch.Generator(device=f"cuda:{self.ctx.device}")
        self._gen.manual_seed(self.seed + self.rank)
        self.profile_comm = False
        self._comm_events = []
        self.last_comm_ms = 0.0
        self._bufs = {}
        self._graph = None
        self._set_mode()
        self._push_counters(self.optimizer.iterations, 0)
        self.peer_exchange = False
        if self._world() > 1:
            self.sync_replicas()
            self._connect_peers()

    def _connect_peers(self):
        """Map the other ranks' gradient buffers through CUDA IPC (csrc/dp_peer.cu) so that the gradient exchange is the library's
        own NVLink kernel (graph-capturable) instead of an NCCL call.  Used when every rank of the group sits on this node and the
        mapping works on all of them (agreed collectively); otherwise, or with RDG_PEER_ALLREDUCE=0, the exchange stays on NCCL."""
        import socket
        import zlib
        import torch.distributed as dist
        world, ctx = self._world(), self.ctx
        if world > 8 or os.environ.get("RDG_PEER_ALLREDUCE", "1") == "0" or dist.get_backend(self.pg) != "nccl":
            return
        if getattr(ctx, "_peer_world", 0) == world:
            self.peer_exchange = True
            return
        dev = f"cuda:{ctx.device}"
        nb = int(ctx.lib.rdg_peer_handle_bytes())
        buf = (C.c_ubyte * nb)()
        ok = int(ctx.lib.rdg_peer_export(ctx.handle, buf) == 0)
        info = torch.tensor([zlib.crc32(socket.gethostname().encode()), ok], dtype=torch.int64, device=dev)
        infos = [torch.empty_like(info) for _ in range(world)]
        dist.all_gather(infos, info, group=self.pg)
        mine = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
        handles = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(handles, mine, group=self.pg)
        good = all(int(i[1]) == 1 and int(i[0]) == int(infos[0][0]) for i in infos)
        if good:
            blob = torch.cat(handles).cpu().numpy().tobytes()
            good = ctx.lib.rdg_peer_connect(ctx.handle, world, self.rank, (C.c_ubyte * len(blob)).from_buffer_copy(blob)) == 0
        agreed = torch.tensor([int(good)], dtype=torch.int64, device=dev)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=self.pg)
        if int(agreed.item()) == 1:
            ctx._peer_world = world
            self.peer_exchange = True
        elif good:
            ctx.lib.rdg_peer_disconnect(ctx.handle)

    def check_exchange(self):
        """Raise if a barrier of the peer-memory gradient exchange ever gave up waiting for another rank (a rank died or fell more
        than a minute behind): the replicas are then no longer identical.  Synchronises the device; call it at epoch boundaries."""
        if not self.peer_exchange:
            return
        w, t = C.c_int(0), C.c_int(0)
        _lib.check(self.ctx.lib.rdg_peer_status(self.ctx.handle, C.byref(w), C.byref(t)))
        if t.value:
            raise RuntimeError("data-parallel gradient exchange: a barrier timed out waiting for a peer rank")

    def _set_mode(self):
        _lib.check(self.ctx.lib.rdg_set_train_mode(self.ctx.handle, 1 if self.train_mode == "tf32" else 0))

    def _rank(self):
        import torch.distributed as dist
        return dist.get_rank(self.pg) if dist.is_available() and dist.is_initialized() else 0

    def _push_counters(self, adam_t, rng_ctr):
        """rng_ctr: (critic-step counter, generator-step counter) of the device Philox state, or one int for both."""
        rc = tuple(int(x) for x in np.atleast_1d(rng_ctr))
        rc = rc * 2 if len(rc) == 1 else rc
        a, r = C.c_longlong(int(adam_t)), (C.c_ulonglong * 2)(rc[0], rc[1])
        _lib.check(self.ctx.lib.rdg_train_state(self.ctx.handle, 1, C.byref(a), r))

    def _pull_counters(self):
        a, r = C.c_longlong(0), (C.c_ulonglong * 2)(0, 0)
        _lib.check(self.ctx.lib.rdg_train_state(self.ctx.handle, 0, C.byref(a), r))
        return int(a.value), (int(r[0]), int(r[1]))

    def sync_replicas(self, src=0):
        """Make every rank's weights and optimizer state those of `src` (data-parallel replicas must start identical: the
        gradient exchange only keeps them identical).  Derived weight images are rebuilt by the set_weights calls."""
        from .dist import broadcast_
        torch.cuda.synchronize(self.ctx.device)
        for which, net in ((0, self.generator), (1, self.critic)):
            broadcast_(self.param_tensor(which), src, self.pg)
            m, v = self._adam_tensors(which)
            broadcast_(m, src, self.pg)
            broadcast_(v, src, self.pg)
        it = torch.tensor([self.optimizer.iterations], device=f"cuda:{self.ctx.device}", dtype=torch.int64)
        broadcast_(it, src, self.pg)
        torch.cuda.synchronize(self.ctx.device)
        self.optimizer.iterations = int(it.item())
        self.generator.set_weights(self.generator.get_weights())     # re-derive the packed / folded images from the new master
        self.critic.set_weights(self.critic.get_weights())
        self._push_counters(self.optimizer.iterations, self._pull_counters()[1])

    # -- helpers
    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size(self.pg) if dist.is_available() and dist.is_initialized() else 1

    def grad_tensor(self, which):
        """Flat FP32 gradient buffer of the generator (0) / critic (1) as a torch view."""
        if ("g", which) not in self._bufs:
            p, n = C.c_void_p(), C.c_size_t()
            _lib.check(self.ctx.lib.rdg_grad_buffer(self.ctx.handle, which, C.byref(p), C.byref(n)))
            self._bufs[("g", which)] = torch.as_tensor(_DevBuf(p.value, n.value), device=f"cuda:{self.ctx.device}")
        return self._bufs[("g", which)]

    def param_tensor(self, which):
        if ("p", which) not in self._bufs:
            p, n = C.c_void_p(), C.c_size_t()
            _lib.check(self.ctx.lib.rdg_param_buffer(self.ctx.handle, which, C.byref(p), C.byref(n)))
            self._bufs[("p", which)] = torch.as_tensor(_DevBuf(p.value, n.value), device=f"cuda:{self.ctx.device}")
        return self._bufs[("p", which)]

    def _adam_tensors(self, which):
        if ("a", which) not in self._bufs:
            m, v, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
            _lib.check(self.ctx.lib.rdg_adam_buffers(self.ctx.handle, which, C.byref(m), C.byref(v), C.byref(n)))
            dev = f"cuda:{self.ctx.device}"
            self._bufs[("a", which)] = (torch.as_tensor(_DevBuf(m.value, n.value), device=dev),
                                        torch.as_tensor(_DevBuf(v.value, n.value), device=dev))
        return self._bufs[("a", which)]

    # -- checkpoint / resume.  The reference writes weights only (:520-521) and has no resume path
    # (`start_epoch` just labels plots, :526-529); this adds the optimizer state so training continues bit-identically.
    def state_dict(self):
        self.finish()
        torch.cuda.synchronize(self.ctx.device)
        self.check_exchange()
        adam_t, rng_ctr = self._pull_counters()
        assert adam_t == self.optimizer.iterations, "host / device Adam step counters diverged"
        sd = {"iterations": np.int64(self.optimizer.iterations), "rng_ctr": np.array(rng_ctr, np.uint64),
              "adam": np.array([self.optimizer.lr, self.optimizer.beta_1, self.optimizer.beta_2, self.optimizer.epsilon], np.float64),
              "rng_state": self._gen.get_state().cpu().numpy()}
        for which, name, net in ((0, "gen", self.generator), (1, "critic", self.critic)):
            for i, w in enumerate(net.get_weights()):
                sd[f"{name}_w{i}"] = w
            m, v = self._adam_tensors(which)
            sd[f"{name}_adam_m"], sd[f"{name}_adam_v"] = m.cpu().numpy(), v.cpu().numpy()
        return sd

    def load_state_dict(self, sd):
        self.generator.set_weights([sd[f"gen_w{i}"] for i in range(10)])
        self.critic.set_weights([sd[f"critic_w{i}"] for i in range(10)])
        for which, name in ((0, "gen"), (1, "critic")):
            m, v = self._adam_tensors(which)
            m.copy_(torch.as_tensor(np.asarray(sd[f"{name}_adam_m"], np.float32)))
            v.copy_(torch.as_tensor(np.asarray(sd[f"{name}_adam_v"], np.float32)))
        self.optimizer.iterations = int(sd["iterations"])
        self.optimizer.lr, self.optimizer.beta_1, self.optimizer.beta_2, self.optimizer.epsilon = (float(x) for x in sd["adam"])
        self._gen.set_state(torch.as_tensor(np.asarray(sd["rng_state"], np.uint8)))
        self._push_counters(self.optimizer.iterations, sd.get("rng_ctr", 0))
        torch.cuda.synchronize(self.ctx.device)

    def save_checkpoint(self, path):
        """One .npz with both nets' weights (Keras order), the Adam moments, the shared step counter and the RNG state."""
        np.savez(path, **self.state_dict())

    def load_checkpoint(self, path):
        with np.load(path) as f:
            self.load_state_dict({k: f[k] for k in f.files})

    def critic_mask_shapes(self, B):
        shapes, chans = [], (64, 128, 256, 256)
        for (_, out, _), ch in zip(W.critic_geometry(self.ctx.nd), chans):
            shapes.append((B,) + tuple(out) + (ch,))
        return shapes

    def draw_masks(self, B):
        """Bernoulli(keep=0.75) dropout masks for one critic invocation (Dropout(0.25), :289-301)."""
        dev = f"cuda:{self.ctx.device}"
        return [(torch.rand(s, device=dev, generator=self._gen) < 0.75).float() for s in self.critic_mask_shapes(B)]

    @staticmethod
    def _maskptrs(masks):
        if masks is None:
            return None
        return (C.c_void_p * 4)(*[C.c_void_p(m.data_ptr()) for m in masks])

    def _apply(self, which):
        """Gradient exchange (SUM over ranks, 1/world folded into Adam) + Keras-Adam update with the shared step counter, which
        lives in device memory (rdg_adam_apply_dev) so that the same call sequence can be captured in a CUDA graph."""
        from .dist import allreduce_sum_
        world = self._world()
        if world > 1:
            g = self.grad_tensor(which)
            if self.profile_comm:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if self.peer_exchange:               # own two-shot all-reduce over NVLink peer memory, stream-ordered, capturable
                _lib.check(self.ctx.lib.rdg_peer_allreduce(self.ctx.handle, which, self.ctx._stream()))
            else:
                allreduce_sum_(g, self.pg)       # 1/world is folded into the fused Adam kernel (grad_scale)
            if self.profile_comm:
                e1.record()
                self._comm_events.append((e0, e1))
        opt = self.optimizer
        opt.iterations += 1
        _lib.check(self.ctx.lib.rdg_adam_apply_dev(self.ctx.handle, which, opt.lr, opt.beta_1, opt.beta_2, opt.epsilon,
                                                   1.0 / world, _lib.MODES[self.gen_mode], self.ctx._stream()))

    # -- device-resident steps (tensor-core mode): noise / alpha / dropout masks drawn on the GPU, nothing read back.
    # Each step is issued in two phases: phase 1 does not read the critic's weights (random draws, generator forward,
    # interpolation, critic inputs) and runs on the caller's stream WHILE a second stream still all-reduces and applies the previous
    # step's critic gradients; phase 2 waits for that update.  A generator update is awaited before anything else runs.
    def _update_stream(self):
        if getattr(self, "_upd_stream", None) is None:
            # the update sits on the critical path between two steps: high priority, like the capture stream of an iteration
            self._upd_stream = torch.cuda.Stream(device=self.ctx.device, priority=self._critical_priority())
            self._pending, self._pending_which = None, None
        return self._upd_stream

    @staticmethod
    def _critical_priority():
        """Stream priority of the critical chain (critic passes, updates): -3; the filter-gradient side streams of the library
        run at -2, the critic steps' prefetch stream at -1, the generator step's forward at 0 (default, lowest) -- in the order
        in which their results are needed, so that when several have thread blocks pending the SMs go to the launch the others
        wait for.  Measured 4.43 -> 4.32 ms per iteration (chain above everything else; the grading below it adds 0.01 ms).
        RDG_STREAM_PRIORITY=0 puts the Python-side streams on the default priority (A/B timing)."""
        return -3 if os.environ.get("RDG_STREAM_PRIORITY", "1") == "1" else 0

    def finish(self):
        """Make the caller's stream wait for the last overlapped update (call before reading weights / losses on another path)."""
        if getattr(self, "_pending", None) is not None:
            torch.cuda.current_stream(self.ctx.device).wait_event(self._pending)
            self._pending, self._pending_which = None, None

    def _launch_update(self, which):
        main, upd = torch.cuda.current_stream(self.ctx.device), self._update_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        upd.wait_event(ev)
        with torch.cuda.stream(upd):
            self._apply(which)
            self._pending = torch.cuda.Event()
            self._pending.record(upd)
        self._pending_which = which

    def _critic_calls(self, x_real, cond, losses_out, slot=0):
        """slot: which of the context's two sets of step buffers the step uses (RDG_STEP_SLOT1 = 16 in `phases`)."""
        ctx = self.ctx
        args = (ctx.handle, C.c_void_p(x_real.data_ptr()), C.c_void_p(cond.data_ptr()), int(x_real.shape[0]),
                _lib.MODES[self.gen_mode], self.seed + self.rank, int(self.dropout), C.c_void_p(losses_out.data_ptr()))
        return [lambda ph=ph: _lib.check(ctx.lib.rdg_critic_step_dev(*args, ph | (16 if slot else 0), ctx._stream())) for ph in (1, 2)]

    def _generator_calls(self, cond, loss_out):
        ctx = self.ctx
        args = (ctx.handle, C.c_void_p(cond.data_ptr()), int(cond.shape[0]), self.seed + self.rank, int(self.dropout),
                C.c_void_p(loss_out.data_ptr()))
        return [lambda ph=ph: _lib.check(ctx.lib.rdg_generator_step_dev(*args, ph, ctx._stream())) for ph in (1, 2)]

    def _run_step(self, which, phase1, phase2):
        """phase1 / phase2: callables that enqueue the two phases on the current stream (C calls or graph replays)."""
        self._set_mode()
        self._update_stream()
        if self._pending_which == self.WHICH_GEN:
            self.finish()                   # everything reads the generator: wait for its update first
        phase1()
        self.finish()                       # phase 2 reads the critic's weights and overwrites its gradient buffer
        phase2()
        self._launch_update(which)

    def critic_step_device(self, x_real, cond, losses_out):
        """One critic step (gradients + exchange + update) on device tensors; losses_out: cuda float32 [4]."""
        self._run_step(self.WHICH_CRITIC, *self._critic_calls(x_real, cond, losses_out))

    def generator_step_device(self, cond, loss_out):
        self._run_step(self.WHICH_GEN, *self._generator_calls(cond, loss_out))

    def _run_iteration(self, critic_calls, gen_calls):
        """One iteration = the critic steps, then the generator step (reference :468-482).  Phase 1 of the generator step (noise,
        generator forward with saved activations) reads neither the critic nor anything the critic steps write, so it is issued
        on a third stream at the START of the iteration and runs next to the critic steps, whose launches are too small to fill the
        GPU; the generator step's second phase (critic pass + backward) follows the last critic update as before.
        The critic steps' own first phases (draws, frozen generator forward, interpolation: they read only the generator, which
        the critic steps do not change) run on a prefetch stream one step ahead: the steps alternate between the context's two
        buffer sets, so phase 1 of step k+1 overlaps phase 2 of step k and only phase 2 stays on the critical path."""
        dev = self.ctx.device
        main = torch.cuda.current_stream(dev)
        self._set_mode()
        self._update_stream()
        if getattr(self, "_gen_stream", None) is None:
            self._gen_stream = torch.cuda.Stream(device=dev)
        if self._pending_which == self.WHICH_GEN:
            self.finish()                   # the previous iteration's generator update
        if os.environ.get("RDG_GEN_OVERLAP", "1") == "0":          # A/B switch: everything in stream order
            for p1, p2 in critic_calls:
                self._run_step(self.WHICH_CRITIC, p1, p2)
            self._run_step(self.WHICH_GEN, gen_calls[0], gen_calls[1])
            return
        ev0 = torch.cuda.Event()
        ev0.record(main)
        self._gen_stream.wait_event(ev0)
        with torch.cuda.stream(self._gen_stream):
            gen_calls[0]()
            ev_g = torch.cuda.Event()
            ev_g.record(self._gen_stream)
        if os.environ.get("RDG_PREFETCH", "1") == "0":             # A/B switch: critic phases in stream order
            for p1, p2 in critic_calls:
                self._run_step(self.WHICH_CRITIC, p1, p2)
        else:
            if getattr(self, "_pre_stream", None) is None:
                self._pre_stream = torch.cuda.Stream(device=dev, priority=-1 if self._critical_priority() else 0)
            pre = self._pre_stream
            pre.wait_event(ev0)
            slot_free, ready = {}, {}

            def prefetch(k):
                with torch.cuda.stream(pre):
                    if k % 2 in slot_free:
                        pre.wait_event(slot_free[k % 2])       # phase 2 of step k-2 has finished with this buffer set
                    critic_calls[k][0]()
                    ready[k] = torch.cuda.Event()
                    ready[k].record(pre)

            prefetch(0)
            for k, (_, p2) in enumerate(critic_calls):
                if k + 1 < len(critic_calls):
                    prefetch(k + 1)
                main.wait_event(ready[k])
                self._run_step(self.WHICH_CRITIC, lambda: None, p2)
                slot_free[k % 2] = torch.cuda.Event()
                slot_free[k % 2].record(main)
        main.wait_event(ev_g)
        self._run_step(self.WHICH_GEN, lambda: None, gen_calls[1])

    def iteration_device(self, x_real, cond, cond_gen, d_losses, g_loss):
        """x_real [n_critic,B,24,nd,nd,1], cond [n_critic,B,nd,nd,ncond], cond_gen [B,nd,nd,ncond], d_losses [n_critic,4], g_loss [1]:
        device tensors.  Issues one whole iteration (see _run_iteration); call finish() before reading weights."""
        n = int(x_real.shape[0])
        self._run_iteration([self._critic_calls(x_real[k], cond[k], d_losses[k], slot=k % 2) for k in range(n)],
                            self._generator_calls(cond_gen, g_loss))

    def capture_iteration(self, batch, n_critic=5, segmented=None, preserve_state=False):
        """Capture one training iteration (n_critic critic steps + 1 generator step, reference :468-482, with the gradient
        exchange and the Adam updates) in a CUDA graph.  Returns an IterationGraph: fill `x_real` [n_critic,B,24,nd,nd,1],
        `cond` [n_critic,B,nd,nd,ncond] and `cond_gen` [B,nd,nd,ncond] (static device tensors) and call replay();
        `d_losses` [n_critic,4] and `g_loss` [1] are device tensors read whenever the caller wants (no sync per step).
        segmented (default: data-parallel over NCCL): one graph per step phase with the gradient exchange issued between them;
        with the peer-memory exchange (`peer_exchange`) a data-parallel iteration is one graph like the single-GPU one.
        The capture is preceded by one REAL iteration on constant placeholder data (it sizes every workspace outside the capture);
        preserve_state=True puts weights, Adam moments, step and random counters back afterwards, so that a training run can
        capture at its start without taking a step on the placeholders.  Under data parallelism the call is collective (the
        placeholder iteration exchanges gradients): every rank must make it, with the same arguments."""
        if self.train_mode != "tf32":
            raise RuntimeError("capture_iteration needs train_mode='tf32' (device-resident step inputs)")
        dev = f"cuda:{self.ctx.device}"
        nd, nc = self.ctx.nd, self.ctx.ncond
        ig = IterationGraph()
        ig.n_critic, ig.trainer = n_critic, self
        ig.x_real = torch.zeros((n_critic, batch, W.NHOURS, nd, nd, 1), device=dev)
        ig.cond = torch.zeros((n_critic, batch, nd, nd, nc), device=dev)
        ig.cond_gen = torch.zeros((batch, nd, nd, nc), device=dev)
        ig.d_losses = torch.zeros((n_critic, 4), device=dev)
        ig.g_loss = torch.zeros((1,), device=dev)
        ig.x_real[:] = 1.0 / W.NHOURS

        def body():
            self.iteration_device(ig.x_real, ig.cond, ig.cond_gen, ig.d_losses, ig.g_loss)
            self.finish()               # the iteration ends with the generator update joined back

        self.finish()
        saved = self.state_dict() if preserve_state else None
        it0 = self.optimizer.iterations
        s = torch.cuda.Stream(device=dev, priority=self._critical_priority())       # kernel nodes inherit the capturing stream's priority
        s.wait_stream(torch.cuda.current_stream(self.ctx.device))
        with torch.cuda.stream(s):
            body()                      # eager pass: allocations, attribute calls and weight images happen outside the capture
        torch.cuda.current_stream(self.ctx.device).wait_stream(s)
        torch.cuda.synchronize(self.ctx.device)
        if segmented is None:
            segmented = self._world() > 1 and not self.peer_exchange      # NCCL calls stay outside the graphs
        if not segmented:
            ig.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ig.graph, stream=s):
                body()
            # the eager pass was a real iteration; the capture itself launches nothing, only the host mirror moved
            self.optimizer.iterations = it0 + n_critic + 1
        else:
            # data-parallel: the NCCL exchange stays outside the graphs (one graph per step phase; the all-reduce + Adam of a
            # step are issued eagerly on the update stream between them and overlap the next step's first phase)
            ig.graph = None
            ig.segments = []
            calls = [(self.WHICH_CRITIC, self._critic_calls(ig.x_real[k], ig.cond[k], ig.d_losses[k], slot=k % 2)) for k in range(n_critic)]
            calls.append((self.WHICH_GEN, self._generator_calls(ig.cond_gen, ig.g_loss)))
            for which, (p1, p2) in calls:
                gs = []
                for fn in (p1, p2):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=s if fn is p2 else None):       # first phases run off the critical chain
                        fn()
                    gs.append(g)
                ig.segments.append((which, gs[0], gs[1]))
        torch.cuda.synchronize(self.ctx.device)
        if saved is not None:
            self.load_state_dict(saved)
        return ig

    # -- the two train_on_batch calls
    def critic_grads(self, x_real, cond, latent, alpha=None, masks3="draw"):
        """Losses + gradients of one critic step, no update.  masks3: "draw" | None | (mf, mr, mh)."""
        ctx = self.ctx
        self.finish()
        xr, c, z = ctx.dev(x_real), ctx.dev(cond), ctx.dev(latent)
        B = int(xr.shape[0])
        if alpha is None:
            alpha = torch.rand(B, device=xr.device, generator=self._gen)      # tf.random.uniform((batch_size,1,1,1,1)) :223
        al = ctx.dev(alpha).reshape(-1)
        if isinstance(masks3, str):
            masks3 = (self.draw_masks(B), self.draw_masks(B), self.draw_masks(B))
        mf, mr, mh = (None, None, None) if masks3 is None else [[ctx.dev(m) for m in ms] for ms in masks3]
        losses = torch.empty(4, device=xr.device, dtype=torch.float32)
        self._keep = (xr, c, z, al, mf, mr, mh)
        self._set_mode()
        _lib.check(ctx.lib.rdg_critic_step_grads(
            ctx.handle, C.c_void_p(xr.data_ptr()), C.c_void_p(c.data_ptr()), C.c_void_p(z.data_ptr()),
            C.c_void_p(al.data_ptr()), self._maskptrs(mf), self._maskptrs(mr), self._maskptrs(mh), B,
            _lib.MODES[self.gen_mode], C.c_void_p(losses.data_ptr()), ctx._stream()))
        return losses

    def critic_train_on_batch(self, inputs, targets=None, alpha=None, masks3="draw"):
        """critic_model.train_on_batch([X_real, cond_real, latent], [valid, fake, dummy]) ->
        [total, l_valid, l_fake, l_gp]  (gan_train_cwgangp_pixelnorm.py:472).  The targets are the
        reference's constants (-1, +1, 0; :452-454) and are accepted only for signature parity."""
        x_real, cond, latent = inputs
        losses = self.critic_grads(x_real, cond, latent, alpha, masks3)
        self._apply(self.WHICH_CRITIC)
        return [float(v) for v in losses.cpu().numpy()]

    def generator_grads(self, latent, cond, masks="draw"):
        ctx = self.ctx
        self.finish()
        z, c = ctx.dev(latent), ctx.dev(cond)
        B = int(z.shape[0])
        if isinstance(masks, str):
            masks = self.draw_masks(B)
        m = None if masks is None else [ctx.dev(x) for x in masks]
        loss = torch.empty(1, device=z.device, dtype=torch.float32)
        self._keep = (z, c, m)
        self._set_mode()
        _lib.check(ctx.lib.rdg_generator_step_grads(ctx.handle, C.c_void_p(z.data_ptr()), C.c_void_p(c.data_ptr()),
                                                    self._maskptrs(m), B, C.c_void_p(loss.data_ptr()), ctx._stream()))
        return loss

    def critic_input_gradient(self, sample, cond):
        """d sum_b D(sample_b, cond_b) / d sample without dropout -> (B,24,nd,nd,1) numpy: the gradient GradientPenalty takes the
        norm of (gan_train_cwgangp_pixelnorm.py:238-241), from rdg_critic_input_grad (critic forward + transposed-conv chain)."""
        ctx = self.ctx
        self.finish()
        x, c = ctx.dev(sample), ctx.dev(cond)
        B = int(x.shape[0])
        g = torch.empty((B, W.NHOURS, ctx.nd, ctx.nd), device=x.device, dtype=torch.float32)
        _lib.check(ctx.lib.rdg_critic_input_grad(ctx.handle, C.c_void_p(x.data_ptr()), C.c_void_p(c.data_ptr()), None,
                                                 C.c_void_p(g.data_ptr()), B, ctx._stream()))
        torch.cuda.synchronize(ctx.device)
        return g.cpu().numpy().reshape(B, W.NHOURS, ctx.nd, ctx.nd, 1)

    def generator_train_on_batch(self, inputs, target=None, masks="draw"):
        """generator_model.train_on_batch([latent, cond], valid) -> g_loss (:482)."""
        latent, cond = inputs
        loss = self.generator_grads(latent, cond, masks)
        self._apply(self.WHICH_GEN)
        return float(loss.item())

    def comm_ms(self):
        """Sum of the all-reduce durations recorded while `profile_comm` was set (device time), then reset."""
        torch.cuda.synchronize(self.ctx.device)
        ms = sum(a.elapsed_time(b) for a, b in self._comm_events)
        self._comm_events = []
        return ms


class IterationGraph:
    """A captured training iteration (GanTrainer.capture_iteration)."""

    def replay(self):
        t = self.trainer
        if self.graph is not None:
            self.graph.replay()
            t.optimizer.iterations += self.n_critic + 1
            return
        # _apply (all-reduce + Adam) advances the host step mirror itself
        t._run_iteration([(g1.replay, g2.replay) for w, g1, g2 in self.segments if w == t.WHICH_CRITIC],
                         [(g1.replay, g2.replay) for w, g1, g2 in self.segments if w == t.WHICH_GEN][0])
        t.finish()
