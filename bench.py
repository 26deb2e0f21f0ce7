#!/usr/bin/env python
"""bench.py -- scenarios/s of the RainDisaggGAN generator forward (BASELINE.json metric).

Workload (config #2): 10 000 synthetic gamma(0.8,12) daily-sum 16x16 conditions x 100
scenarios each = 1 000 000 scenarios per step and per GPU (weak scaling: every rank runs its own
full ensemble, no data-path collective), seeded random-init generator of the reference architecture.

  value  : scenarios/s with latent + conditions resident in HBM, output written to HBM
  e2e    : same metric through the host-buffer C-ABI call rdg_generate_host (pinned host latent in,
           pinned host mm/h fields out, copies inside the timed region)
  roofline: dominant kernel (tc_upconv64_planes_kernel: tcgen05 upsample-folded conv 128->64 + PixelNorm +
           LeakyReLU + fused output-conv tap products), CUDA-event duration on the launch stream inside the
           timed steps, algorithmic FLOPs; `traffic` = DRAM bytes of one launch from the committed
           ncu --set full capture (profiles/r1h_ncu_planes_cg2_summary.md), scaled to the units of a launch
  cpu_baseline / --impl reference: BASELINE.md section 3: `import tensorflow` + the shipped .h5 if both exist (kind "tf"),
           else the oracle (torch-CPU restatement of the Keras graph, kind "port") on a bounded sample of the same workload;
           the cpu_baseline sample is the FIRST 32 conditions x 100 scenarios of the timed workload with the same latent
           noise, and its output is also the parity check of the timed GPU output (`parity`).
  train / dp: one cWGAN-GP iteration (5 critic + 1 generator step, batch 32 per GPU) replayed from a CUDA graph in the
           tensor-core training mode; at N > 1 the N-rank averaged gradients are compared with rank 0 running the gathered batch.
The JSON line is ordered so that `e2e_stats`, `train` and `dp` come last (the driver keeps the tail of stdout).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))

METRIC = "scenarios/sec (16x16x24h cGAN gen)"
N_COND, SCEN_PER_COND = 10000, 100
# algorithmic work per scenario: MAC counts of the reference graph
MAC_TOTAL = 2_220_011_520
MAC_CONV3 = 1_358_954_496          # Conv3D(128->64) on the 24x16x16 grid, direct form
MAC_CONV3_FOLDED = 402_653_184     # what the folded kernel executes
MAC_TOTAL_FOLDED = 666_021_888     # whole generator with the three upsample folds
# dram__bytes_read.sum + dram__bytes_write.sum of one 4736-unit launch of the dominant kernel (profiles/r1h_ncu_planes_cg2_summary.md)
CONV3_DRAM_BYTES_PER_UNIT = (932.31e6 + 111.21e6) / 4736


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synth_conditions(n_cond, nd, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    cond_mm = np.clip(rng.gamma(shape=0.8, scale=12.0, size=(n_cond, nd, nd, 1)), 0, 200).astype(np.float32)
    return cond_mm / np.float32(127.4)


def cpu_reference_rate(n_cond, spc, threads, reps=1, latent=None, cond=None, keep=False):
    """Oracle (torch-CPU FP32 restatement of gen.predict + rescale) on n_cond x spc scenarios.  latent [n_cond*spc,100] / cond
    [n_cond,16,16,1] given: that sample (the head of the timed GPU workload); keep: also return the mm/h fields."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import torch
    import rdg_oracle as O
    from rdg_b200 import weights as W
    torch.set_num_threads(threads)
    gw = W.init_generator_weights(0)
    if cond is None:
        cond = synth_conditions(n_cond, 16, 354)
    rng = np.random.default_rng(1)
    best, fields = None, None
    for _ in range(reps):
        outs = []
        t0 = time.perf_counter()
        for i in range(n_cond):   # one predict per condition, like generate_and_evaluate_crps.py:177-188
            z = rng.standard_normal((spc, 100)).astype(np.float32) if latent is None else latent[i * spc:(i + 1) * spc]
            cb = np.repeat(cond[i:i + 1], spc, axis=0)
            out = O.generator_forward(gw, z, cb, torch.float32)
            mm = out[..., 0] * cond[i, :, :, 0] * np.float32(127.4)
            if keep:
                outs.append(mm)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        fields = outs
    if keep:
        return n_cond * spc / best, best, np.concatenate(fields)
    return n_cond * spc / best, best


def tf_reference_rate(n_cond, spc):
    """the reference's own TensorFlow generate_scenarios, if TensorFlow and the shipped
    trained_models/*.h5 exist on this box (they do not in the build image; the attempt is made at every run)."""
    try:
        import tensorflow  # noqa: F401
    except Exception as e:          # ImportError, or a broken install
        return None, f"tensorflow not importable ({type(e).__name__})"
    import glob
    import numpy as np
    ref_dirs = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]
    for d in ref_dirs:
        if glob.glob(os.path.join(d, "trained_models", "gen_*.h5")) and os.path.exists(os.path.join(d, "raindisagg_gan_pretrained.py")):
            cwd = os.getcwd()
            try:
                os.chdir(d)
                sys.path.insert(0, d)
                import raindisagg_gan_pretrained as ref
                cond = synth_conditions(n_cond, 16, 354) * np.float32(127.4)
                ref.generate_scenarios(cond[0], spc)
                t0 = time.perf_counter()
                for i in range(n_cond):
                    ref.generate_scenarios(cond[i], spc)
                dt = time.perf_counter() - t0
                return n_cond * spc / dt, dt
            finally:
                os.chdir(cwd)
    return None, "tensorflow imports but trained_models/gen_*.h5 is not shipped"


def workload_config(n_cond, spc):
    """config block shared by both arms: the reference arm times a bounded sample of the SAME workload."""
    return {"workload": f"ensemble generation: {n_cond} synthetic gamma(0.8,12) daily-sum 16x16 conditions x "
                        f"{spc} scenarios per GPU, seeded random-init generator (pretrained .h5 not shipped)",
            "scenarios_per_step_per_gpu": n_cond * spc}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_cond = args.ref_conditions
    times = []
    tf_rate, tf_note = tf_reference_rate(1, SCEN_PER_COND)
    kind = "tf" if tf_rate is not None else "port"
    for i in range(args.warmup + args.steps):
        if kind == "tf":
            rate, dt = tf_reference_rate(n_cond, SCEN_PER_COND)
        else:
            rate, dt = cpu_reference_rate(n_cond, SCEN_PER_COND, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = n_cond * SCEN_PER_COND / (ms / 1e3)
    sample = f"{n_cond} conditions x {SCEN_PER_COND} scenarios per step (of {N_COND} x {SCEN_PER_COND})"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "scenarios/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.conditions, args.scen_per_cond), sample=sample,
                           note="each step is a bounded sample of the workload; rate = sample scenarios / sample time"),
            "cpu_baseline": {"value": value, "unit": "scenarios/s", "cores": threads, "kind": kind, "sample": sample,
                             "note": "reference TensorFlow generate_scenarios" if kind == "tf" else
                                     f"torch-CPU FP32 restatement of the Keras generator ({tf_note})"},
            "e2e": {"value": value, "unit": "scenarios/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default=os.environ.get("RDG_MODE", "fp16"), choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--conditions", type=int, default=N_COND)
    ap.add_argument("--scen-per-cond", type=int, default=SCEN_PER_COND)
    ap.add_argument("--chunk", type=int, default=0, help="samples per internal pass (0 = library default)")
    ap.add_argument("--ref-conditions", type=int, default=10, help="conditions per step of the CPU reference arm")
    ap.add_argument("--cpu-conditions", type=int, default=32, help="conditions of the cpu_baseline / parity sample (x100 scenarios)")
    ap.add_argument("--no-extras", action="store_true", help="skip the bf16 rate and the large-domain block")
    ap.add_argument("--e2e-group", type=int, default=1000, help="conditions per host-API call in the e2e leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind each rank to its GPU's NUMA node")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-iteration timing")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from rdg_b200 import _lib, weights as W
    from rdg_b200.engine import Context, Generator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    from rdg_b200.dist import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local) if world > 1 and not args.no_numa else None   # local pinned staging for the e2e legs

    n_cond, spc = args.conditions, args.scen_per_cond
    B = n_cond * spc
    ctx = Context(16, 1, device=local, max_chunk=args.chunk)
    gen = Generator(W.init_generator_weights(0), ctx=ctx, mode=args.mode)
    lib = ctx.lib

    # inputs resident in HBM: each rank has its own conditions and its own Philox stream
    cond_h = synth_conditions(n_cond, 16, 354 + rank)
    cond = torch.as_tensor(cond_h, device=dev)
    latent = torch.empty((B, 100), device=dev, dtype=torch.float32)
    _lib.check(lib.rdg_fill_normal(C.c_void_p(latent.data_ptr()), B * 100, 1234 + rank, 0, ctx._stream()))
    out = torch.empty((B, 24, 16, 16), device=dev, dtype=torch.float32)
    torch.cuda.synchronize()

    def step():
        gen.forward_device(latent, cond, scen_per_cond=spc, mode=args.mode, out_mm=True, out=out, check=False)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.rdg_profile_enable(ctx.handle, 1)
    l0 = lib.rdg_launch_count(ctx.handle)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    lib.rdg_profile_enable(ctx.handle, 0)
    launches = lib.rdg_launch_count(ctx.handle) - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    ms_sum = (C.c_double * 8)()
    n_l = (C.c_longlong * 8)()
    n_u = (C.c_longlong * 8)()
    _lib.check(lib.rdg_profile_collect(ctx.handle, ms_sum, n_l, n_u))
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)

    # conservation check on the last step's output (daily sum reproduces the condition)
    chk = out[:spc * 4].sum(dim=1)
    want = (cond[:4, :, :, 0] * 127.4).repeat_interleave(spc, dim=0)
    cons = float(((chk - want).abs() / want.clamp_min(1e-6)).max().item())
    n_par = min(n_cond, args.cpu_conditions)
    head_out = out[:n_par * spc].cpu().numpy()          # first conditions of the LAST TIMED step, checked against the oracle below
    head_lat = latent[:n_par * spc].cpu().numpy()

    # ---- end-to-end through the host-buffer C-ABI call
    e2e = None
    if not args.no_e2e:
        # pinned host latent for the whole step; results stream through a reusable pinned buffer in groups of
        # conditions (the reference's loop also consumes each condition's ensemble before the next,
        # generate_and_evaluate_crps.py:177-195) so host memory stays bounded at any rank count
        grp_cond = min(n_cond, args.e2e_group)
        grp = grp_cond * spc
        lat_h = torch.empty((B, 100), dtype=torch.float32, pin_memory=True)
        lat_h.copy_(latent)
        out_h = torch.empty((grp, 24, 16, 16), dtype=torch.float32, pin_memory=True)
        cond_p = torch.as_tensor(cond_h).pin_memory()
        torch.cuda.synchronize()

        def e2e_step():
            for c0 in range(0, n_cond, grp_cond):
                c1 = min(n_cond, c0 + grp_cond)
                gen.generate_ensemble_host(lat_h[c0 * spc:c1 * spc], cond_p[c0:c1], spc, out=out_h[:(c1 - c0) * spc],
                                           mode=args.mode, out_mm=True)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B / float(t.item()), "unit": "scenarios/s",
               "h2d_bytes_per_step": int(lat_h.numel() * 4 + cond_p.numel() * 4),
               "d2h_bytes_per_step": int(B * 24 * 16 * 16 * 4), "steps": n_e2e,
               "api": f"rdg_generate_host, pinned buffers, {grp_cond} conditions per call"}
        # the same call with a PAGEABLE result buffer (what gen.predict hands over): internal pinned staging ring + host callbacks
        n_pg = min(n_cond, 2 * grp_cond)
        out_pg = np.empty((grp, 24, 16, 16), np.float32)
        gen.generate_ensemble_host(lat_h[:grp], cond_p[:grp_cond], spc, out=out_pg, mode=args.mode, out_mm=True)
        barrier()
        t0 = time.perf_counter()
        for c0 in range(0, n_pg, grp_cond):
            gen.generate_ensemble_host(lat_h[c0 * spc:(c0 + grp_cond) * spc], cond_p[c0:c0 + grp_cond], spc, out=out_pg, mode=args.mode, out_mm=True)
        barrier()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e["pageable_out_value"] = world * n_pg * spc / float(t.item())
        c_last = max(0, n_pg - grp_cond)
        assert np.array_equal(out_pg[:1000], out[c_last * spc:c_last * spc + 1000].cpu().numpy()), "pageable e2e output differs"
        del out_pg
        # the last group's e2e result equals the device-resident result
        last0 = ((n_cond - 1) // grp_cond) * grp_cond * spc
        assert torch.equal(out_h[:1000], out[last0:last0 + 1000].cpu()), "e2e output differs from device-resident output"
        del out_h, lat_h

    # ---- the same ensemble workload through the statistics API: fields stay on the GPU, only the
    # area means [B,24] and the area-mean CRPS [n_cond,24] come back (generate_and_evaluate_crps.py:177-191 per condition)
    e2e_stats = None
    if not args.no_e2e:
        grp_cond = min(n_cond, args.e2e_group)
        lat_h = torch.empty((B, 100), dtype=torch.float32, pin_memory=True)
        lat_h.copy_(latent)
        cond_p = torch.as_tensor(cond_h).pin_memory()
        rng = np.random.default_rng(99 + rank)
        obs_frac = rng.random((grp_cond, 24, 16, 16), dtype=np.float32)
        obs_frac /= obs_frac.sum(axis=1, keepdims=True)
        obs_p = torch.as_tensor(obs_frac).pin_memory()      # synthetic observations, one block reused by every group
        torch.cuda.synchronize()

        def stats_step():
            last = None
            for c0 in range(0, n_cond, grp_cond):
                c1 = min(n_cond, c0 + grp_cond)
                last = gen.generate_ensemble_stats_host(lat_h[c0 * spc:c1 * spc], cond_p[c0:c1], spc, obs_p[:c1 - c0],
                                                        mode=args.mode, out_mm=True)
            return last

        stats_step()
        barrier()
        t0 = time.perf_counter()
        am, cr = stats_step()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert np.isfinite(am).all() and np.isfinite(cr).all()
        e2e_stats = {"value": world * B / float(t.item()), "unit": "scenarios/s",
                     "h2d_bytes_per_step": int(lat_h.numel() * 4 + cond_p.numel() * 4 + n_cond * 24 * 256 * 4),
                     "d2h_bytes_per_step": int(B * 24 * 4 + n_cond * 24 * 4),
                     "api": "rdg_generate_stats_host (area means + CRPS on device)"}
        del lat_h

    # ---- training (configs #3/#4): one cWGAN-GP iteration = 5 critic steps + 1 generator step, batch 32 per GPU, data-parallel over
    # the ranks; tensor-core training mode (tcgen05 kind::tf32), the whole iteration replayed from one CUDA graph
    train = dp = None
    if not args.no_train:
        from rdg_b200.engine import Critic, GanTrainer
        tctx = Context(16, 1, device=local, max_chunk=1024)
        tgen = Generator(W.init_generator_weights(0), ctx=tctx, mode="fp16")
        tcrit = Critic(W.init_critic_weights(1), ctx=tctx)
        TB = 32
        rng = np.random.default_rng(7 + rank)
        lg = rng.standard_normal((5, TB, 24, 16, 16, 1)).astype(np.float32) * 2
        ex = np.exp(lg - lg.max(axis=2, keepdims=True))
        x_real = tctx.dev((ex / ex.sum(axis=2, keepdims=True)).astype(np.float32))
        tcond = tctx.dev(synth_conditions(5 * TB, 16, 1000 + rank).reshape(5, TB, 16, 16, 1))

        def timed(fn, n):
            for _ in range(3):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b) / n], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        tr = GanTrainer(tgen, tcrit, gen_mode="fp16", seed=100, train_mode="tf32")
        ig = tr.capture_iteration(TB)
        ig.x_real.copy_(x_real); ig.cond.copy_(tcond); ig.cond_gen.copy_(tcond[0])
        ms_graph = timed(ig.replay, 20)
        losses, gl = ig.d_losses.cpu().numpy(), float(ig.g_loss.item())
        dl, gls = torch.zeros((5, 4), device=dev), torch.zeros(1, device=dev)
        ms_crit = timed(lambda: (tr.critic_step_device(x_real[0], tcond[0], dl[0]), tr.finish()), 20)
        ms_gen = timed(lambda: (tr.generator_step_device(tcond[0], gls), tr.finish()), 20)
        tg = torch.Generator(device=dev); tg.manual_seed(5 + rank)
        tr32 = GanTrainer(tgen, tcrit, gen_mode="fp16", seed=100, train_mode="fp32")

        def keras_iteration():
            for k in range(5):
                tr32.critic_train_on_batch([x_real[k], tcond[k], torch.randn((TB, 100), device=dev, generator=tg)])
            tr32.generator_train_on_batch([torch.randn((TB, 100), device=dev, generator=tg), tcond[0]])
        ms_fp32 = timed(keras_iteration, 3)
        flop_iter = 40.2e9 * TB          # direct-form work of one iteration per sample
        train = {"ms_per_iteration": ms_graph, "samples_per_s": world * TB / (ms_graph * 1e-3), "critic_step_ms": ms_crit,
                 "generator_step_ms": ms_gen, "fp32_simt_ms_per_iteration": ms_fp32,
                 "direct_form_tflops_per_gpu": flop_iter / (ms_graph * 1e-3) / 1e12,
                 "mode": "tcgen05 kind::tf32 (3xTF32 forward), CUDA-graph replay, batch 32/GPU, device Philox noise/alpha/dropout",
                 "parity": "gradients <= 1e-2 rel. L2 per tensor vs the FP64 oracle (tests/test_gpu_train_tc.py; measured <= 7.4e-3)",
                 "finite": bool(np.isfinite(losses).all() and np.isfinite(gl))}
        if world > 1:
            # data-parallel parity: N-rank averaged gradients == rank 0 on the gathered N*32 batch, FP32 mode,
            # explicit inputs, no dropout; and the share of the gradient exchange in an eagerly issued iteration
            z = torch.randn((TB, 100), device=dev, generator=tg)
            al = torch.rand((TB,), device=dev, generator=tg)
            errs = {}
            for which, name in ((1, "critic"), (0, "generator")):
                if which == 1:
                    tr32.critic_grads(x_real[0], tcond[0], z, al, None)
                else:
                    tr32.generator_grads(z, tcond[0], None)
                g = tr32.grad_tensor(which).clone()
                dist.all_reduce(g)
                g /= world
                parts = [torch.empty_like(t) for t in (x_real[0], tcond[0], z, al) for _ in range(world)]
                gathered = []
                for i, t in enumerate((x_real[0], tcond[0], z, al)):
                    lst = parts[i * world:(i + 1) * world]
                    dist.all_gather(lst, t.contiguous())
                    gathered.append(torch.cat(lst))
                if rank == 0:
                    if which == 1:
                        tr32.critic_grads(gathered[0], gathered[1], gathered[2], gathered[3], None)
                    else:
                        tr32.generator_grads(gathered[2], gathered[1], None)
                    ref = tr32.grad_tensor(which)
                    errs[name] = float(((g - ref).norm() / ref.norm()).item())
                    # the same sum taken rank batch by rank batch on ONE GPU (identical kernel configuration per batch, so no
                    # LeakyReLU unit sits on the other side of its kink): isolates the exchange itself
                    acc = torch.zeros_like(g)
                    for r_ in range(world):
                        sl = slice(r_ * TB, (r_ + 1) * TB)
                        if which == 1:
                            tr32.critic_grads(gathered[0][sl], gathered[1][sl], gathered[2][sl], gathered[3][sl], None)
                        else:
                            tr32.generator_grads(gathered[2][sl], gathered[1][sl], None)
                        acc += tr32.grad_tensor(which)
                    acc /= world
                    errs[name + "_rankwise"] = float(((g - acc).norm() / acc.norm()).item())
                barrier()
            tr.profile_comm = True
            tr.comm_ms()
            ms_eager = timed(lambda: (tr.iteration_device(x_real, tcond, tcond[0], dl, gls), tr.finish()), 10)
            comm = tr.comm_ms() / 13.0
            tr.profile_comm = False
            ms_nccl = None
            if tr.peer_exchange:             # the same iteration with the exchange on NCCL (one graph per step phase), for comparison
                os.environ["RDG_PEER_ALLREDUCE"] = "0"
                tr_n = GanTrainer(tgen, tcrit, gen_mode="fp16", seed=100, train_mode="tf32")
                os.environ.pop("RDG_PEER_ALLREDUCE")
                ig_n = tr_n.capture_iteration(TB)
                ig_n.x_real.copy_(x_real); ig_n.cond.copy_(tcond); ig_n.cond_gen.copy_(tcond[0])
                ms_nccl = timed(ig_n.replay, 20)
                tr_n.finish()
                w_, t_ = C.c_int(0), C.c_int(0)
                _lib.check(tctx.lib.rdg_peer_status(tctx.handle, C.byref(w_), C.byref(t_)))
            dp = {"ranks": world, "grad_parity_rel_l2": errs,
                  "grad_parity_note": "N-rank averaged gradients vs rank 0 on the gathered N*32 batch (FP32 mode); *_rankwise: vs rank 0 summing the same "
                                      "rank batches one by one (same split-K configuration, no LeakyReLU kink crossings)", "allreduce_ms_per_iteration": comm,
                  "allreduce_share_of_eager_iteration": comm / ms_eager,
                  "exchange": ("own two-shot all-reduce over NVLink peer memory (CUDA IPC, csrc/dp_peer.cu) + Adam on an update stream, captured "
                               "inside the ONE iteration graph") if tr.peer_exchange else
                              ("NCCL all-reduce of the flat FP32 gradient buffer + Adam on an update stream between per-phase graphs, "
                               "overlapping the next step's generator forward"),
                  "peer_exchange": bool(tr.peer_exchange), "ms_per_iteration_nccl_between_phase_graphs": ms_nccl,
                  "peer_barrier_timeouts": int(t_.value) if tr.peer_exchange else None}
        # critic scoring (config #3, forward only): synthetic hourly fraction fields, tensor-core scoring mode vs the FP32 path
        CB = 20000
        cx = torch.rand((CB, 24, 16, 16), device=dev); cx = cx / cx.sum(dim=1, keepdim=True)
        cc = torch.as_tensor(synth_conditions(CB, 16, 77 + rank), device=dev)
        critic_rates = {}
        for cmode in ("fp16", "fp32"):
            for _ in range(2):
                tcrit.forward_device(cx, cc, mode=cmode)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(3):
                tcrit.forward_device(cx, cc, mode=cmode)
            c1.record(); torch.cuda.synchronize()
            critic_rates[cmode] = round(CB / (c0.elapsed_time(c1) / 3 / 1e3))
        train["critic_scoring_samples_per_s_per_gpu"] = critic_rates
        del cx, cc, ig
        tctx.close()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- config #1 (example.py:8-11): cond = 10 mm/day everywhere, n_scenarios = 10, through the Keras-like predict call
    ex_cond = np.repeat((10.0 * np.ones((1, 16, 16, 1)) / 127.4).astype(np.float32), 10, axis=0)
    np.random.seed(354)
    ex_lat = np.random.normal(size=(10, 100)).astype(np.float32)
    for _ in range(5):
        ex_out = gen.predict([ex_lat, ex_cond], mode=args.mode)
    ts = []
    for _ in range(50):
        t0 = time.perf_counter()
        ex_out = gen.predict([ex_lat, ex_cond], mode=args.mode)
        ts.append(time.perf_counter() - t0)
    example = {"workload": "example.py: cond = 10 mm/day, n_scenarios = 10, gen.predict with host arrays",
               "latency_us_median": 1e6 * statistics.median(ts), "scenarios_per_s": 10 / statistics.median(ts),
               "conservation_err": float(np.abs(ex_out.sum(axis=1) - 1).max())}

    # ---- extras: the bf16 operand mode's rate on the same workload, and the large-domain variant (config #5)
    extras = {}
    if not args.no_extras and world == 1:
        for _ in range(2):
            gen.forward_device(latent, cond, scen_per_cond=spc, mode="bf16", out_mm=True, out=out, check=False)
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(2):
            gen.forward_device(latent, cond, scen_per_cond=spc, mode="bf16", out_mm=True, out=out, check=False)
        b1.record(); torch.cuda.synchronize()
        extras["bf16_scenarios_per_s"] = round(B / (b0.elapsed_time(b1) / 2 / 1e3))
        del out, latent
        torch.cuda.empty_cache()
        lctx = Context(64, 1, device=local)
        lgen = Generator(W.init_generator_weights(0, 64), ctx=lctx, mode="fp16")
        ld = {}
        for LB in (32, 1024):
            lc = lctx.dev(synth_conditions(LB, 64, 3))
            lz = torch.randn((LB, 100), device=dev)
            lo = torch.empty((LB, 24, 64, 64), device=dev)
            for _ in range(2):
                lgen.forward_device(lz, lc, mode="fp16", out_mm=True, out=lo, check=False)
            torch.cuda.synchronize()
            b0.record()
            nrep = 3 if LB > 100 else 10
            for _ in range(nrep):
                lgen.forward_device(lz, lc, mode="fp16", out_mm=True, out=lo, check=False)
            b1.record(); torch.cuda.synchronize()
            ms = b0.elapsed_time(b1) / nrep
            ld[f"B{LB}"] = {"scenarios_per_s": round(LB / ms * 1e3, 1), "executed_tflops": round(LB * 2 * 10.845e9 / ms / 1e9, 1)}
        # one cWGAN-GP iteration at nd = 64, batch 32, tensor-core training mode, replayed from a CUDA graph
        from rdg_b200.engine import Critic, GanTrainer
        lcrit = Critic(W.init_critic_weights(1, 64), ctx=lctx)
        ltr = GanTrainer(lgen, lcrit, gen_mode="fp16", seed=3, train_mode="tf32")
        lig = ltr.capture_iteration(32)
        lx = torch.rand((5, 32, 24, 64, 64, 1), device=dev); lx /= lx.sum(dim=2, keepdim=True)
        lig.x_real.copy_(lx); lig.cond.copy_(lctx.dev(synth_conditions(160, 64, 5).reshape(5, 32, 64, 64, 1))); lig.cond_gen.copy_(lig.cond[0])
        lig.replay(); torch.cuda.synchronize()
        b0.record()
        for _ in range(3):
            lig.replay()
        b1.record(); torch.cuda.synchronize()
        ld["train_ms_per_iteration_batch32"] = round(b0.elapsed_time(b1) / 3, 2)
        ld["train_finite"] = bool(torch.isfinite(lig.d_losses).all().item() and torch.isfinite(lig.g_loss).all().item())
        extras["largedomain_nd64"] = ld
        del lig, ltr
        lctx.close()

    burst, sustained, hbm, src = load_peaks()
    conv3_ms = ms_sum[5] / max(1, n_l[5])
    units_per_launch = n_u[5] / max(1, n_l[5])
    direct = units_per_launch * 2 * MAC_CONV3 / (conv3_ms * 1e-3) / 1e12 if conv3_ms > 0 else 0.0
    executed = direct * MAC_CONV3_FOLDED / MAC_CONV3
    issued = executed * (1.0 - 16.0 / 384.0)      # the hour-boundary taps (16 of 384 M-tile steps) are skipped, not issued
    layer_names = ["concat", "dense_front", "cvt16", "upconv256", "upconv128", "upconv64_planes", "softmax_hours", "pixelnorm"]
    shares = {layer_names[i]: round(ms_sum[i] / args.steps, 2) for i in range(8) if n_l[i]}
    roofline = {"bound": "tensor", "kernel": "tc_upconv64_planes_kernel",
                "achieved": issued, "peak": sustained, "unit": "TFLOP/s", "frac": issued / sustained,
                "peak_kind": f"{src} sustained bf16 cuBLAS (burst {burst})", "frac_of_burst": issued / burst,
                "counts": "MACs issued to the tensor pipe (upsample-folded form, skipped boundary taps excluded)",
                "direct_form_tflops": direct, "fold_factor": MAC_CONV3 / MAC_CONV3_FOLDED,
                "ms_per_launch": conv3_ms, "units_per_launch": units_per_launch,
                "traffic": CONV3_DRAM_BYTES_PER_UNIT * units_per_launch,
                "algorithmic_bytes_per_launch": units_per_launch * (12 * 64 * 128 * 2 + 24 * 256 * 4),
                "layer_ms_per_step": shares,
                "whole_path_executed_frac": value / world * 2 * MAC_TOTAL_FOLDED / 1e12 / sustained}

    # ---- CPU baseline + parity: the oracle on the first conditions of the timed workload, same latent noise
    cpu = parity = None
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        tf_rate, tf_note = tf_reference_rate(1, spc)
        rate, dt, ref = cpu_reference_rate(n_par, spc, threads, latent=head_lat, cond=cond_h[:n_par], keep=True)
        cpu = {"value": rate, "unit": "scenarios/s", "cores": threads, "kind": "port",
               "sample": f"first {n_par} conditions x {spc} scenarios ({dt:.1f} s) of the {n_cond} x {spc} workload, same noise as the GPU run",
               "note": f"torch-CPU FP32 restatement (oracle/rdg_oracle.py); {tf_note}"}
        if tf_rate is not None:
            cpu["tf_reference_scenarios_per_s"] = tf_rate
        ok = ref > 1e-12
        rel = np.abs(head_out.astype(np.float64) - ref)[ok] / ref[ok]
        parity = {"checked_scenarios": int(n_par * spc), "max_rel_err": float(rel.max()), "rms_rel_err": float(np.sqrt(np.mean(rel ** 2))),
                  "bound": {"fp16": 1e-2, "bf16": 4e-2, "fp32": 1e-5}[args.mode], "against": "FP32 CPU oracle, first scenarios of the last timed step"}
        parity["ok"] = bool(parity["max_rel_err"] <= parity["bound"])

    line = {"metric": METRIC, "value": value, "unit": "scenarios/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.mode], "data": "synthetic",
            "config": dict(workload_config(n_cond, spc), **{"chunk": ctx.max_chunk, "mode": args.mode,
                       "l2": "inputs (410 MB) and outputs (24.6 GB) per step exceed the 126 MB L2",
                       "parallelism": f"shard{world}-independent", "numa_node_rank0": numa}),
            "gpu_launches": int(launches), "clocks": clocks, "conservation_rel_err": cons, "parity": parity,
            "roofline": roofline, "cpu_baseline": cpu, "example": example, "extras": extras,
            "e2e": e2e, "e2e_stats": e2e_stats, "train": train, "dp": dp}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
