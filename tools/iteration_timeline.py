#!/usr/bin/env python
"""Where one training iteration spends its time on the critical path: CUDA events around every phase of an EAGERLY issued
iteration (5 critic steps + 1 generator step, batch 32, tensor-core mode, all overlap streams active), median over the timed
iterations, in microseconds since the iteration's first launch.  A debugging aid for the stream layout of
GanTrainer._run_iteration (the graph replay has the same dependencies and ~0.25 ms less launch overhead)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
import numpy as np
import torch


def main():
    from rdg_b200 import weights as W
    from rdg_b200.engine import Context, Generator, Critic, GanTrainer
    ctx = Context(16, 1, device=0, max_chunk=1024)
    gen = Generator(W.init_generator_weights(0), ctx=ctx, mode="fp16")
    crit = Critic(W.init_critic_weights(1), ctx=ctx)
    B = 32
    rng = np.random.default_rng(7)
    lg = rng.standard_normal((5, B, 24, 16, 16, 1)).astype(np.float32) * 2
    ex = np.exp(lg - lg.max(axis=2, keepdims=True))
    x_real = ctx.dev((ex / ex.sum(axis=2, keepdims=True)).astype(np.float32))
    cond = ctx.dev((np.clip(rng.gamma(0.8, 12.0, size=(5, B, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32))
    tr = GanTrainer(gen, crit, gen_mode="fp16", seed=100, train_mode="tf32")
    dl = torch.zeros((5, 4), device="cuda"); gl = torch.zeros(1, device="cuda")
    marks = []

    def mark(name, stream=None):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream or torch.cuda.current_stream())
        marks.append((name, e))

    orig_run_step, orig_launch = tr._run_step, tr._launch_update
    state = {"k": 0}

    def run_step(which, p1, p2):
        k = state["k"]; state["k"] += 1
        name = f"critic{k}" if which == tr.WHICH_CRITIC else "gen"
        def p2m():
            mark(name + ".p2_begin")
            p2()
            mark(name + ".p2_end")
        orig_run_step(which, p1, p2m)

    def launch_update(which):
        orig_launch(which)
        mark("update_end", tr._upd_stream)

    tr._run_step, tr._launch_update = run_step, launch_update
    rows = []
    for it in range(13):
        marks.clear(); state["k"] = 0
        torch.cuda.synchronize()
        mark("start")
        tr.iteration_device(x_real, cond, cond[0], dl, gl)
        tr.finish()
        mark("end")
        torch.cuda.synchronize()
        if it >= 3:
            t0 = marks[0][1]
            rows.append([(n, t0.elapsed_time(e) * 1e3) for n, e in marks])
    names = [n for n, _ in rows[0]]
    med = [float(np.median([r[i][1] for r in rows])) for i in range(len(names))]
    print(json.dumps({"unit": "us since iteration start (median of 10 eager iterations)", "timeline": [[n, round(t, 1)] for n, t in zip(names, med)]}))


if __name__ == "__main__":
    main()
