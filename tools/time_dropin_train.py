#!/usr/bin/env python
"""Wall time per iteration of the drop-in train() (gan_train_cwgangp_pixelnorm.py: 5 critic steps + 1 generator step, batch 32,
synthetic radar-shaped data), through its three forms: the reference's six train_on_batch calls in the FP32 mode, the same calls in
the tensor-core mode, and the graph replay (default in the tensor-core mode) with the host / the device batch sampler."""
import contextlib, importlib, io, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))


def run(mode, graph, device_sampler, iters=100):
    os.environ["RDG_TRAIN_MODE"] = mode
    os.environ["RDG_TRAIN_GRAPH"] = "1" if graph else "0"
    sys.modules.pop("gan_train_cwgangp_pixelnorm", None)
    m = importlib.import_module("gan_train_cwgangp_pixelnorm")
    m.hist = {'d_loss': [], 'g_loss': []}
    m.setup(seed=3, device_sampler=device_sampler)
    import torch
    with contextlib.redirect_stdout(io.StringIO()):
        m.train(1, 32, bat_per_epo=5, save=False)          # warm-up (allocations, capture)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.train(1, 32, bat_per_epo=iters, save=False)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


if __name__ == "__main__":
    os.chdir("/tmp")
    res = {"unit": "ms per iteration (wall, batch 32, includes batch sampling, the per-iteration loss read and print)"}
    res["fp32_train_on_batch"] = run("fp32", False, False, 10)
    res["tf32_train_on_batch"] = run("tf32", False, False, 20)
    res["tf32_graph_host_sampler"] = run("tf32", True, False)
    res["tf32_graph_device_sampler"] = run("tf32", True, True)
    print(json.dumps(res))
