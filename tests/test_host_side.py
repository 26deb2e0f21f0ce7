"""Host-side pieces that need no GPU: HDF5 subset, weight initialisers, C-ABI symbol table, sharding,
and the data-parallel gradient exchange on 2 gloo ranks."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

from rdg_b200 import _lib, hdf5, weights as W
from rdg_b200 import dist as rdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rdg_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rdg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = _lib.load()                      # loading makes no CUDA call
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/rdg_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes binding table out of sync with the header"
    assert lib.rdg_version() == 100


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rdg_b200.engine import Context
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Context(16, 1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import, include, link or execute it."""
    pkg = os.path.join(ROOT, "pr-disagg-radar-gan_b200")
    pat = re.compile(r"^\s*(import|from)\s+rdg_oracle|sys\.path.*oracle|#include\s*[\"<].*oracle|CDLL\(.*oracle|oracle/_ref", re.M)
    n = 0
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                n += 1
                assert not pat.search(open(os.path.join(dp, f), errors="ignore").read()), f
    assert n >= 12


def test_initialisers():
    gw = W.init_generator_weights(0)
    assert abs(float(np.std(gw[2])) - 0.02) < 2e-4 and all(np.all(b == 0) for b in gw[1::2])
    cw = W.init_critic_weights(1)
    lim = np.sqrt(6.0 / (27 * 2 + 27 * 64))
    assert cw[0].max() <= lim and cw[0].min() >= -lim and abs(float(cw[0].std()) - lim / np.sqrt(3)) < 0.05 * lim
    with pytest.raises(ValueError):
        W.check_shapes(gw[:-1], W.generator_shapes(16), "generator")
    with pytest.raises(ValueError):
        W.check_shapes(gw, W.generator_shapes(64), "generator")


def test_hdf5_keras_roundtrip(tmp_path):
    gw = W.randomize_biases(W.init_generator_weights(5))
    p = str(tmp_path / "gen.h5")
    hdf5.save_keras_weights(p, gw, "generator", model_config='{"class_name": "Model"}')
    f = hdf5.H5File(p)
    assert f.keys() == ["model_weights"]
    assert f.attrs["keras_version"] == b"2.2.4-tf" and f.attrs["backend"] == b"tensorflow"
    mw = f["model_weights"]
    assert [x.decode() for x in mw.attrs["layer_names"]] == ["input_1", "input_2", "flatten", "concatenate", "sequential"]
    assert mw["sequential"].attrs["weight_names"][0] == b"dense/kernel:0"
    assert mw["input_1"].attrs["weight_names"].shape == (0,)
    k = f["model_weights/sequential/conv3d_2/kernel:0"]
    assert k.shape == (3, 3, 3, 128, 64) and k.dtype == np.float32
    back = hdf5.load_keras_weights(p)
    assert len(back) == 10 and all(np.array_equal(a, b) for a, b in zip(back, gw))
    cw = W.init_critic_weights(6)
    hdf5.save_keras_weights(str(tmp_path / "disc.h5"), cw, "critic")
    assert all(np.array_equal(a, b) for a, b in zip(hdf5.load_keras_weights(str(tmp_path / "disc.h5")), cw))


def test_hdf5_generic_tree_and_errors(tmp_path):
    tree = {"attrs": {"n": np.int32(7), "names": np.array([b"a", b"bcd"])},
            "children": {"g": {"children": {"x": np.arange(6, dtype=np.float64).reshape(2, 3),
                                            "i": np.arange(4, dtype=np.int32), "empty": np.zeros((0,), np.float32)}},
                         "top": np.float32([1.5])}}
    data = hdf5.write_h5(None, tree)
    f = hdf5.H5File(data)
    assert f.attrs["n"] == 7 and list(f.attrs["names"]) == [b"a", b"bcd"]
    assert np.array_equal(f["g/x"], np.arange(6.0).reshape(2, 3)) and np.array_equal(f["g/i"], np.arange(4))
    assert f["g/empty"].shape == (0,) and f["top"][0] == 1.5
    with pytest.raises(KeyError):
        f["g/missing"]
    with pytest.raises(ValueError):
        hdf5.H5File(b"not an hdf5 file at all")
    bad = bytearray(data); bad[8] = 2
    with pytest.raises(NotImplementedError):
        hdf5.H5File(bytes(bad))
    with pytest.raises(ValueError):
        hdf5.load_keras_weights(_write(tmp_path, {"children": {"x": np.zeros(3, np.float32)}}))


def test_hdf5_reads_a_libhdf5_written_file_with_user_block():
    """External fixture, NOT produced by this repo's writer: scipy's MATLAB-v7.3 test file (written by libhdf5 through MATLAB 7.4
    on GLNX86; copied from scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat).  512-byte user block -> superblock at 512,
    base address 512, v0 superblock, symbol-table root group, v1 object headers, contiguous float64 dataset, string attribute."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "external_libhdf5_userblock.mat")
    raw = open(path, "rb").read()
    assert raw[:6] == b"MATLAB" and raw[512:520] == hdf5.SIG          # the signature is NOT at offset 0
    f = hdf5.H5File(path)
    assert f.keys() == ["testdouble"]
    x = f["testdouble"]
    assert x.dtype == np.float64 and x.shape == (9, 1)
    np.testing.assert_allclose(x[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
    leaf = f._r.children(f._r.root["ohdr"])["testdouble"]
    assert f._r.attributes(leaf)["MATLAB_class"] == b"double"
    # the same bytes with the user block cut off and the base address patched to 0 read identically
    cut = bytearray(raw[512:]); cut[24:32] = (0).to_bytes(8, "little")
    np.testing.assert_array_equal(hdf5.H5File(bytes(cut))["testdouble"], x)


def test_hdf5_own_file_behind_a_user_block(tmp_path):
    gw = W.init_critic_weights(3)
    p = str(tmp_path / "c.h5")
    hdf5.save_keras_weights(p, gw, "critic")
    raw = bytearray(open(p, "rb").read())
    ver = raw[8]
    pos = 24 if ver == 0 else 28
    raw[pos:pos + 8] = (1024).to_bytes(8, "little")                     # base address = user-block size
    blob = bytes(1024) + bytes(raw)
    f = hdf5.H5File(blob)
    names = [x.decode() for x in f["model_weights"].attrs["layer_names"]]
    seq = [n for n in names if n.startswith("sequential")][0]
    wn = [x.decode() for x in f["model_weights"][seq].attrs["weight_names"]]
    np.testing.assert_array_equal(f["model_weights/" + seq + "/" + wn[0]], gw[0])
    with pytest.raises(ValueError):
        bad = bytearray(blob); bad[1024 + pos:1024 + pos + 8] = (4096).to_bytes(8, "little")
        hdf5.H5File(bytes(bad))

def _write(tmp_path, tree):
    p = str(tmp_path / "t.h5")
    hdf5.write_h5(p, tree)
    return p


def test_shard_range_partitions():
    for total, world, align in [(1_000_000, 8, 100), (10, 4, 1), (7, 8, 1), (1000, 3, 100), (0, 2, 1)]:
        spans = [rdist.shard_range(total, r, world, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(lo % align == 0 for lo, _ in spans if lo < total)
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= align
    with pytest.raises(ValueError):
        rdist.shard_range(10, 2, 2)


def _dp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rdg_oracle as O
    torch.set_num_threads(2)
    gw = W.init_generator_weights(0); cw = W.randomize_biases(W.init_critic_weights(1))
    rng = np.random.default_rng(7)
    B = 2 * world
    x = rng.standard_normal((B, 24, 16, 16, 1)); x = np.exp(x - x.max(1, keepdims=True)); x = (x / x.sum(1, keepdims=True)).astype(np.float32)
    cond = (rng.gamma(0.8, 12.0, size=(B, 16, 16, 1)) / 127.4).astype(np.float32)
    z = rng.standard_normal((B, 100)).astype(np.float32)
    alpha = rng.random((B, 1, 1, 1, 1)).astype(np.float32)
    lo, hi = rdist.shard_range(B, rank, world)
    _, grads, _ = O.critic_step(gw, cw, x[lo:hi], cond[lo:hi], z[lo:hi], alpha[lo:hi], None, torch.float64)
    flat = torch.cat([torch.as_tensor(g).reshape(-1) for g in grads])
    rdist.allreduce_sum_(flat)
    flat /= world                                          # grad_scale = 1/world in rdg_adam_apply
    if rank == 0:
        _, full, _ = O.critic_step(gw, cw, x, cond, z, alpha, None, torch.float64)
        ref = torch.cat([torch.as_tensor(g).reshape(-1) for g in full])
        q.put(float((flat - ref).norm() / ref.norm()))
    dist.destroy_process_group()


def test_data_parallel_gradients_match_single_rank_gloo():
    """SURVEY 8d config #4 parity rule: N-rank averaged gradients == single-rank gradients on the concatenated
    batch (the WGAN-GP loss is a batch mean, incl. the gradient-penalty term)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    err = q.get(timeout=300)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert err <= 1e-12


def test_numa_binding_is_best_effort_without_a_gpu():
    """bind_to_gpu_numa_node must never raise (no GPU here, single-node VMs on the GPU boxes) and must not shrink the CPU set."""
    import os
    from rdg_b200.dist import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before
