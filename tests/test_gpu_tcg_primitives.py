"""Tensor-core training primitives (tcgen05 kind::tf32, csrc/tcg_gemm.cu: forward / backward-data / backward-filter, plain and
upsample-folded) vs torch CPU float64, on the geometries the cWGAN-GP step uses: critic stride-2 TF-'same' convs with their
asymmetric pads, the generator's UpSampling3D + Conv3D blocks (folded: backward lands on the LOW-RES grid) and Dense.
Bound: tf32 operands (10-bit mantissa, round-to-nearest) with FP32 accumulation -> 2e-3 relative L2 per tensor."""
import ctypes as C
import zlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from rdg_b200 import _lib

pytestmark = pytest.mark.gpu

TOL = 2e-3

CASES = [  # name, B, (Ti,Hi,Wi), Ci, Co, k, stride, pads((before,after) per axis), up
    ("critic2_same", 5, (11, 7, 7), 64, 128, 3, 2, ((1, 1),) * 3, 0),
    ("critic2_same_b96", 96, (11, 7, 7), 64, 128, 3, 2, ((1, 1),) * 3, 0),
    ("critic3_same_asym", 7, (6, 4, 4), 128, 256, 3, 2, ((0, 1),) * 3, 0),
    ("critic3_same_asym_b64", 64, (6, 4, 4), 128, 256, 3, 2, ((0, 1),) * 3, 0),
    ("critic4_same_mixed", 6, (3, 2, 2), 256, 256, 3, 2, ((1, 1), (0, 1), (0, 1)), 0),
    ("critic4_same_mixed_b96", 96, (3, 2, 2), 256, 256, 3, 2, ((1, 1), (0, 1), (0, 1)), 0),
    ("critic2_nd64", 2, (11, 31, 31), 64, 128, 3, 2, ((1, 1),) * 3, 0),
    ("stride1_same", 3, (5, 6, 7), 32, 64, 3, 1, ((1, 1),) * 3, 0),
    ("gen_up1", 4, (3, 2, 2), 256, 256, 3, 1, ((1, 1),) * 3, 1),
    ("gen_up2", 3, (6, 4, 4), 256, 128, 3, 1, ((1, 1),) * 3, 1),
    ("gen_up3", 3, (12, 8, 8), 128, 64, 3, 1, ((1, 1),) * 3, 1),
    ("gen_up3_b32", 32, (12, 8, 8), 128, 64, 3, 1, ((1, 1),) * 3, 1),
    ("gen_dense", 32, (1, 1, 1), 356, 3072, 1, 1, ((0, 0),) * 3, 0),
    ("critic1_valid", 4, (24, 16, 16), 2, 64, 3, 2, ((0, 0),) * 3, 0),
    ("critic1_valid_b96", 96, (24, 16, 16), 2, 64, 3, 2, ((0, 0),) * 3, 0),
    ("critic1_valid_ncond3", 3, (24, 16, 16), 4, 64, 3, 2, ((0, 0),) * 3, 0),
    ("critic1_valid_ncond2", 3, (24, 16, 16), 3, 64, 3, 2, ((0, 0),) * 3, 0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tcg_primitives(ctx16, case):
    name, B, (Ti, Hi, Wi), Ci, Co, k, stride, pads, up = case
    rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)
    x = rng.standard_normal((B, Ti, Hi, Wi, Ci)).astype(np.float32)
    w = (rng.standard_normal((k, k, k, Ci, Co)) * 0.05).astype(np.float32)
    b = rng.standard_normal(Co).astype(np.float32)
    xt = torch.as_tensor(x, dtype=torch.float64).requires_grad_(True)
    wt = torch.as_tensor(w, dtype=torch.float64).requires_grad_(True)
    bt = torch.as_tensor(b, dtype=torch.float64).requires_grad_(True)
    xin = xt.permute(0, 4, 1, 2, 3)
    if up:
        xin = xin.repeat_interleave(2, 2).repeat_interleave(2, 3).repeat_interleave(2, 4)
    xp = F.pad(xin, [pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]])
    y = F.conv3d(xp, wt.permute(4, 3, 0, 1, 2), bt, stride=stride)
    dy = rng.standard_normal(tuple(y.shape)).astype(np.float32)          # (B,Co,To,Ho,Wo)
    y.backward(torch.as_tensor(dy, dtype=torch.float64))
    y_ref = y.detach().permute(0, 2, 3, 4, 1).numpy()
    dx_ref = xt.grad.numpy()                      # w.r.t. the conv's (low-res, if upsampled) input
    dw_ref = wt.grad.numpy()
    To, Ho, Wo = y_ref.shape[1:4]
    geom = (C.c_int * 17)(B, Ti, Hi, Wi, Ci, To, Ho, Wo, Co, k, k, k, stride, pads[0][0], pads[1][0], pads[2][0], up)
    dev, lib = ctx16.dev, ctx16.lib
    xd, wd, bd = dev(x), dev(w), dev(b)
    dyd = dev(np.ascontiguousarray(np.transpose(dy, (0, 2, 3, 4, 1))))
    yd = torch.full(y_ref.shape, float("nan"), device=xd.device)
    dxd = torch.full(dx_ref.shape, float("nan"), device=xd.device)
    dwd = torch.zeros(dw_ref.shape, device=xd.device)
    P = lambda t: C.c_void_p(t.data_ptr())
    rel = lambda a, r: float(np.linalg.norm(a.astype(np.float64) - r) / np.linalg.norm(r))
    _lib.check(lib.rdg_conv3d(10, geom, P(xd), P(wd), P(bd), P(yd), None, 0, None))
    torch.cuda.synchronize()
    assert rel(yd.cpu().numpy(), y_ref) <= TOL
    yd.fill_(float("nan"))
    _lib.check(lib.rdg_conv3d(13, geom, P(xd), P(wd), P(bd), P(yd), None, 0, None))      # 3xTF32: FP32-grade forward
    torch.cuda.synchronize()
    assert rel(yd.cpu().numpy(), y_ref) <= 2e-5
    if Ci % 32 == 0:
        _lib.check(lib.rdg_conv3d(11, geom, P(dyd), P(wd), None, P(dxd), None, 0, None))
        torch.cuda.synchronize()
        assert rel(dxd.cpu().numpy(), dx_ref) <= TOL
    if k == 3:
        _lib.check(lib.rdg_conv3d(12, geom, P(xd), P(dyd), None, P(dwd), None, 0, None))
        torch.cuda.synchronize()
        assert rel(dwd.cpu().numpy(), dw_ref) <= TOL


def test_tcg_forward_epilogue(ctx16):
    """bias + LeakyReLU through the fused epilogue and through the split-K reduction give the same function."""
    rng = np.random.default_rng(5)
    for B in (2, 96):     # 2: one tile, split-K; 96: many tiles
        x = rng.standard_normal((B, 6, 4, 4, 128)).astype(np.float32)
        w = (rng.standard_normal((3, 3, 3, 128, 256)) * 0.05).astype(np.float32)
        b = rng.standard_normal(256).astype(np.float32)
        xp = F.pad(torch.as_tensor(x, dtype=torch.float64).permute(0, 4, 1, 2, 3), [0, 1, 0, 1, 0, 1])
        y = F.conv3d(xp, torch.as_tensor(w, dtype=torch.float64).permute(4, 3, 0, 1, 2), torch.as_tensor(b, dtype=torch.float64), stride=2)
        y_ref = F.leaky_relu(y, 0.2).permute(0, 2, 3, 4, 1).numpy()
        geom = (C.c_int * 17)(B, 6, 4, 4, 128, 3, 2, 2, 256, 3, 3, 3, 2, 0, 0, 0, 0)
        dev, lib = ctx16.dev, ctx16.lib
        xd, wd, bd = dev(x), dev(w), dev(b)
        yd = torch.full(y_ref.shape, float("nan"), device=xd.device)
        P = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(lib.rdg_conv3d(10, geom, P(xd), P(wd), P(bd), P(yd), None, 1, None))
        torch.cuda.synchronize()
        assert float(np.linalg.norm(yd.cpu().numpy() - y_ref) / np.linalg.norm(y_ref)) <= TOL
