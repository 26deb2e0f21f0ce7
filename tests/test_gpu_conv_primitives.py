"""FP32 SIMT Conv3D primitives (forward / backward-data / backward-filter) vs torch CPU float64,
for the generator (stride 1, fused nearest upsample) and critic (stride 2, TF 'same'/'valid') geometries."""
import ctypes as C

import zlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from rdg_b200 import _lib

pytestmark = pytest.mark.gpu


def _torch_conv(x, w, b, stride, pads, up):
    xt = torch.as_tensor(x, dtype=torch.float64).permute(0, 4, 1, 2, 3)
    if up:
        xt = xt.repeat_interleave(2, 2).repeat_interleave(2, 3).repeat_interleave(2, 4)
    xt = F.pad(xt, [pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]])
    wt = torch.as_tensor(w, dtype=torch.float64).permute(4, 3, 0, 1, 2)
    return F.conv3d(xt, wt, None if b is None else torch.as_tensor(b, dtype=torch.float64), stride=stride)


CASES = [  # name, B, (Ti,Hi,Wi), Ci, Co, stride, pads((before,after) per axis), up
    ("gen_conv3", 3, (12, 8, 8), 128, 64, 1, ((1, 1),) * 3, 1),
    ("gen_conv3_b9", 9, (12, 8, 8), 128, 64, 1, ((1, 1),) * 3, 1),
    ("gen_out", 5, (24, 16, 16), 64, 1, 1, ((1, 1),) * 3, 0),
    ("critic1_valid", 4, (24, 16, 16), 2, 64, 2, ((0, 0),) * 3, 0),
    ("critic3_same_asym", 7, (6, 4, 4), 128, 256, 2, ((0, 1),) * 3, 0),
    ("critic4_same_mixed", 6, (3, 2, 2), 256, 256, 2, ((1, 1), (0, 1), (0, 1)), 0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_primitives(ctx16, case):
    name, B, (Ti, Hi, Wi), Ci, Co, stride, pads, up = case
    rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)      # str hashes are salted per process: keep the inputs fixed
    x = rng.standard_normal((B, Ti, Hi, Wi, Ci)).astype(np.float32)
    w = (rng.standard_normal((3, 3, 3, Ci, Co)) * 0.05).astype(np.float32)
    b = rng.standard_normal(Co).astype(np.float32)
    xt = torch.as_tensor(x, dtype=torch.float64).requires_grad_(True)
    wt = torch.as_tensor(w, dtype=torch.float64).requires_grad_(True)
    bt = torch.as_tensor(b, dtype=torch.float64).requires_grad_(True)
    xin = xt.permute(0, 4, 1, 2, 3)
    if up:
        xin = xin.repeat_interleave(2, 2).repeat_interleave(2, 3).repeat_interleave(2, 4)
    xin.retain_grad()
    xp = F.pad(xin, [pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]])
    y = F.conv3d(xp, wt.permute(4, 3, 0, 1, 2), bt, stride=stride)
    dy = rng.standard_normal(tuple(y.shape)).astype(np.float32)          # (B,Co,To,Ho,Wo)
    y.backward(torch.as_tensor(dy, dtype=torch.float64))
    y_ref = y.detach().permute(0, 2, 3, 4, 1).numpy()
    dx_ref = xin.grad.permute(0, 2, 3, 4, 1).numpy()                      # w.r.t. the (upsampled) conv input
    dw_ref, db_ref = wt.grad.numpy(), bt.grad.numpy()
    To, Ho, Wo = y_ref.shape[1:4]
    geom = (C.c_int * 17)(B, Ti, Hi, Wi, Ci, To, Ho, Wo, Co, 3, 3, 3, stride, pads[0][0], pads[1][0], pads[2][0], up)
    dev = ctx16.dev
    lib = ctx16.lib
    xd, wd, bd = dev(x), dev(w), dev(b)
    dyd = dev(np.ascontiguousarray(np.transpose(dy, (0, 2, 3, 4, 1))))
    yd = torch.empty(y_ref.shape, device=xd.device)
    dxd = torch.empty(dx_ref.shape, device=xd.device)
    dwd = torch.zeros(dw_ref.shape, device=xd.device)
    dbd = torch.zeros(db_ref.shape, device=xd.device)
    P = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.rdg_conv3d(0, geom, P(xd), P(wd), P(bd), P(yd), None, 0, None))
    _lib.check(lib.rdg_conv3d(1, geom, P(dyd), P(wd), None, P(dxd), None, 0, None))
    _lib.check(lib.rdg_conv3d(2, geom, P(xd), P(dyd), None, P(dwd), P(dbd), 0, None))
    torch.cuda.synchronize()
    rel = lambda a, r: float(np.linalg.norm(a.astype(np.float64) - r) / np.linalg.norm(r))
    assert rel(yd.cpu().numpy(), y_ref) <= 5e-6
    assert rel(dxd.cpu().numpy(), dx_ref) <= 5e-6
    assert rel(dwd.cpu().numpy(), dw_ref) <= 5e-6
    assert rel(dbd.cpu().numpy(), db_ref) <= 5e-6
