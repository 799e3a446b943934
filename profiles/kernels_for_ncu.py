"""Launch each solver-owned kernel a few times at its benchmark shape (for `ncu --set full`).

    ncu --set full --clock-control none --import-source on -k regex:"tsde" -c 24 -o gpurun_out/prof_kernels \
        python profiles/kernels_for_ncu.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torchsde_b200 import _cabi  # noqa: E402

dev = torch.device('cuda')
lib = _cabi.lib()
key = torch.tensor([123456789], dtype=torch.int64, device=dev)
dt = 2.0 ** -10


def noise(want_u=False):
    nz = _cabi.Noise()
    nz.source, nz.want_u, nz.key, nz.cell_id, nz.n_cells, nz.h, nz.h_total = _cabi.SRC_COUNTER, int(want_u), key.data_ptr(), 42, 1, dt, dt
    return nz


# cfg2 shapes: diagonal, B=65536, D=64
B, D = 65536, 64
s = [torch.rand(B, D, device=dev) for _ in range(9)]
L = _cabi.make_launch(torch.float32, _cabi.NOISE_DIAGONAL, B, D, D)
nz, nzu = noise(), noise(True)
for _ in range(3):
    lib.tsde_milstein_vjp_seed(ctypes.byref(L), ctypes.byref(nz), s[0].data_ptr(), dt, 1, s[1].data_ptr())
    lib.tsde_step_milstein(ctypes.byref(L), ctypes.byref(nz), s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), s[3].data_ptr(), dt, s[4].data_ptr())
    lib.tsde_step_euler(ctypes.byref(L), ctypes.byref(nz), s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), dt, s[4].data_ptr())
    lib.tsde_step_srk_diag(ctypes.byref(L), ctypes.byref(nzu), *(x.data_ptr() for x in s[:8]), dt, 1 / dt, dt ** .5, 3 * dt, s[8].data_ptr())
    lib.tsde_brownian_cells(ctypes.byref(L), ctypes.byref(nz), s[0].data_ptr(), None, None)
    lib.tsde_brownian_cells(ctypes.byref(L), ctypes.byref(nzu), s[0].data_ptr(), s[1].data_ptr(), None)

# cfg3 shapes: general, B=8192, D=32, M=16
B, D, M = 8192, 32, 16
g = [torch.rand(B, D, M, device=dev) for _ in range(2)]
e = [torch.rand(B, D, device=dev) for _ in range(4)]
Lg = _cabi.make_launch(torch.float32, _cabi.NOISE_GENERAL, B, D, M)
for _ in range(3):
    lib.tsde_step_euler(ctypes.byref(Lg), ctypes.byref(nz), e[0].data_ptr(), e[1].data_ptr(), g[0].data_ptr(), dt, e[3].data_ptr())
    lib.tsde_step_heun(ctypes.byref(Lg), ctypes.byref(nz), e[0].data_ptr(), e[1].data_ptr(), e[2].data_ptr(), g[0].data_ptr(), g[1].data_ptr(), dt, e[3].data_ptr())
# a larger general problem (HBM-bound regime): B=65536, D=64, M=16 -> g = 256 MiB
B, D, M = 65536, 64, 16
g2 = torch.rand(B, D, M, device=dev)
e2 = [torch.rand(B, D, device=dev) for _ in range(3)]
Lg2 = _cabi.make_launch(torch.float32, _cabi.NOISE_GENERAL, B, D, M)
for _ in range(3):
    lib.tsde_step_euler(ctypes.byref(Lg2), ctypes.byref(nz), e2[0].data_ptr(), e2[1].data_ptr(), g2.data_ptr(), dt, e2[2].data_ptr())
# Levy area (cfg5 foster): B=131072, M=16
B, M = 131072, 16
w, h = torch.randn(B, M, device=dev), torch.randn(B, M, device=dev)
a = torch.empty(B, M, M, device=dev)
Lb = _cabi.make_launch(torch.float32, _cabi.NOISE_DIAGONAL, B, M, M)
for _ in range(3):
    lib.tsde_brownian_levy_area(ctypes.byref(Lb), key.data_ptr(), 0, 77, w.data_ptr(), h.data_ptr(), 2.0 ** -6, 1, a.data_ptr())
# ... and the fused whole-cell query (W, U and A drawn in one launch), same shape
wo, uo = torch.empty(B, M, device=dev), torch.empty(B, M, device=dev)
nzc = noise()
nzc.h = nzc.h_total = 2.0 ** -6
for _ in range(3):
    rc = lib.tsde_brownian_cell_levy(ctypes.byref(Lb), ctypes.byref(nzc), 77, 1, wo.data_ptr(), uo.data_ptr(), a.data_ptr())
    assert rc == 0, rc
# bmm(g, A) of the log-ODE correction at the cfg3 size
B, D, M = 8192, 32, 16
gg = torch.randn(B, D, M, device=dev)
aa = torch.randn(B, M, M, device=dev)
oo = torch.empty(M, B, D, device=dev)
Lbm = _cabi.make_launch(torch.float32, _cabi.NOISE_GENERAL, B, D, M)
for _ in range(3):
    lib.tsde_bmm_ga(ctypes.byref(Lbm), gg.data_ptr(), aa.data_ptr(), oo.data_ptr())
torch.cuda.synchronize()
print('ok')
