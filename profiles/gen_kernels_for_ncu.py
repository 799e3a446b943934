"""General-noise tile kernels at an HBM-bound shape, once with per-thread loads and once TMA-staged (for ncu).

    ncu --set full --clock-control none --import-source on -k regex:"gen_" -c 8 -o gpurun_out/prof_gen_tma \
        python profiles/gen_kernels_for_ncu.py
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
for (B, D, M) in ((65536, 32, 64), (65536, 64, 16)):
    L = _cabi.make_launch(torch.float32, _cabi.NOISE_GENERAL, B, D, M)
    nz = _cabi.Noise()
    nz.source, nz.key, nz.cell_id, nz.n_cells, nz.h, nz.h_total = _cabi.SRC_COUNTER, key.data_ptr(), 42, 1, dt, dt
    y, f, o = (torch.rand(B, D, device=dev) for _ in range(3))
    g = torch.rand(B, D, M, device=dev)
    for mode in ('0', '2'):
        os.environ['TSDE_GEN_TMA'] = mode
        for _ in range(2):
            _cabi.check(lib.tsde_step_euler(ctypes.byref(L), ctypes.byref(nz), y.data_ptr(), f.data_ptr(), g.data_ptr(),
                                            dt, o.data_ptr()), 'tsde_step_euler')
    torch.cuda.synchronize()
