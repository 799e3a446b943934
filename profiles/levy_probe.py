"""CUDA-event timing of the Levy-area kernels at the cfg5 shape (131072 x 16 x 16, fp32; each launch writes 134 MB,
more than L2): tsde_brownian_levy_area (W, H read) and tsde_brownian_cell_levy (W, U, A drawn in one launch).
    TORCHSDE_B200_LIB=profiles/_ab/libtsde_levy_4_3.so python profiles/levy_probe.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torchsde_b200 import _cabi  # noqa: E402

dev = torch.device('cuda')
lib = _cabi.lib()
key = torch.tensor([123456789], dtype=torch.int64, device=dev)
B, M = int(os.environ.get('LEVY_B', 131072)), int(os.environ.get('LEVY_M', 16))
w, h = torch.randn(B, M, device=dev), torch.randn(B, M, device=dev)
a = [torch.empty(B, M, M, device=dev) for _ in range(2)]
wo, uo = torch.empty(B, M, device=dev), torch.empty(B, M, device=dev)
L = _cabi.make_launch(torch.float32, _cabi.NOISE_DIAGONAL, B, M, M)
nz = _cabi.Noise()
nz.source, nz.want_u, nz.key, nz.cell_id, nz.n_cells, nz.h, nz.h_total = _cabi.SRC_COUNTER, 1, key.data_ptr(), 42, 1, 2.0 ** -6, 2.0 ** -6


def area(i, foster=1):
    return lib.tsde_brownian_levy_area(ctypes.byref(L), key.data_ptr(), 0, 77, w.data_ptr(), h.data_ptr(), 2.0 ** -6, foster, a[i & 1].data_ptr())


def cell(i, foster=1):
    return lib.tsde_brownian_cell_levy(ctypes.byref(L), ctypes.byref(nz), 77, foster, wo.data_ptr(), uo.data_ptr(), a[i & 1].data_ptr())


def timed(fn, n=50):
    for i in range(5):
        assert fn(i) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


out = {'lib': os.environ.get('TORCHSDE_B200_LIB', 'in-tree'), 'rows': B, 'm': M}
for name, fn in (('levy_area_foster', area), ('levy_area_davie', lambda i: area(i, 0)), ('cell_levy_foster', cell),
                 ('cell_levy_davie', lambda i: cell(i, 0))):
    out[name + '_us'] = round(timed(fn), 2)
bytes_area, bytes_cell = B * (M * M + 2 * M) * 4, B * (M * M + 2 * M) * 4
out['levy_area_foster_TBps'] = round(bytes_area / out['levy_area_foster_us'] / 1e6, 3)
out['cell_levy_foster_TBps'] = round(bytes_cell / out['cell_levy_foster_us'] / 1e6, 3)
print(out)
