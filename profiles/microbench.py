"""Micro-benchmarks of the cfg2 step's kernels in isolation (CUDA events, rotating buffers > L2).
Used for tuning; numbers quoted in profiles/*.md come from ncu, not from here."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torchsde_b200 import _cabi  # noqa: E402

dev = torch.device('cuda')
B, D = 65536, 64
NSET = 12  # 12 x 5 x 16 MiB = 960 MiB >> 126 MB L2
lib = _cabi.lib()
key = torch.tensor([123456789], dtype=torch.int64, device=dev)
sets = [[torch.rand(B, D, device=dev) for _ in range(6)] for _ in range(NSET)]
L = _cabi.make_launch(torch.float32, _cabi.NOISE_DIAGONAL, B, D, D)
nz = _cabi.Noise()
nz.source = _cabi.SRC_COUNTER
nz.key = key.data_ptr()
nz.cell_id = 42
nz.n_cells = 1
nz.h = 2.0 ** -10
nz.h_total = 2.0 ** -10
mu = torch.rand(D, device=dev)
MU = mu.expand(B, D).contiguous()
DIAG = torch.diag(mu)


def timeit(name, fn, nbytes, reps=5):
    for i in range(NSET):
        fn(sets[i])
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(NSET):
            fn(sets[i])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / NSET * 1e3)
    print(f"{name:34s} {best:7.2f} us   {nbytes / best / 1e3:7.1f} GB/s")


MB = B * D * 4
dt = 2.0 ** -10
timeit('tsde milstein tableau (5 Ds)', lambda s: lib.tsde_step_milstein(ctypes.byref(L), ctypes.byref(nz), s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), s[3].data_ptr(), dt, s[4].data_ptr()), 5 * MB)
timeit('tsde milstein vjp seed (2 Ds)', lambda s: lib.tsde_milstein_vjp_seed(ctypes.byref(L), ctypes.byref(nz), s[0].data_ptr(), dt, 1, s[1].data_ptr()), 2 * MB)
timeit('tsde euler tableau (4 Ds)', lambda s: lib.tsde_step_euler(ctypes.byref(L), ctypes.byref(nz), s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), dt, s[4].data_ptr()), 4 * MB)
timeit('tsde brownian cells W (1 Ds)', lambda s: lib.tsde_brownian_cells(ctypes.byref(L), ctypes.byref(nz), s[0].data_ptr(), None, None), 1 * MB)
timeit('torch mu * y  (broadcast, 2 Ds)', lambda s: torch.mul(s[0], mu, out=s[1]), 2 * MB)
timeit('torch MU * y  (contiguous, 3 Ds)', lambda s: torch.mul(s[0], MU, out=s[1]), 3 * MB)
timeit('torch y @ diag(mu) (2 Ds)', lambda s: torch.mm(s[0], DIAG, out=s[1]), 2 * MB)
timeit('torch y * 0.5 (2 Ds)', lambda s: torch.mul(s[0], 0.5, out=s[1]), 2 * MB)
timeit('torch copy (2 Ds)', lambda s: s[1].copy_(s[0]), 2 * MB)
