"""CUDA-event probe of the solver-owned element-wise kernels at the cfg2 tensor size (65536 x 64 fp32).

    python profiles/kernel_probe.py            [TORCHSDE_B200_LIB=<other build> for an A/B on the same box]

Each kernel is timed the way the solver issues it — launches captured into a CUDA graph and replayed — in two
regimes: 'cold' = every launch works on a different buffer set (12 sets, larger than the 126 MB L2: HBM traffic),
'warm' = the same buffer set every launch (operands resident in L2, as inside a solver step where the producer
kernel has just written them).  Prints one line per kernel: microseconds per launch, algorithmic GB/s and the
fraction of the measured HBM peak.
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from torchsde_b200 import _cabi  # noqa: E402

dev = torch.device('cuda')
lib = _cabi.lib()
PEAK = 6569.6
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                             'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
B, D = int(os.environ.get('PROBE_B', 65536)), int(os.environ.get('PROBE_D', 64))
dt = 2.0 ** -10
key = torch.tensor([987654321], dtype=torch.int64, device=dev)
L = _cabi.make_launch(torch.float32, _cabi.NOISE_DIAGONAL, B, D, D)


def noise(want_u=False):
    nz = _cabi.Noise()
    nz.source, nz.want_u, nz.key, nz.cell_id, nz.n_cells, nz.h, nz.h_total = \
        _cabi.SRC_COUNTER, int(want_u), key.data_ptr(), 7, 1, dt, dt
    return nz


nz, nzu = noise(), noise(True)
P = lambda t: t.data_ptr()  # noqa
KERNELS = {
    # name: (n_tensors, launch(s))
    'milstein_vjp_seed': (2, lambda s: lib.tsde_milstein_vjp_seed(ctypes.byref(L), ctypes.byref(nz), P(s[0]), dt, 1, P(s[1]))),
    'step_milstein': (5, lambda s: lib.tsde_step_milstein(ctypes.byref(L), ctypes.byref(nz), P(s[0]), P(s[1]), P(s[2]), P(s[3]), dt, P(s[4]))),
    'step_euler': (4, lambda s: lib.tsde_step_euler(ctypes.byref(L), ctypes.byref(nz), P(s[0]), P(s[1]), P(s[2]), dt, P(s[3]))),
    'brownian_cells_W': (1, lambda s: lib.tsde_brownian_cells(ctypes.byref(L), ctypes.byref(nz), P(s[0]), None, None)),
    'brownian_cells_WU': (2, lambda s: lib.tsde_brownian_cells(ctypes.byref(L), ctypes.byref(nzu), P(s[0]), P(s[1]), None)),
    'euler_heun_predict': (3, lambda s: lib.tsde_euler_heun_predict(ctypes.byref(L), ctypes.byref(nz), P(s[0]), P(s[1]), P(s[2]))),
    'step_srk_diag': (9, lambda s: lib.tsde_step_srk_diag(ctypes.byref(L), ctypes.byref(nzu), *(P(x) for x in s[:8]), dt, 1 / dt, dt ** .5, 3 * dt, P(s[8]))),
}


def probe(name, nt, launch, cold):
    nset = 12 if cold else 1
    reps = 12
    sets = [[torch.rand(B, D, device=dev) for _ in range(nt)] for _ in range(nset)]
    for s in sets:
        assert launch(s) == 0
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        L.stream = torch.cuda.current_stream(dev).cuda_stream
        for i in range(reps):
            launch(sets[i % nset])
    L.stream = torch.cuda.current_stream(dev).cuda_stream
    graph.replay()
    torch.cuda.synchronize()
    times = []
    for _ in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3 / reps)
    us = float(np.median(times))
    nbytes = nt * B * D * 4
    return us, nbytes / us / 1e3


print(f"lib = {_cabi.LIB_PATH}  B={B} D={D}")
for name, (nt, launch) in KERNELS.items():
    for cold in (True, False):
        us, gbs = probe(name, nt, launch, cold)
        print(f"{name:22s} {'cold' if cold else 'warm'}: {us:7.2f} us  {gbs:7.1f} GB/s  {gbs / PEAK * 100:5.1f} % of {PEAK}")
