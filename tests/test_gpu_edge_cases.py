"""Edge cases of the CUDA path that round 1 only dry-ran on the CPU: empty batches, a single
state / Brownian channel, and launches of 2^31 quads and more (the 64-bit index path of the generic row-wise kernel)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import brownian as obm
from oracle import solvers
from . import helpers, problems

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _tsde():
    import torchsde_b200
    return torchsde_b200


@pytest.mark.parametrize('kind,sde_type,method,levy', [
    ('gbm', 'ito', 'euler', 'none'), ('gbm', 'ito', 'milstein', 'none'), ('gbm', 'ito', 'srk', 'space-time'),
    ('general', 'stratonovich', 'heun', 'none'), ('additive', 'ito', 'srk', 'space-time'),
    ('scalar', 'stratonovich', 'midpoint', 'none'), ('gbm', 'stratonovich', 'reversible_heun', 'none')])
@pytest.mark.parametrize('graph', [False, True])
def test_empty_batch(kind, sde_type, method, levy, graph):
    """Zero trajectories: every launcher returns early on an empty launch; the solve yields a (T, 0, d) series."""
    tsde = _tsde()
    d, m = 3, {'gbm': 3, 'scalar': 1}.get(kind, 2)
    sde = problems.make(kind, d, m, sde_type, dtype=torch.float32).to(DEV)
    ts = torch.tensor([0.0, 0.125, 0.25], device=DEV)
    bm = tsde.BrownianInterval(0.0, 0.25, size=(0, m), dtype=torch.float32, device=DEV, levy_area_approximation=levy)
    with torch.no_grad():
        ys = tsde.sdeint(sde, torch.ones(0, d, device=DEV), ts, bm=bm, method=method, dt=2.0 ** -4,
                         options={'cuda_graph': graph})
    assert ys.shape == (3, 0, d)
    assert bm(0.0, 0.125).shape == (0, m)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
@pytest.mark.parametrize('kind,sde_type,method,d,m', [
    ('gbm', 'ito', 'milstein', 1, 1), ('gbm', 'ito', 'srk', 1, 1), ('gbm', 'stratonovich', 'heun', 1, 1),
    ('additive', 'ito', 'srk', 5, 1), ('additive', 'ito', 'euler', 1, 1), ('general', 'stratonovich', 'midpoint', 1, 1),
    ('general', 'stratonovich', 'reversible_heun', 7, 1), ('general', 'ito', 'euler', 1, 3), ('scalar', 'ito', 'milstein', 1, 1)])
def test_single_channel_shapes_vs_oracle(dtype, kind, sde_type, method, d, m):
    """d = 1 and / or m = 1 (no 128-bit vector path, quads with one valid lane, GEMV over a single column) against the
    oracle on the same counter-based path."""
    tsde = _tsde()
    npdt = np.float32 if dtype == torch.float32 else np.float64
    B = 67
    sde = problems.make(kind, d, m, sde_type, dtype=dtype, seed=3).to(DEV)
    sde_cpu = problems.make(kind, d, m, sde_type, dtype=dtype, seed=3)
    bm_m = d if kind == 'gbm' else m
    levy = 'space-time' if method == 'srk' else 'none'
    y0 = (0.2 + 0.3 * torch.rand(B, d, generator=torch.Generator().manual_seed(1), dtype=torch.float64)).to(dtype)
    ts = np.array([0.0, 0.125, 0.25], dtype=npdt)
    bm = tsde.BrownianInterval(0.0, 0.25, size=(B, bm_m), dtype=dtype, device=DEV, entropy=55,
                               levy_area_approximation=levy)
    with torch.no_grad():
        ys = tsde.sdeint(sde, y0.to(DEV), torch.from_numpy(ts).to(DEV), bm=bm, method=method, dt=2.0 ** -4)
    ref, _ = solvers.make(method, problems.NumpySDE(sde_cpu),
                          helpers.oracle_grid_bm(bm, np.arange(B), bm_m, npdt, levy != 'none'), 2.0 ** -4).integrate(
        y0.numpy(), ts)
    tol = dict(rtol=1e-11, atol=1e-12) if dtype == torch.float64 else dict(rtol=5e-5, atol=1e-5)
    np.testing.assert_allclose(ys.cpu().numpy(), ref, **tol)


def test_more_than_2_31_quads():
    """rows x d/4 = 2^31 quads: beyond the 32-bit fast path, the generic kernel indexes with 64 bits
    (csrc/ew.cuh ew_kernel).  y' = y0 + g.dW with y0 = 0, g = 1 makes the output the increment itself, which is
    compared — first, middle and last rows — with the increments materialised for those rows alone."""
    from torchsde_b200 import _cabi
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    rows, d = 1 << 23, 1024                       # 2^33 elements = 32 GiB per fp32 tensor, 2^31 quads
    if free < 3.2 * rows * d * 4:
        pytest.skip('needs ~105 GB of free device memory')
    lib = _cabi.lib()
    key = torch.tensor([20260923], dtype=torch.int64, device=DEV)
    nz = _cabi.Noise()
    nz.source, nz.key, nz.cell_id, nz.n_cells, nz.h, nz.h_total = _cabi.SRC_COUNTER, key.data_ptr(), 11, 1, 0.25, 0.25
    y0 = torch.zeros(rows, d, device=DEV)
    g = torch.ones(rows, d, device=DEV)
    out = torch.empty(rows, d, device=DEV)
    L = _cabi.make_launch(torch.float32, _cabi.NOISE_DIAGONAL, rows, d, d)
    _cabi.check(lib.tsde_euler_heun_predict(ctypes.byref(L), ctypes.byref(nz), y0.data_ptr(), g.data_ptr(),
                                            out.data_ptr()), 'tsde_euler_heun_predict')
    torch.cuda.synchronize()
    for r in (0, 1, rows // 2 - 1, rows // 2, rows - 2, rows - 1):
        w = torch.empty(1, d, device=DEV)
        L1 = _cabi.make_launch(torch.float32, _cabi.NOISE_DIAGONAL, 1, d, d)
        nz.row_offset = r
        _cabi.check(lib.tsde_brownian_cells(ctypes.byref(L1), ctypes.byref(nz), w.data_ptr(), None, None),
                    'tsde_brownian_cells')
        assert torch.equal(out[r], w[0]), r
    # ... and against the oracle's definition of the last row (the key tensor IS the 64-bit Philox key)
    W, _ = obm.cell(int(key.item()), 11, 0.25, 1, d, np.float32, False, row_ids=np.array([rows - 1]))
    np.testing.assert_allclose(out[rows - 1].cpu().numpy(), W[0], rtol=2e-5, atol=5e-6)
    del y0, g, out
    torch.cuda.empty_cache()


def _sync_warnings(fn):
    """Run fn with PyTorch's synchronisation debug mode on; return the warnings about synchronising CUDA calls."""
    import warnings
    torch.cuda.synchronize()
    prev = torch.cuda.get_sync_debug_mode()
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter('always')
        torch.cuda.set_sync_debug_mode('warn')
        try:
            out = fn()
        finally:
            torch.cuda.set_sync_debug_mode(prev)
    torch.cuda.synchronize()
    # (the mode announces itself once per process with a 'prototype feature' warning: not a finding)
    return out, [f"{w.filename}:{w.lineno}: {w.message}" for w in caught
                 if 'synchroniz' in str(w.message).lower() and 'prototype feature' not in str(w.message)]


@pytest.mark.parametrize('graph', [True, False])
def test_cached_solve_does_not_synchronise_the_host(graph):
    """A training or sampling loop builds a new BrownianInterval per solve and calls sdeint again: with the schedule
    and the graph plan cached, nothing on that path may wait for the device (the host prepares solve k+1 while solve k
    replays).  Until late in r02 the interval's Philox key was uploaded with a synchronous host-to-device copy.
    The eager loop (no CUDA graph) uploads its per-step time table once per solve — one blocking copy, at
    `BaseSDESolver._contexts`, and nothing else."""
    tsde = _tsde()
    sde = problems.GBMDiagonal(8, 'ito', seed=1, dtype=torch.float32).to(DEV)
    ts = (torch.arange(5, dtype=torch.float32) * 2.0 ** -4).to(DEV)
    y0 = torch.full((64, 8), 0.1, device=DEV)

    def solve(entropy, graph):
        bm = tsde.BrownianInterval(0.0, 0.25, size=(64, 8), dtype=torch.float32, device=DEV, entropy=entropy)
        with torch.no_grad():
            return tsde.sdeint(sde, y0, ts, bm=bm, method='milstein', dt=2.0 ** -4, options={'cuda_graph': graph})

    first = solve(1, graph)
    solve(2, graph)
    again, syncs = _sync_warnings(lambda: solve(1, graph))
    if graph:
        assert syncs == [], syncs
    else:
        assert len(syncs) <= 1 and all('base_solver.py' in w for w in syncs), syncs
    assert torch.equal(first, again)


@pytest.mark.parametrize('levy', ['none', 'space-time', 'foster'])
def test_whole_cell_queries_do_not_synchronise_the_host(levy):
    tsde = _tsde()
    h = 2.0 ** -6

    def sweep():
        bm = tsde.BrownianInterval(0.0, 1.0, size=(256, 16), dtype=torch.float32, device=DEV, entropy=9, dt=h,
                                   levy_area_approximation=levy)
        return [bm(k * h, (k + 1) * h, return_U=levy != 'none', return_A=levy == 'foster') for k in (0, 7, 3)]

    sweep()
    _, syncs = _sync_warnings(sweep)
    assert syncs == [], syncs
