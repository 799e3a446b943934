"""The TMA-staged general-noise path (`gen_tma_kernel`, TSDE_GEN_TMA) against the default tile kernel.

Both kernels use the same chunk -> lane mapping, the same fused multiply-add chain inside a 4-chunk and the
same xor-tree across chunks, so for identical increments their outputs must be BIT-IDENTICAL; the default
kernel itself is pinned against the oracle / the reference's golden files in test_gpu_solver.py.
"""
import pytest
import torch

from . import problems

pytestmark = pytest.mark.gpu


def _solve(method, sde_type, kind, B, d, m, dtype, levy, materialise, graph=False):
    import torchsde_b200 as tsde
    dev = torch.device('cuda')
    sde = problems.make(kind, d, m, sde_type, dtype=dtype, seed=11).to(dev)
    y0 = torch.full((B, d), 0.25, dtype=dtype, device=dev)
    ts = torch.tensor([0.0, 0.125, 0.25], dtype=dtype, device=dev)
    bm = tsde.BrownianInterval(0.0, 0.25, size=(B, m), dtype=dtype, device=dev, entropy=99,
                               levy_area_approximation=levy)
    if materialise:  # increments handed over as tensors (TSDE_SRC_MEMORY) instead of regenerated from the counter
        inner = bm

        class Materialised:
            shape, levy_area_approximation = inner.shape, inner.levy_area_approximation

            def __call__(self, ta, tb=None, return_U=False, return_A=False):
                return inner(ta, tb, return_U=return_U)

        bm = Materialised()
    with torch.no_grad():
        return tsde.sdeint(sde, y0, ts, bm=bm, method=method, dt=2.0 ** -5,
                           options={'cuda_graph': True} if graph else None).clone()


CASES = [
    # method, sde_type, problem kind, B, d, m, levy       (eligible: d a power of two >= 4, m in {8, 16, 32, 64})
    ('euler', 'ito', 'general', 1000, 32, 16, 'none'),
    ('heun', 'stratonovich', 'general', 777, 64, 16, 'none'),          # ragged last tile, two g operands
    ('midpoint', 'stratonovich', 'general', 513, 8, 8, 'none'),
    ('euler_heun', 'stratonovich', 'general', 300, 16, 32, 'none'),
    ('reversible_heun', 'stratonovich', 'general', 260, 16, 8, 'none'),
    ('srk', 'ito', 'additive', 515, 32, 16, 'space-time'),              # (W, U) weights
    ('euler', 'ito', 'general', 3, 4, 64, 'none'),                      # fewer tiles than stages
    ('euler', 'ito', 'general', 2100, 128, 64, 'none'),                 # one 32 KiB row per stage
]


def _launches(family):
    from torchsde_b200 import _cabi
    return _cabi.lib().tsde_kernel_launches(family)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64], ids=['f32', 'f64'])
@pytest.mark.parametrize('materialise', [False, True], ids=['counter', 'memory'])
@pytest.mark.parametrize('case', CASES, ids=lambda c: '-'.join(map(str, c)))
def test_tma_path_bit_identical(case, materialise, dtype, monkeypatch):
    method, sde_type, kind, B, d, m, levy = case
    monkeypatch.setenv('TSDE_GEN_TMA', '0')
    n_tma = _launches(1)
    base = _solve(method, sde_type, kind, B, d, m, dtype, levy, materialise)
    assert _launches(1) == n_tma, "TSDE_GEN_TMA=0 must keep the per-thread-load kernel"
    monkeypatch.setenv('TSDE_GEN_TMA', '2')
    n_cta = _launches(0)
    tma = _solve(method, sde_type, kind, B, d, m, dtype, levy, materialise)
    if d * m * base.element_size() <= 32 * 1024:  # (fp64 rows of 64 KiB exceed a stage: stays on the default kernel)
        assert _launches(1) > n_tma and _launches(0) == n_cta, "shape was not routed to the TMA-staged kernel"
    assert torch.isfinite(base).all()
    assert torch.equal(base, tma), f"max abs diff {(base - tma).abs().max().item()}"


def test_default_routing(monkeypatch):
    """Without TSDE_GEN_TMA (r02 routing): once the batch fills the pipeline m = 64 goes to the TMA-staged kernel, and so
    does m = 16 for tableaus with a single g operand (Euler); two-operand tableaus (Heun) at m = 16 and small batches stay
    on the per-thread-load kernel."""
    import ctypes
    from torchsde_b200 import _cabi
    monkeypatch.delenv('TSDE_GEN_TMA', raising=False)
    dev = torch.device('cuda')
    lib = _cabi.lib()
    key = torch.tensor([7], dtype=torch.int64, device=dev)
    for (B, D, M), expect_tma in (((65536, 32, 64), True), ((256, 32, 64), False), ((65536, 32, 16), True),
                                  ((256, 32, 16), False)):
        L = _cabi.make_launch(torch.float32, _cabi.NOISE_GENERAL, B, D, M)
        nz = _cabi.Noise()
        nz.source, nz.key, nz.cell_id, nz.n_cells, nz.h, nz.h_total = _cabi.SRC_COUNTER, key.data_ptr(), 5, 1, 0.01, 0.01
        y, f, g = torch.rand(B, D, device=dev), torch.rand(B, D, device=dev), torch.rand(B, D, M, device=dev)
        outs = []
        for mode in (None, '0'):
            if mode is None:
                monkeypatch.delenv('TSDE_GEN_TMA', raising=False)
            else:
                monkeypatch.setenv('TSDE_GEN_TMA', mode)
            o = torch.empty(B, D, device=dev)
            before = lib.tsde_kernel_launches(1)
            _cabi.check(lib.tsde_step_euler(ctypes.byref(L), ctypes.byref(nz), y.data_ptr(), f.data_ptr(),
                                            g.data_ptr(), 0.01, o.data_ptr()), 'tsde_step_euler')
            if mode is None:
                assert (lib.tsde_kernel_launches(1) > before) == expect_tma, (B, D, M)
            outs.append(o)
        assert torch.equal(outs[0], outs[1])
    # two g operands at m = 16: per-thread-load kernel by default
    B, D, M = 65536, 32, 16
    L = _cabi.make_launch(torch.float32, _cabi.NOISE_GENERAL, B, D, M)
    nz = _cabi.Noise()
    nz.source, nz.key, nz.cell_id, nz.n_cells, nz.h, nz.h_total = _cabi.SRC_COUNTER, key.data_ptr(), 5, 1, 0.01, 0.01
    e = [torch.rand(B, D, device=dev) for _ in range(4)]
    g2 = [torch.rand(B, D, M, device=dev) for _ in range(2)]
    monkeypatch.delenv('TSDE_GEN_TMA', raising=False)
    before = lib.tsde_kernel_launches(1)
    _cabi.check(lib.tsde_step_heun(ctypes.byref(L), ctypes.byref(nz), e[0].data_ptr(), e[1].data_ptr(), e[2].data_ptr(),
                                   g2[0].data_ptr(), g2[1].data_ptr(), 0.01, e[3].data_ptr()), 'tsde_step_heun')
    assert lib.tsde_kernel_launches(1) == before


def test_tma_path_in_cuda_graph(monkeypatch):
    monkeypatch.setenv('TSDE_GEN_TMA', '0')
    base = _solve('heun', 'stratonovich', 'general', 2048, 32, 16, torch.float32, 'none', False)
    monkeypatch.setenv('TSDE_GEN_TMA', '2')
    tma = _solve('heun', 'stratonovich', 'general', 2048, 32, 16, torch.float32, 'none', False, graph=True)
    assert torch.equal(base, tma)
