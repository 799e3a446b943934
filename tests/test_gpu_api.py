"""API-surface behaviour of sdeint on the GPU (reference sdeint.py:27-112 argument handling)."""
import pytest
import torch

from . import problems

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _tsde():
    import torchsde_b200
    return torchsde_b200


@pytest.fixture(autouse=True)
def _no_grad():
    with torch.no_grad():
        yield


def test_argument_forms_agree():
    tsde = _tsde()
    B, d = 12, 8
    sde = problems.GBMDiagonal(d, 'ito', seed=1, dtype=torch.float32).to(DEV)
    y0 = torch.rand(B, d, device=DEV) + 0.1

    def run(y, ts, dt, **kw):
        bm = tsde.BrownianInterval(0.0, 0.5, size=(B, d), dtype=torch.float32, device=DEV, entropy=5)
        return tsde.sdeint(sde, y, ts, bm=bm, method='milstein', dt=dt, **kw)

    base = run(y0, torch.tensor([0.0, 0.25, 0.5], device=DEV), 0.125)
    assert torch.equal(base, run(y0, [0.0, 0.25, 0.5], 0.125))                       # list of floats
    assert torch.equal(base, run(y0, (0.0, 0.25, 0.5), torch.tensor(0.125)))          # tuple, 0-d tensor dt
    nc = y0.t().contiguous().t()                                                      # non-contiguous y0
    assert not nc.is_contiguous() and torch.equal(base, run(nc, [0.0, 0.25, 0.5], 0.125))
    ts64 = torch.tensor([0.0, 0.25, 0.5], dtype=torch.float64, device=DEV)           # fp64 grid, fp32 state
    assert torch.equal(base, run(y0, ts64, 0.125))
    ys, extra = run(y0, [0.0, 0.25, 0.5], 0.125, extra=True)
    assert extra == () and torch.equal(ys, base)
    with pytest.warns(UserWarning, match='Unexpected arguments'):
        assert torch.equal(base, run(y0, [0.0, 0.25, 0.5], 0.125, foo=3))
    with pytest.raises(ValueError, match='must not require gradient'):
        run(y0, [0.0, 0.25, 0.5], torch.tensor(0.125, requires_grad=True))


def test_names_and_default_method_and_default_bm():
    tsde = _tsde()
    B, d = 6, 4

    class Custom(torch.nn.Module):
        noise_type, sde_type = 'diagonal', 'ito'

        def forward(self, t, y):
            return -y

        def diffusion(self, t, y):
            return 0.3 * y

    y0 = torch.ones(B, d, device=DEV)
    ys = tsde.sdeint(Custom(), y0, [0.0, 0.1, 0.2], dt=0.05, names={'drift': 'forward', 'diffusion': 'diffusion'})
    assert ys.shape == (3, B, d) and torch.isfinite(ys).all()   # default method srk, default bm (space-time)
    strat = problems.TanhGeneral(d, 3, 'stratonovich', dtype=torch.float32).to(DEV)
    ys = tsde.sdeint(strat, y0, [0.0, 0.1], dt=0.05)            # default method midpoint
    assert ys.shape == (2, B, d)
    # two solves with default bm differ (fresh entropy from numpy's global RNG, brownian_interval.py:489-490)
    a = tsde.sdeint(strat, y0, [0.0, 0.1], dt=0.05)
    b = tsde.sdeint(strat, y0, [0.0, 0.1], dt=0.05)
    assert not torch.equal(a, b)


def test_dt_larger_than_output_spacing_and_single_interval():
    """dt > spacing of ts: several outputs interpolated inside one step (base_solver.py:114-147)."""
    tsde = _tsde()
    B, d = 5, 4
    sde = problems.GBMDiagonal(d, 'ito', seed=2, dtype=torch.float64).to(DEV)
    y0 = torch.full((B, d), 0.7, dtype=torch.float64, device=DEV)
    bm = tsde.BrownianInterval(0.0, 0.2, size=(B, d), dtype=torch.float64, device=DEV, entropy=1)
    ts = torch.tensor([0.0, 0.01, 0.02, 0.03, 0.2], dtype=torch.float64, device=DEV)
    ys = tsde.sdeint(sde, y0, ts, bm=bm, method='euler', dt=0.05)
    # rows 1..3 lie on the segment between y0 and the first step's end point
    step_end = y0 + (ys[1] - y0) * (0.05 / 0.01)
    torch.testing.assert_close(ys[2], y0 + (step_end - y0) * (0.02 / 0.05), rtol=1e-10, atol=1e-12)
    torch.testing.assert_close(ys[3], y0 + (step_end - y0) * (0.03 / 0.05), rtol=1e-10, atol=1e-12)


def test_step_operator_contract():
    """Reference per-step operator contract: solver.step(t0, t1, y0, extra0) -> (y1, extra1) and class attrs."""
    tsde = _tsde()
    from torchsde_b200._core import methods
    from torchsde_b200._core.base_sde import ForwardSDE
    B, d = 4, 8
    sde = ForwardSDE(problems.GBMDiagonal(d, 'stratonovich', seed=2, dtype=torch.float32).to(DEV))
    bm = tsde.BrownianInterval(0.0, 1.0, size=(B, d), dtype=torch.float32, device=DEV, entropy=3)
    cls = methods.select('reversible_heun', 'stratonovich')
    assert cls.sde_type == 'stratonovich' and cls.weak_order == 1.0 and 'general' in cls.noise_types
    solver = cls(sde=sde, bm=bm, dt=0.1, adaptive=False, rtol=1e-5, atol=1e-4, dt_min=1e-5, options={})
    assert solver.strong_order == 0.5 and 'ReversibleHeun' in repr(solver)
    y0 = torch.full((B, d), 0.5, device=DEV)
    extra0 = solver.init_extra_solver_state(torch.tensor(0.0, device=DEV), y0)
    y1, extra1 = solver.step(torch.tensor(0.0), torch.tensor(0.1), y0, extra0)
    assert y1.shape == y0.shape and len(extra1) == 3
    # equals the first step of a full solve on the same Brownian path
    ys = tsde.sdeint(sde._base_sde, y0, [0.0, 0.1], bm=bm, method='reversible_heun', dt=0.1)
    torch.testing.assert_close(ys[1], y1, rtol=1e-6, atol=1e-7)


def test_wrong_device_or_dtype_errors():
    tsde = _tsde()
    sde = problems.GBMDiagonal(4, 'ito', dtype=torch.float32).to(DEV)
    y0 = torch.ones(3, 4, device=DEV, dtype=torch.float16)
    with pytest.raises((ValueError, RuntimeError)):
        tsde.sdeint(sde.half(), y0, [0.0, 0.1], dt=0.05, method='euler')
