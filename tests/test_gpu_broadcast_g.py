"""Batch-broadcast diffusion operands (TSDE_FLAG_G_BROADCAST): an additive-noise SDE that returns
`sigma.expand(B, d, m)` (row stride 0) must give exactly the numbers of the same SDE returning a dense
`repeat` (what the reference's additive test problem does, tests/problems.py:113-116) — the tile kernels read
the shared (d, m) block instead of a densified copy, with the same chunk -> lane map and summation order."""
import pytest
import torch

from . import problems

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _tsde():
    import torchsde_b200
    return torchsde_b200


@pytest.mark.parametrize('method,sde_type,levy', [('srk', 'ito', 'space-time'), ('euler', 'ito', 'none'),
                                                  ('milstein', 'ito', 'none'), ('heun', 'stratonovich', 'none'),
                                                  ('midpoint', 'stratonovich', 'none'),
                                                  ('euler_heun', 'stratonovich', 'none'),
                                                  ('reversible_heun', 'stratonovich', 'none')])
@pytest.mark.parametrize('B,d,m', [(257, 32, 16), (64, 8, 4), (33, 5, 3), (1000, 16, 64), (2, 3, 2)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_expand_equals_repeat(method, sde_type, levy, B, d, m, dtype):
    tsde = _tsde()
    from torchsde_b200 import _cabi
    dense = problems.make('additive', d, m, sde_type, dtype=dtype, seed=3).to(DEV)
    bcast = problems.make('additive_expand', d, m, sde_type, dtype=dtype, seed=3).to(DEV)
    y0 = torch.rand(B, d, dtype=dtype, device=DEV)
    ts = torch.tensor([0.0, 0.125, 0.3, 0.5], dtype=dtype, device=DEV)
    out = []
    for sde in (dense, bcast):
        bm = tsde.BrownianInterval(0.0, 0.5, size=(B, m), dtype=dtype, device=DEV, entropy=11,
                                   levy_area_approximation=levy)
        with torch.no_grad():
            out.append(tsde.sdeint(sde, y0, ts, bm=bm, method=method, dt=2.0 ** -4))
    assert torch.equal(out[0], out[1])
    # ... and under a CUDA graph (the expanded view's storage lives in the graph's pool)
    bm = tsde.BrownianInterval(0.0, 0.5, size=(B, m), dtype=dtype, device=DEV, entropy=11,
                               levy_area_approximation=levy)
    with torch.no_grad():
        g = tsde.sdeint(bcast, y0, ts, bm=bm, method=method, dt=2.0 ** -4, options={'cuda_graph': True})
    if bm._root.kind == 2:
        assert torch.equal(out[0], g)
    assert _cabi.FLAG_G_BROADCAST == 1


def test_broadcast_views_that_must_be_densified():
    """Mixed dense / broadcast operands, a user Brownian object (memory source) and gradients through the solve all
    keep working: the host densifies whenever the flag cannot describe the launch."""
    tsde = _tsde()
    B, d, m = 65, 8, 4
    sde = problems.make('additive_expand', d, m, 'stratonovich', dtype=torch.float64, seed=5).to(DEV)
    ref = problems.make('additive', d, m, 'stratonovich', dtype=torch.float64, seed=5).to(DEV)
    y0 = torch.rand(B, d, dtype=torch.float64, device=DEV)
    ts = torch.tensor([0.0, 0.25, 0.5], dtype=torch.float64, device=DEV)

    class UserBM:  # memory source: the increments are tensors handed over by a duck-typed Brownian motion
        def __init__(self):
            self.inner = tsde.BrownianInterval(0.0, 0.5, size=(B, m), dtype=torch.float64, device=DEV, entropy=2)
            self.shape, self.dtype, self.device = self.inner.shape, self.inner.dtype, self.inner.device
            self.levy_area_approximation = 'none'

        def __call__(self, ta, tb=None, return_U=False, return_A=False):
            return self.inner(ta, tb)

    with torch.no_grad():
        a = tsde.sdeint(sde, y0, ts, bm=UserBM(), method='heun', dt=0.125)
        b = tsde.sdeint(ref, y0, ts, bm=UserBM(), method='heun', dt=0.125)
    assert torch.equal(a, b)
    # gradients through the solver (autograd nodes see dense operands)
    grads = []
    for s in (sde, ref):
        s.zero_grad()
        y = y0.clone().requires_grad_()
        bm = tsde.BrownianInterval(0.0, 0.5, size=(B, m), dtype=torch.float64, device=DEV, entropy=2)
        tsde.sdeint(s, y, ts, bm=bm, method='heun', dt=0.125)[-1].sum().backward()
        grads.append((y.grad.clone(), [p.grad.clone() for p in s.parameters()]))
    torch.testing.assert_close(grads[0][0], grads[1][0], rtol=1e-12, atol=1e-12)
    for p, q in zip(grads[0][1], grads[1][1]):
        torch.testing.assert_close(p, q, rtol=1e-12, atol=1e-12)
