"""sdeint_adjoint (reversible_heun / adjoint_reversible_heun) on the GPU."""
import numpy as np
import pytest
import torch

from . import helpers, problems

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _tsde():
    import torchsde_b200
    return torchsde_b200


@pytest.mark.parametrize('path', helpers.golden_files('adjoint_'), ids=helpers.case_id)
def test_adjoint_golden_replay(path):
    """Gradients of the reference's sdeint_adjoint on identical increments."""
    tsde = _tsde()
    case = helpers.load(path)
    sde = problems.make(str(case['kind']), int(case['d']), int(case['m']), 'stratonovich', dtype=torch.float64,
                        seed=3).to(DEV)
    y0 = torch.from_numpy(case['y0']).to(DEV).requires_grad_(True)
    ts = torch.from_numpy(case['ts']).to(DEV)
    Ws = [torch.from_numpy(w).to(DEV) for w in case['W']]
    bm = problems.ReplayBM(case['ta'], case['tb'], Ws)
    ys = tsde.sdeint_adjoint(sde, y0, ts, bm=bm, method='reversible_heun',
                             adjoint_method='adjoint_reversible_heun', dt=float(case['dt']))
    np.testing.assert_allclose(ys.detach().cpu().numpy(), case['ys'], rtol=1e-12, atol=1e-13)
    (ys * torch.from_numpy(case['weights']).to(DEV)).sum().backward()
    np.testing.assert_allclose(y0.grad.cpu().numpy(), case['grad_y0'], rtol=1e-9, atol=1e-11)
    for n, p in sde.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), case['grad.' + n], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize('kind,d,m', [('gbm', 16, 16), ('general', 4, 8)])
def test_adjoint_counter_path_matches_replay(kind, d, m):
    """Backward regenerates the forward's Brownian cells from the counter, in reverse: gradients
    equal those obtained when the same increments are served from memory."""
    tsde = _tsde()
    B = 24
    sde = problems.make(kind, d, m, 'stratonovich', dtype=torch.float64, seed=6).to(DEV)
    bm_m = d if kind == 'gbm' else m
    ts = torch.tensor([0.0, 0.25, 0.5], dtype=torch.float64, device=DEV)
    dt = 2.0 ** -4
    y_a = torch.full((B, d), 0.4, dtype=torch.float64, device=DEV).requires_grad_(True)
    bm = tsde.BrownianInterval(0.0, 0.5, size=(B, bm_m), dtype=torch.float64, device=DEV, entropy=17)
    ys = tsde.sdeint_adjoint(sde, y_a, ts, bm=bm, method='reversible_heun', dt=dt)
    ys.pow(2).sum().backward()
    g_fast = [y_a.grad.clone()] + [p.grad.clone() for p in sde.parameters()]
    for p in sde.parameters():
        p.grad = None

    class Mat:
        shape = bm.shape
        levy_area_approximation = 'none'

        def __call__(self, ta, tb=None, return_U=False, return_A=False):
            return bm(ta, tb)

    y_b = torch.full((B, d), 0.4, dtype=torch.float64, device=DEV).requires_grad_(True)
    ys2 = tsde.sdeint_adjoint(sde, y_b, ts, bm=Mat(), method='reversible_heun', dt=dt)
    assert torch.equal(ys, ys2)
    ys2.pow(2).sum().backward()
    g_slow = [y_b.grad] + [p.grad for p in sde.parameters()]
    for a, b in zip(g_fast, g_slow):
        torch.testing.assert_close(a, b, rtol=1e-12, atol=1e-13)


def test_reversibility():
    """Reference tests/test_sdeint.py:219-252: forward with reversible Heun, then the forward solver
    again on the reversed Brownian motion with negated extras reconstructs ys."""
    tsde = _tsde()
    B, d = 8, 6
    sde = problems.GBMDiagonal(d, 'stratonovich', seed=2, dtype=torch.float64).to(DEV)
    y0 = torch.full((B, d), 0.3, dtype=torch.float64, device=DEV)
    ts = torch.linspace(0.0, 1.0, 5, dtype=torch.float64, device=DEV)
    bm = tsde.BrownianInterval(0.0, 1.0, size=(B, d), dtype=torch.float64, device=DEV, entropy=5)
    ys, (f, g, z) = tsde.sdeint(sde, y0, ts, bm=bm, method='reversible_heun', dt=0.125, extra=True)

    class Neg(torch.nn.Module):
        noise_type, sde_type = 'diagonal', 'stratonovich'

        def f_and_g(self, t, y):
            return -sde.f(-t, y), -sde.g(-t, y)

    back, _ = tsde.sdeint(Neg(), ys[-1], -ts.flip(0), bm=tsde.ReverseBrownian(bm), method='reversible_heun',
                          dt=0.125, extra=True, extra_solver_state=(-f, -g, z))
    torch.testing.assert_close(back.flip(0), ys, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('kind,d,m', [('gbm', 16, 16), ('general', 4, 8)])
def test_adjoint_cuda_graph_equals_eager(kind, d, m):
    """Forward and backward sweeps replayed from CUDA graphs give the eager results bit for bit,
    also on a second call with a different Brownian key and different y0 (static buffers refreshed)."""
    tsde = _tsde()
    B = 32
    sde = problems.make(kind, d, m, 'stratonovich', dtype=torch.float32, seed=6).to(DEV)
    bm_m = d if kind == 'gbm' else m
    ts = torch.tensor([0.0, 0.25, 0.5], device=DEV)
    dt = 2.0 ** -4

    def run(graph, entropy, fill):
        for p in sde.parameters():
            p.grad = None
        y = torch.full((B, d), fill, device=DEV).requires_grad_(True)
        bm = tsde.BrownianInterval(0.0, 0.5, size=(B, bm_m), dtype=torch.float32, device=DEV, entropy=entropy)
        ys = tsde.sdeint_adjoint(sde, y, ts, bm=bm, method='reversible_heun', dt=dt,
                                 options={'cuda_graph': graph}, adjoint_options={'cuda_graph': graph})
        ys.pow(2).sum().backward()
        return ys.detach().clone(), y.grad.clone(), [p.grad.clone() for p in sde.parameters()]

    for entropy, fill in ((3, 0.4), (4, 0.7), (3, 0.4)):
        e = run(False, entropy, fill)
        g = run(True, entropy, fill)
        assert torch.equal(e[0], g[0]) and torch.equal(e[1], g[1])
        for a, b in zip(e[2], g[2]):
            assert torch.equal(a, b)


@pytest.mark.parametrize('path', helpers.golden_files('genadj_'), ids=helpers.case_id)
def test_generic_adjoint_golden_replay(path):
    """sdeint_adjoint through AdjointSDE (default adjoint methods): gradients of the reference on identical
    increments."""
    tsde = _tsde()
    case = helpers.load(path)
    sde = problems.make(str(case['kind']), int(case['d']), int(case['m']), str(case['sde_type']),
                        dtype=torch.float64, seed=int(case['seed'])).to(DEV)
    y0 = torch.from_numpy(case['y0']).to(DEV).requires_grad_(True)
    ts = torch.from_numpy(case['ts']).to(DEV)
    Ws = [torch.from_numpy(w).to(DEV) for w in case['W']]
    Us = [torch.from_numpy(u).to(DEV) for u in case['U']] if 'U' in case else None
    bm = problems.ReplayBM(case['ta'], case['tb'], Ws, Us, levy='space-time' if Us is not None else 'none')
    ys = tsde.sdeint_adjoint(sde, y0, ts, bm=bm, method=str(case['method']),
                             adjoint_method=str(case['adjoint_method']) or None, dt=float(case['dt']))
    np.testing.assert_allclose(ys.detach().cpu().numpy(), case['ys'], rtol=1e-11, atol=1e-13)
    (ys * torch.from_numpy(case['weights']).to(DEV)).sum().backward()
    np.testing.assert_allclose(y0.grad.cpu().numpy(), case['grad_y0'], rtol=1e-8, atol=1e-10)
    for n, p in sde.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), case['grad.' + n], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize('path', helpers.golden_files('backprop_'), ids=helpers.case_id)
def test_backprop_through_solver_golden_replay(path):
    """Plain sdeint under autograd (every tableau launch an autograd node): gradients equal those of
    backpropagating through the REFERENCE solver on identical increments."""
    tsde = _tsde()
    case = helpers.load(path)
    sde = problems.make(str(case['kind']), int(case['d']), int(case['m']), str(case['sde_type']),
                        dtype=torch.float64, seed=int(case['seed'])).to(DEV)
    y0 = torch.from_numpy(case['y0']).to(DEV).requires_grad_(True)
    ts = torch.from_numpy(case['ts']).to(DEV)
    Ws = [torch.from_numpy(w).to(DEV) for w in case['W']]
    Us = [torch.from_numpy(u).to(DEV) for u in case['U']] if 'U' in case else None
    bm = problems.ReplayBM(case['ta'], case['tb'], Ws, Us, levy='space-time' if Us is not None else 'none')
    ys = tsde.sdeint(sde, y0, ts, bm=bm, method=str(case['method']), dt=float(case['dt']),
                     options={'grad_free': True} if bool(case['grad_free']) else None)
    assert ys.requires_grad
    np.testing.assert_allclose(ys.detach().cpu().numpy(), case['ys'], rtol=1e-11, atol=1e-13)
    (ys * torch.from_numpy(case['weights']).to(DEV)).sum().backward()
    np.testing.assert_allclose(y0.grad.cpu().numpy(), case['grad_y0'], rtol=1e-9, atol=1e-11)
    for n, p in sde.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), case['grad.' + n], rtol=1e-8, atol=1e-10)


def test_backprop_counter_path_matches_adjoint():
    """Reference tests/test_adjoint.py:100-154 (test_against_sdeint): the reversible adjoint's gradients equal
    backprop through the solver, here both on the counter-based Brownian motion."""
    tsde = _tsde()
    B, d = 16, 8
    sde = problems.GBMDiagonal(d, 'stratonovich', seed=4, dtype=torch.float64).to(DEV)
    ts = torch.tensor([0.0, 0.25, 0.5], dtype=torch.float64, device=DEV)
    grads = []
    for fn in (tsde.sdeint, tsde.sdeint_adjoint):
        for p in sde.parameters():
            p.grad = None
        y0 = torch.full((B, d), 0.3, dtype=torch.float64, device=DEV).requires_grad_(True)
        bm = tsde.BrownianInterval(0.0, 0.5, size=(B, d), dtype=torch.float64, device=DEV, entropy=8)
        ys = fn(sde, y0, ts, bm=bm, method='reversible_heun', dt=2.0 ** -4)
        ys.pow(2).sum().backward()
        grads.append([y0.grad.clone()] + [p.grad.clone() for p in sde.parameters()])
    for a, b in zip(*grads):
        torch.testing.assert_close(a, b, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize('path', helpers.golden_files('gradgrad_'), ids=helpers.case_id)
def test_double_backward_golden_replay(path):
    """Double backward through sdeint_adjoint (reference adjoint.py:97-113: the backward pass re-enters the Function,
    so the second-order gradient is the continuous adjoint of the adjoint solve): first- and second-order gradients
    against the reference's on identical increments."""
    tsde = _tsde()
    case = helpers.load(path)
    dev = torch.device('cuda')
    kind, d, m = str(case['kind']), int(case['d']), int(case['m'])
    sde = problems.make(kind, d, m, str(case['sde_type']), dtype=torch.float64, seed=int(case['seed'])).to(dev)
    params = list(sde.parameters())
    y0 = torch.from_numpy(case['y0']).to(dev).requires_grad_(True)
    ts = torch.from_numpy(case['ts']).to(dev)
    bm = helpers.replay_torch(case, dev)
    ys = tsde.sdeint_adjoint(sde, y0, ts, bm=bm, method=str(case['method']),
                             adjoint_method=str(case['adjoint_method']) or None, dt=float(case['dt']))
    np.testing.assert_allclose(ys.detach().cpu().numpy(), case['ys'], rtol=1e-11, atol=1e-12)
    loss = (ys * torch.from_numpy(case['w1']).to(dev)).sum()
    first = torch.autograd.grad(loss, [y0] + params, create_graph=True, allow_unused=True)
    first = [torch.zeros_like(x) if g is None else g for g, x in zip(first, [y0] + params)]
    np.testing.assert_allclose(first[0].detach().cpu().numpy(), case['first_y0'], rtol=1e-8, atol=1e-10)
    for (n, _), g in zip(sde.named_parameters(), first[1:]):
        np.testing.assert_allclose(g.detach().cpu().numpy(), case['first.' + n], rtol=1e-8, atol=1e-10, err_msg=n)
    w2 = [torch.linspace(1.0, 2.0, g.numel(), dtype=torch.float64, device=dev).reshape(g.shape) for g in first]
    second = sum((g * w).sum() for g, w in zip(first, w2))
    gg = torch.autograd.grad(second, [y0] + params, allow_unused=True)
    gg = [torch.zeros_like(x) if g is None else g for g, x in zip(gg, [y0] + params)]
    np.testing.assert_allclose(gg[0].cpu().numpy(), case['second_y0'], rtol=1e-7, atol=1e-9)
    for (n, _), g in zip(sde.named_parameters(), gg[1:]):
        np.testing.assert_allclose(g.cpu().numpy(), case['second.' + n], rtol=1e-7, atol=1e-9, err_msg=n)


@pytest.mark.parametrize('kind,d,m', [('gbm', 4, 4), ('general', 3, 2)])
def test_double_backward_reversible_pair(kind, d, m):
    """The reversible pair cannot be double-backwarded in the reference at all (its re-entered Function finds no saved
    solver state: reversible_heun.py:93-96 raises).  Here the backward sweep falls back to differentiable torch
    operations when create_graph=True; since reversible Heun's adjoint is the exact gradient of the discrete solve,
    its second-order gradients must equal those obtained by double-backpropagating through plain `sdeint`."""
    tsde = _tsde()
    dev = torch.device('cuda')
    sde = problems.make(kind, d, m, 'stratonovich', dtype=torch.float64, seed=9).to(dev)
    params = list(sde.parameters())
    B = 5
    y0 = (0.2 + 0.3 * torch.rand(B, d, dtype=torch.float64, device=dev)).requires_grad_(True)
    ts = torch.tensor([0.0, 0.125, 0.25], dtype=torch.float64, device=dev)
    bm_m = d if kind == 'gbm' else m
    out = []
    for adjoint in (True, False):
        bm = tsde.BrownianInterval(0.0, 0.25, size=(B, bm_m), dtype=torch.float64, device=dev, entropy=21)
        if adjoint:
            ys = tsde.sdeint_adjoint(sde, y0, ts, bm=bm, method='reversible_heun',
                                     adjoint_method='adjoint_reversible_heun', dt=2.0 ** -4)
        else:
            ys = tsde.sdeint(sde, y0, ts, bm=bm, method='reversible_heun', dt=2.0 ** -4)
        loss = (ys ** 2).sum()
        first = torch.autograd.grad(loss, [y0] + params, create_graph=True)
        second = sum((g ** 2).sum() for g in first)
        out.append(([g.detach() for g in first], torch.autograd.grad(second, [y0] + params)))
    for a, b in zip(out[0][0], out[1][0]):
        torch.testing.assert_close(a, b, rtol=1e-8, atol=1e-10)
    for a, b in zip(out[0][1], out[1][1]):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize('path', helpers.golden_files('adjadaptive_'), ids=helpers.case_id)
def test_adjoint_adaptive_reversible_pair_golden_replay(path):
    """adjoint_adaptive=True with the reversible pair: the reference warns and integrates the adjoint adaptively
    (adjoint.py:245-249); the backward engine must take the same accept / reject decisions (its queries are replayed
    from the reference's own log: an unknown (ta, tb) raises) and return the same gradients."""
    import warnings
    tsde = _tsde()
    case = helpers.load(path)
    dev = torch.device('cuda')
    kind, d, m = str(case['kind']), int(case['d']), int(case['m'])
    sde = problems.make(kind, d, m, 'stratonovich', dtype=torch.float64, seed=int(case['seed'])).to(dev)
    y0 = torch.from_numpy(case['y0']).to(dev).requires_grad_(True)
    ts = torch.from_numpy(case['ts']).to(dev)
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter('always')
        ys = tsde.sdeint_adjoint(sde, y0, ts, bm=helpers.replay_torch(case, dev), method='reversible_heun',
                                 adjoint_method='adjoint_reversible_heun', dt=float(case['dt']), adjoint_adaptive=True,
                                 adjoint_rtol=float(case['rtol']), adjoint_atol=float(case['atol']),
                                 dt_min=float(case['dt_min']))
        (ys * torch.from_numpy(case['weights']).to(dev)).sum().backward()
    assert any('does not save the time steps' in str(w.message) for w in caught)
    np.testing.assert_allclose(ys.detach().cpu().numpy(), case['ys'], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(y0.grad.cpu().numpy(), case['grad_y0'], rtol=1e-8, atol=1e-10)
    for n, p in sde.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), case['grad.' + n], rtol=1e-8, atol=1e-10, err_msg=n)
