"""Dry run of the host side on the CPU: the solver / adjoint / Brownian control flow that normally needs a GPU is
executed against a *recording stand-in* for the C library (every entry point checks its arity against the ctypes
binding, returns 0 and computes nothing).  Numbers are meaningless here — parity lives in the `-m gpu` tests — but every
Python path from `sdeint` / `sdeint_adjoint` down to the launch sites runs, so a broken call site, a wrong argument
count or a stale attribute fails in the CPU suite instead of on the GPU box.  This is test scaffolding: the product
itself has no CPU path (tests/test_host_logic.py::test_no_cpu_fallback)."""
import warnings

import pytest
import torch

import torchsde_b200 as tsde
from torchsde_b200 import _cabi
from torchsde_b200._brownian import interval as interval_mod
from . import problems


class _RecordingLib:
    def __init__(self):
        self.calls = {}

    def __getattr__(self, name):
        if name not in _cabi.SIGNATURES:
            raise AttributeError(name)
        arity = len(_cabi.SIGNATURES[name])

        def entry(*args):
            assert len(args) == arity, f"{name}: {len(args)} arguments, the C ABI declares {arity}"
            self.calls[name] = self.calls.get(name, 0) + 1
            return 0
        return entry

    def tsde_abi_version(self):
        return 1

    def tsde_kernel_launches(self, family):
        return 0


@pytest.fixture
def dry(monkeypatch):
    lib = _RecordingLib()
    monkeypatch.setattr(_cabi, '_lib', lib)
    monkeypatch.setattr(_cabi, 'lib', lambda: lib)
    monkeypatch.setattr(_cabi, 'require_cuda', lambda *a, **k: None)
    monkeypatch.setattr(interval_mod.BrownianInterval, '_require_cuda', lambda self: None)
    # let CPU-resident intervals bind a solver grid (the product only binds on CUDA devices)
    real_bind = interval_mod.BrownianInterval.bind_grid

    def bind_grid(self, bounds):
        device, self._device = self._device, torch.device('cuda')
        try:
            return real_bind(self, bounds)
        finally:
            self._device = device

    monkeypatch.setattr(interval_mod.BrownianInterval, 'bind_grid', bind_grid)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    # stand-ins for stream capture: the "captured" body simply runs once, replay does nothing
    import contextlib

    class _Stream:
        cuda_stream = 0

        def __init__(self, *a, **k):
            pass

        def wait_stream(self, other):
            pass

        def record_event(self):
            return object()

        def wait_event(self, event):
            pass

    class _Graph:
        replays = 0
        captures = 0

        def __init__(self):
            type(self).captures += 1

        def replay(self):
            type(self).replays += 1

    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, 'Stream', _Stream)
    monkeypatch.setattr(torch.cuda, 'stream', lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, 'CUDAGraph', _Graph)
    monkeypatch.setattr(torch.cuda, 'graph', lambda g, **k: contextlib.nullcontext())
    lib.graph_cls = _Graph
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        yield lib


CASES = [('gbm', 'ito', 'euler', 'none'), ('gbm', 'ito', 'milstein', 'none'), ('gbm', 'ito', 'srk', 'space-time'),
         ('gbm', 'stratonovich', 'milstein', 'none'), ('scalar', 'stratonovich', 'midpoint', 'none'),
         ('additive', 'ito', 'srk', 'space-time'), ('additive', 'ito', 'milstein', 'none'),
         ('general', 'ito', 'euler', 'none'), ('general', 'stratonovich', 'heun', 'none'),
         ('general', 'stratonovich', 'reversible_heun', 'none'), ('gbm', 'stratonovich', 'euler_heun', 'none'),
         ('gbm', 'stratonovich', 'reversible_heun', 'none'), ('general', 'stratonovich', 'log_ode', 'foster'),
         ('scalar', 'ito', 'srk', 'davie')]
TS = [0.0, 0.09375, 0.25]  # dyadic (exact in fp32, so the Brownian grid binds); the middle output lies between two
DT = 0.0625                # steps and exercises tsde_linear_interp


def _setup(kind, sde_type, levy):
    d, m = 3, {'gbm': 3, 'scalar': 1}.get(kind, 2)
    sde = problems.make(kind, d, m, sde_type, dtype=torch.float32)
    bm = tsde.BrownianInterval(0.0, TS[-1], size=(4, m), dtype=torch.float32, device='cpu',
                               levy_area_approximation=levy)
    return sde, torch.ones(4, d), bm


@pytest.mark.parametrize('kind,sde_type,method,levy', CASES)
def test_forward_paths(dry, kind, sde_type, method, levy):
    sde, y0, bm = _setup(kind, sde_type, levy)
    with torch.no_grad():
        ys = tsde.sdeint(sde, y0, TS, bm=bm, method=method, dt=DT)
    assert ys.shape == (3, 4, 3) and dry.calls and 'tsde_linear_interp' in dry.calls
    if method != 'log_ode':                      # gradients through the solver: every launch is an autograd node
        sde, y0, bm = _setup(kind, sde_type, levy)
        ys = tsde.sdeint(sde, y0.requires_grad_(), TS, bm=bm, method=method, dt=DT)
        ys.sum().backward()
        assert y0.grad is not None and y0.grad.shape == y0.shape


@pytest.mark.parametrize('kind,sde_type,method,levy', [c for c in CASES if c[2] not in ('log_ode',)])
def test_adjoint_paths(dry, kind, sde_type, method, levy):
    sde, y0, bm = _setup(kind, sde_type, levy)
    ys = tsde.sdeint_adjoint(sde, y0.requires_grad_(), TS, bm=bm, method=method, dt=DT)
    ys.sum().backward()
    assert y0.grad is not None and all(p.grad is not None for p in sde.parameters())
    if method == 'reversible_heun':
        assert dry.calls.get('tsde_adjoint_reversible_heun_a', 0) > 0 and dry.calls.get('tsde_adjoint_reversible_heun_b', 0) > 0


def test_adaptive_logqp_names_extra_and_queries(dry):
    sde, y0, bm = _setup('gbm', 'ito', 'none')
    with torch.no_grad():
        ys = tsde.sdeint(sde, y0, TS, bm=bm, method='euler', dt=DT, adaptive=True)
        assert ys.shape == (3, 4, 3) and dry.calls.get('tsde_adaptive_error_sumsq', 0) > 0

        class Renamed(torch.nn.Module):
            noise_type, sde_type = 'diagonal', 'ito'

            def drift(self, t, y):
                return -y

            def diffusion(self, t, y):
                return 0.1 + 0 * y

            def prior(self, t, y):
                return 0 * y

        ys, logqp = tsde.sdeint(Renamed(), y0, TS, method='euler', dt=DT, logqp=True,
                                names={'drift': 'drift', 'diffusion': 'diffusion', 'prior_drift': 'prior'},
                                bm=tsde.BrownianInterval(0., TS[-1], size=(4, 4), device='cpu'))
        assert ys.shape == (3, 4, 3) and logqp.shape == (2, 4)
        sde, y0, bm = _setup('gbm', 'stratonovich', 'none')
        ys, extra = tsde.sdeint(sde, y0, TS, bm=bm, method='reversible_heun', dt=DT, extra=True)
        assert len(extra) == 3
        # arbitrary queries: bridge below the bound grid, merges across cells, Levy area, derived objects
        bm = tsde.BrownianInterval(0.0, 1.0, size=(4, 2), device='cpu', levy_area_approximation='foster')
        w, u, a = bm(0.1, 0.7, return_U=True, return_A=True)
        assert w.shape == u.shape == (4, 2) and a.shape == (4, 2, 2)
        assert tsde.ReverseBrownian(bm)(-0.5, -0.25).shape == (4, 2)
        assert tsde.BrownianPath(t0=0.0, w0=torch.zeros(4, 2))(0.5).shape == (4, 2)
        assert tsde.BrownianTree(t0=0.0, w0=torch.zeros(4, 2), t1=1.0)(0.3).shape == (4, 2)
        assert tsde.brownian_interval_like(y0)(0.0, 0.5).shape == y0.shape
    for name in ('tsde_brownian_bridge', 'tsde_brownian_levy_area', 'tsde_brownian_h_to_u'):
        assert dry.calls.get(name, 0) > 0, name


def test_every_solver_entry_point_is_reached(dry):
    """Between them the dry runs drive (almost) the whole C ABI from Python; what is left is listed explicitly."""
    for kind, sde_type, method, levy in CASES:
        sde, y0, bm = _setup(kind, sde_type, levy)
        with torch.no_grad():
            tsde.sdeint(sde, y0, TS, bm=bm, method=method, dt=DT)
    sde, y0, bm = _setup('gbm', 'ito', 'none')
    with torch.no_grad():
        tsde.sdeint(sde, y0, TS, bm=bm, method='milstein', dt=DT, options={'grad_free': True})
    sde, y0, bm = _setup('general', 'stratonovich', 'none')
    tsde.sdeint_adjoint(sde, y0.requires_grad_(), TS, bm=bm, method='reversible_heun', dt=DT).sum().backward()
    sde, y0, bm = _setup('gbm', 'ito', 'none')
    with torch.no_grad():
        tsde.sdeint(sde, y0, TS, bm=bm, method='milstein', dt=DT, adaptive=True)
    bm = tsde.BrownianInterval(0.0, 1.0, size=(4, 2), device='cpu', levy_area_approximation='davie')
    bm(0.0, 0.5, return_U=True, return_A=True)
    bm(0.25, 0.75, return_U=True, return_A=True)      # covers pieces of two nodes: increments and areas are merged
    class Latent(torch.nn.Module):   # diagonal-noise logqp under no_grad: the fused KL-integrand launch
        noise_type, sde_type = 'diagonal', 'ito'
        f = staticmethod(lambda t, y: -y)
        h = staticmethod(lambda t, y: -0.5 * y)
        g = staticmethod(lambda t, y: 0.3 + 0.0 * y)
    with torch.no_grad():
        ys, logqp = tsde.sdeint(Latent(), torch.ones(4, 3), TS, method='euler', dt=DT, logqp=True,
                                bm=tsde.BrownianInterval(0.0, TS[-1], size=(4, 4), dtype=torch.float32, device='cpu'))
    assert ys.shape == (3, 4, 3) and logqp.shape == (2, 4)
    grid = tsde.BrownianInterval(0.0, 1.0, size=(4, 4), device='cpu', levy_area_approximation='foster', dt=0.25)
    w, u, a = grid(0.25, 0.5, return_U=True, return_A=True)   # one whole cell of the dt grid: the fused W/U/A launch
    assert w.shape == u.shape == (4, 4) and a.shape == (4, 4, 4)
    not_reached = set(_cabi.SIGNATURES) - set(dry.calls)
    assert not not_reached, sorted(not_reached)


def test_graph_plans_are_cached_and_follow_the_parameters(dry):
    """`options={'cuda_graph': True}` with stand-ins for stream capture: the plan is built once per (sde, shapes, grid,
    Brownian structure), replayed on the next call, and rebuilt when the SDE's parameters are replaced."""
    from torchsde_b200._core import graph
    sde, y0, _ = _setup('gbm', 'ito', 'none')

    def solve(method='milstein', static=True):
        bm = tsde.BrownianInterval(0.0, TS[-1], size=(4, 3), dtype=torch.float32, device='cpu')
        with torch.no_grad():
            return tsde.sdeint(sde, y0, TS, bm=bm, method=method, dt=DT,
                               options={'cuda_graph': True, 'static_output': static})

    ys = solve()
    assert ys.shape == (3, 4, 3)
    plans = graph._PLANS[sde]
    assert len(plans) == 1 and dry.graph_cls.replays == 1
    launches = sum(dry.calls.values())
    assert solve() is ys                                   # same plan, same static output buffer
    assert len(plans) == 1 and dry.graph_cls.replays == 2
    fresh = solve(static=False)                            # the default hands out a copy (results never alias)
    assert fresh is not ys and fresh.data_ptr() != ys.data_ptr() and len(plans) == 1
    dry.graph_cls.replays -= 1
    assert sum(dry.calls.values()) == launches             # a replay issues no new launches from Python
    solve('euler')
    assert len(plans) == 2
    sde.mu = torch.nn.Parameter(sde.mu.detach().clone())   # new storage: the old plan must not be replayed
    solve()
    assert len(plans) == 3
    for _ in range(4):                                     # the per-object cache is bounded
        sde.mu = torch.nn.Parameter(sde.mu.detach().clone())
        solve()
    assert len(plans) == graph.MAX_PLANS_PER_SDE
    # row-split chains and the captured reversible-Heun adjoint sweep run through the same machinery
    bm = tsde.BrownianInterval(0.0, TS[-1], size=(4, 3), dtype=torch.float32, device='cpu')
    with torch.no_grad():
        assert tsde.sdeint(sde, y0, TS, bm=bm, method='milstein', dt=DT,
                           options={'cuda_graph': True, 'row_split': 2}).shape == (3, 4, 3)
    sde2, y02, bm2 = _setup('gbm', 'stratonovich', 'none')
    ys = tsde.sdeint_adjoint(sde2, y02.requires_grad_(), TS, bm=bm2, method='reversible_heun', dt=DT,
                             options={'cuda_graph': True}, adjoint_options={'cuda_graph': True})
    ys.sum().backward()
    assert y02.grad is not None and all(p.grad is not None for p in sde2.parameters())


def test_graph_plans_die_with_their_sde(dry):
    """a plan pins a whole output series, so it must not outlive the SDE it was captured for, and the
    wrappers `sdeint` creates per call (logqp=True, names=...) must not defeat the cache."""
    import gc
    import weakref
    from torchsde_b200._core import graph
    from torchsde_b200._core import adjoint

    class Latent(torch.nn.Module):
        noise_type, sde_type = 'diagonal', 'stratonovich'

        def __init__(self):
            super().__init__()
            self.a = torch.nn.Parameter(torch.tensor([0.3, 0.2, 0.1]))

        def f(self, t, y):
            return -self.a * y

        def h(self, t, y):
            return -y

        def g(self, t, y):
            return 0.2 + 0.1 * torch.sigmoid(y)

    def bm_for(rows, m):
        return tsde.BrownianInterval(0.0, TS[-1], size=(rows, m), dtype=torch.float32, device='cpu')

    refs = []
    for _ in range(5):
        sde, y0, _ = _setup('gbm', 'ito', 'none')
        with torch.no_grad():
            tsde.sdeint(sde, y0, TS, bm=bm_for(4, 3), method='milstein', dt=DT, options={'cuda_graph': True})
        assert len(graph._PLANS[sde]) == 1
        refs.append(weakref.ref(sde))
        del sde
    gc.collect()
    assert all(r() is None for r in refs) and len(graph._PLANS) == 0
    # logqp=True wraps the user's SDE in a fresh SDELogqp per call: one capture, then replays
    sde = Latent()
    y0 = torch.full((4, 3), 0.1)
    captures = dry.graph_cls.captures
    for _ in range(4):
        with torch.no_grad():
            ys, logqp = tsde.sdeint(sde, y0, TS, bm=bm_for(4, 4), method='midpoint', dt=DT, logqp=True,
                                    options={'cuda_graph': True})
    assert dry.graph_cls.captures == captures + 1 and len(graph._PLANS[sde]) == 1
    for _ in range(3):
        # (output times on the step grid: only then can the backward sweep bind the forward pass's cells)
        out = tsde.sdeint_adjoint(sde, y0.clone().requires_grad_(), [0.0, 0.125, 0.25], bm=bm_for(4, 4),
                                  method='reversible_heun', dt=DT,
                                  logqp=True, options={'cuda_graph': True}, adjoint_options={'cuda_graph': True})
        out[0].sum().backward()
    assert len(adjoint._BWD_PLANS[sde]) == 1 and len(graph._PLANS[sde]) == 2
    ref = weakref.ref(sde)
    del sde, out, ys, logqp
    gc.collect()
    assert ref() is None and len(graph._PLANS) == 0 and len(adjoint._BWD_PLANS) == 0


@pytest.mark.parametrize('kind,sde_type,method,levy', [('gbm', 'ito', 'euler', 'none'), ('gbm', 'ito', 'milstein', 'none'),
                                                       ('gbm', 'ito', 'srk', 'space-time'),
                                                       ('general', 'stratonovich', 'heun', 'none')])
def test_empty_batch_flows_through(dry, kind, sde_type, method, levy):
    """Zero trajectories: the host side must still produce a (T, 0, d) series (the C entry points return early on
    empty launches; the reference itself handles B = 0 for most methods)."""
    d, m = 3, {'gbm': 3}.get(kind, 2)
    sde = problems.make(kind, d, m, sde_type, dtype=torch.float32)
    for bm in (None, tsde.BrownianInterval(0., TS[-1], size=(0, m), device='cpu', levy_area_approximation=levy)):
        with torch.no_grad():
            ys = tsde.sdeint(sde, torch.ones(0, d), TS, method=method, dt=DT, bm=bm)
        assert ys.shape == (3, 0, d)


@pytest.mark.parametrize('kind,sde_type,method,adjoint_method', [
    ('general', 'stratonovich', 'midpoint', None), ('gbm', 'stratonovich', 'heun', 'heun'),
    ('gbm', 'stratonovich', 'reversible_heun', 'adjoint_reversible_heun'),
    ('general', 'stratonovich', 'reversible_heun', 'adjoint_reversible_heun')])
def test_double_backward_flows_through(dry, kind, sde_type, method, adjoint_method):
    """create_graph=True through sdeint_adjoint: the generic adjoint re-enters the Function per interval (reference
    adjoint.py:97-113), the reversible pair switches to its differentiable sweep; both must produce second-order
    gradients w.r.t. y0 and the parameters."""
    sde, y0, bm = _setup(kind, sde_type, 'none')
    y0 = y0.clone().requires_grad_()
    ys = tsde.sdeint_adjoint(sde, y0, [0.0, 0.125, 0.25], bm=bm, method=method, adjoint_method=adjoint_method, dt=DT)
    params = list(sde.parameters())
    first = torch.autograd.grad((ys ** 2).sum(), [y0] + params, create_graph=True, allow_unused=True)
    assert all(g is not None for g in first)
    assert first[0].requires_grad
    second = torch.autograd.grad(sum((g ** 2).sum() for g in first), [y0] + params, allow_unused=True)
    assert second[0] is not None and second[0].shape == y0.shape


def test_adjoint_adaptive_reversible_pair_flows_through(dry):
    """adjoint_adaptive=True on the reversible pair: warns like the reference and runs the adaptive backward sweep."""
    import warnings as _w
    sde, y0, bm = _setup('general', 'stratonovich', 'none')
    y0 = y0.clone().requires_grad_()
    with _w.catch_warnings(record=True) as caught:
        _w.simplefilter('always')
        ys = tsde.sdeint_adjoint(sde, y0, [0.0, 0.125, 0.25], bm=bm, method='reversible_heun', dt=DT,
                                 adjoint_adaptive=True, adjoint_rtol=1e-1, adjoint_atol=1e-1)
        ys.sum().backward()
    assert any('does not save the time steps' in str(w.message) for w in caught)
    assert y0.grad is not None and all(p.grad is not None for p in sde.parameters())
    assert dry.calls.get('tsde_adaptive_error_sumsq', 0) > 0


@pytest.mark.parametrize('levy,entry', [('none', 'tsde_brownian_cells'), ('space-time', 'tsde_brownian_cells'),
                                        ('foster', 'tsde_brownian_cell_levy')])
@pytest.mark.parametrize('size', [(6, 8), (3, 2, 8)])
def test_whole_cell_queries_are_one_launch_each(dry, levy, entry, size):
    """dt-spaced queries on an interval created with `dt=` (cfg5's access pattern): every whole-cell query, in any
    order, is answered by exactly one launch — also W-only queries — and the outputs come in the caller-visible
    shape without a reshape (they are allocated in it)."""
    h = 0.125
    bm = tsde.BrownianInterval(0.0, 1.0, size=size, dtype=torch.float32, device='cpu', entropy=3, dt=h,
                               levy_area_approximation=levy)
    want_u, want_a = levy != 'none', levy == 'foster'
    for n, k in enumerate((0, 5, 2, 7)):
        before = dict(dry.calls)
        out = bm(k * h, (k + 1) * h, return_U=want_u, return_A=want_a)
        new = {name: c - before.get(name, 0) for name, c in dry.calls.items() if c != before.get(name, 0)}
        assert new == {entry: 1}, (levy, k, new)
        out = out if isinstance(out, tuple) else (out,)
        assert tuple(out[0].shape) == size and out[0].is_contiguous()
        if want_u:
            assert tuple(out[1].shape) == size
        if want_a:
            assert tuple(out[2].shape) == (*size, size[-1])
    # a query that is not a run of whole cells takes the general path (bridge), with more than one launch
    before = sum(dry.calls.values())
    bm(0.0625, 0.3, return_U=want_u, return_A=want_a)
    assert sum(dry.calls.values()) - before > 1
