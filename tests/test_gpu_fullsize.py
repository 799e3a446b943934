"""Parity at the sizes BASELINE.json names (the configurations the bench numbers are quoted on).

Every other GPU parity test runs toy sizes; here the CUDA path is run at cfg2 / cfg3 / cfg4 / cfg5 size and a
random sample of >= 256 trajectories is compared with the numpy oracle evaluated on exactly those global rows
(rows are independent Philox streams: `oracle/philox.normals(row_ids=...)`), plus a size-independent property
per config: the analytic GBM solution (reference tests/problems.py:39-64) for cfg2, adjoint == backprop through
the solver for cfg4, the additivity / U identities of tests/test_brownian_interval.py:261-288 for cfg5.
Tolerances (fp32): |got - ref| <= 1e-5 * max(|ref|, scale) as BASELINE.json's north star states ("within 1e-5
rel of the reference"); the device normals use SFU approximations that agree with the float64 oracle to ~1e-6
absolute per normal (csrc/philox.cuh).
"""
import math

import numpy as np
import pytest
import torch

from oracle import brownian as obm
from oracle import solvers
from . import helpers, problems

pytestmark = pytest.mark.gpu
N_SAMPLE = 256
DEV = 'cuda'


def _tsde():
    import torchsde_b200
    return torchsde_b200


def _free_gb():
    free, _ = torch.cuda.mem_get_info()
    return free / 2 ** 30


# ----------------------------------------------------------------------------------------------------------
# cfg2: Milstein Ito/diagonal, batch 65536, state 64, 1000 steps, fp32, full series — the headline workload
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('graph', [True, False], ids=['cuda_graph', 'eager'])
def test_cfg2_milstein_sampled_rows_vs_oracle(graph):
    tsde = _tsde()
    B, D, dt = 65536, 64, 2.0 ** -10
    T = 1000 if graph else 64
    if _free_gb() < (T + 1) * B * D * 4 / 2 ** 30 * 1.2 + 2:
        pytest.skip('not enough free device memory for the full series')
    sde = problems.GBMDiagonal(D, 'ito', seed=649, dtype=torch.float32).to(DEV)
    ts = (torch.arange(T + 1, dtype=torch.float32) * dt).to(DEV)
    y0 = torch.full((B, D), 0.1, device=DEV)
    bm = tsde.BrownianInterval(0.0, T * dt, size=(B, D), dtype=torch.float32, device=DEV, entropy=20260923)
    with torch.no_grad():
        ys = tsde.sdeint(sde, y0, ts, bm=bm, method='milstein', dt=dt, options={'cuda_graph': graph})
    assert ys.shape == (T + 1, B, D) and bool(torch.isfinite(ys[-1]).all())
    rows = helpers.sample_rows(B, N_SAMPLE, seed=1)
    got = ys[:, torch.from_numpy(rows).to(DEV)].cpu().numpy()
    sde_cpu = problems.GBMDiagonal(D, 'ito', seed=649, dtype=torch.float32)
    ref, _ = solvers.make('milstein', problems.NumpySDE(sde_cpu), helpers.oracle_grid_bm(bm, rows, D, np.float32, False),
                          dt).integrate(np.full((len(rows), D), 0.1, dtype=np.float32), ts.cpu().numpy())
    err = helpers.rel_err(got, ref)
    print(f"cfg2 graph={graph}: {len(rows)} rows x {T} steps, max rel err vs oracle {err:.3e}")
    assert err <= 1e-5, err
    if not graph:
        return
    # analytic solution of the GBM (reference tests/problems.py:55-64): y_t = y0 exp((mu - sigma^2/2) t + sigma W_t)
    # with W_t the SAME Brownian path (merged query over all cells).  Milstein has strong order 1.
    W = bm(0.0, T * dt)
    mu, sigma = sde.mu.detach(), sde.sigma.detach()
    exact = y0 * torch.exp((mu - 0.5 * sigma ** 2) * (T * dt) + sigma * W)
    rel = ((ys[-1] - exact).abs() / exact.abs())
    print(f"cfg2 analytic GBM check at t={T * dt:.4f}: mean rel err {rel.mean().item():.3e}, max {rel.max().item():.3e}")
    assert rel.mean().item() < 2e-3 and rel.max().item() < 5e-2
    # batch sharding at full size: rows [B/2, B) solved alone with row_offset reproduce the unsharded rows
    half = B // 2
    bm2 = tsde.BrownianInterval(0.0, T * dt, size=(half, D), dtype=torch.float32, device=DEV, entropy=20260923)
    bm2.shard_rows(half)
    ts_short = ts[:33].contiguous()
    bm_full = tsde.BrownianInterval(0.0, 32 * dt, size=(B, D), dtype=torch.float32, device=DEV, entropy=5)
    bm_half = tsde.BrownianInterval(0.0, 32 * dt, size=(half, D), dtype=torch.float32, device=DEV, entropy=5)
    bm_half.shard_rows(half)
    with torch.no_grad():
        a = tsde.sdeint(sde, y0, ts_short, bm=bm_full, method='milstein', dt=dt)
        b = tsde.sdeint(sde, y0[half:].contiguous(), ts_short, bm=bm_half, method='milstein', dt=dt)
    assert torch.equal(a[:, half:], b)


# ----------------------------------------------------------------------------------------------------------
# cfg3 (substituted): batch 8192, state 32, 16 Brownian channels, 500 steps
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('kind,sde_type,method', [('additive', 'ito', 'srk'), ('general', 'ito', 'euler'),
                                                  ('general', 'stratonovich', 'heun'),
                                                  ('additive_expand', 'ito', 'srk')])
def test_cfg3_sampled_rows_vs_oracle(kind, sde_type, method):
    tsde = _tsde()
    B, D, M, T, dt = 8192, 32, 16, 500, 2.0 ** -10
    sde = problems.make(kind, D, M, sde_type, dtype=torch.float32, seed=649).to(DEV)
    sde_cpu = problems.make(kind.replace('_expand', ''), D, M, sde_type, dtype=torch.float32, seed=649)
    levy = 'space-time' if method == 'srk' else 'none'
    ts = (torch.arange(T + 1, dtype=torch.float32) * dt).to(DEV)
    y0 = torch.full((B, D), 0.1, device=DEV)
    bm = tsde.BrownianInterval(0.0, T * dt, size=(B, M), dtype=torch.float32, device=DEV, entropy=77,
                               levy_area_approximation=levy)
    with torch.no_grad():
        ys = tsde.sdeint(sde, y0, ts, bm=bm, method=method, dt=dt, options={'cuda_graph': True})
    rows = helpers.sample_rows(B, N_SAMPLE, seed=2)
    got = ys[:, torch.from_numpy(rows).to(DEV)].cpu().numpy()
    ref, _ = solvers.make(method, problems.NumpySDE(sde_cpu),
                          helpers.oracle_grid_bm(bm, rows, M, np.float32, levy != 'none'), dt).integrate(
        np.full((len(rows), D), 0.1, dtype=np.float32), ts.cpu().numpy())
    scale = float(np.abs(ref).max())
    err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 0.1 * scale)))
    print(f"cfg3 {kind}/{method}: {len(rows)} rows x {T} steps, max rel err vs oracle {err:.3e} (scale {scale:.3g})")
    assert err <= 1e-5 * (3 if kind.startswith('general') else 1), err  # tanh: CUDA vs CPU libm differ by ulps


# ----------------------------------------------------------------------------------------------------------
# cfg4: sdeint_adjoint reversible Heun, Stratonovich/diagonal latent-SDE-like model, batch 32768, state 128
# (reduced number of steps: the sweep is the same per-step program)
# ----------------------------------------------------------------------------------------------------------
def _torch_reversible_heun(sde, y0, ts, dt, bm_np):
    """Differentiable restatement of the reversible Heun forward pass (methods/reversible_heun.py:58-73) in plain
    torch CPU ops on the sampled rows, fed the oracle's increments; autograd through it is the reference gradient."""
    f0, g0 = sde.f_and_g(ts[0], y0)
    y, z = y0, y0
    out = [y0]
    for k in range(len(ts) - 1):
        t0, t1 = ts[k], ts[k + 1]
        dW = torch.from_numpy(bm_np(float(t0), float(t1)))
        h = t1 - t0
        z1 = 2 * y - z + f0 * h + g0 * dW
        f1, g1 = sde.f_and_g(t1, z1)
        y = y + (f0 + f1) * (0.5 * h) + (g0 + g1) * (0.5 * dW)
        z, f0, g0 = z1, f1, g1
        out.append(y)
    return torch.stack(out)


def test_cfg4_adjoint_full_batch():
    tsde = _tsde()
    B, D, T, dt = 32768, 128, 24, 2.0 ** -10
    torch.backends.cuda.matmul.allow_tf32 = False
    sde = problems.LatentLike(D, hidden=128, seed=3).to(DEV)
    sde_cpu = problems.LatentLike(D, hidden=128, seed=3)
    ts = (torch.arange(T + 1, dtype=torch.float32) * dt)
    y0 = torch.full((B, D), 0.1, device=DEV, requires_grad=True)

    def loss_of(ys):
        return ys[-1].pow(2).sum(1).mean()

    grads = {}
    for mode, adjoint_opts in (('eager', {}), ('graph', {'cuda_graph': True})):
        bm = tsde.BrownianInterval(0.0, T * dt, size=(B, D), dtype=torch.float32, device=DEV, entropy=4242)
        sde.zero_grad()
        y0.grad = None
        ys = tsde.sdeint_adjoint(sde, y0, ts.to(DEV), bm=bm, method='reversible_heun',
                                 adjoint_method='adjoint_reversible_heun', dt=dt,
                                 options={'cuda_graph': mode == 'graph'}, adjoint_options=adjoint_opts)
        loss_of(ys).backward()
        grads[mode] = (ys.detach().clone(), y0.grad.clone(), [p.grad.clone() for p in sde.parameters()])
    ys_e, gy_e, gp_e = grads['eager']
    ys_g, gy_g, gp_g = grads['graph']
    assert torch.equal(ys_e, ys_g) and torch.equal(gy_e, gy_g)
    for a, b in zip(gp_e, gp_g):
        assert torch.equal(a, b)

    # (1) forward states and dL/dy0 of a sample of trajectories against the oracle path + CPU autograd
    rows = helpers.sample_rows(B, N_SAMPLE, seed=4)
    bm_np = helpers.oracle_grid_bm(bm, rows, D, np.float32, False)
    y0_cpu = torch.full((len(rows), D), 0.1, requires_grad=True)
    ys_ref = _torch_reversible_heun(sde_cpu, y0_cpu, ts, dt, bm_np)
    (ys_ref[-1].pow(2).sum(1).sum() / B).backward()
    idx = torch.from_numpy(rows).to(DEV)
    err_y = helpers.rel_err(ys_g[:, idx].cpu().numpy(), ys_ref.detach().numpy(), floor=1e-2)
    gref = y0_cpu.grad.numpy()
    gscale = float(np.abs(gref).max())
    err_g = float(np.max(np.abs(gy_g[idx].cpu().numpy() - gref)) / gscale)
    print(f"cfg4: ys max rel err {err_y:.3e}; dL/dy0 max err / max|grad| {err_g:.3e} on {len(rows)} rows")
    assert err_y <= 2e-5 and err_g <= 1e-4

    # (2) parameter gradients (a sum over ALL rows): the adjoint against backprop through the solver
    # (reference tests/test_adjoint.py:100-154), both on the GPU at full batch
    bm = tsde.BrownianInterval(0.0, T * dt, size=(B, D), dtype=torch.float32, device=DEV, entropy=4242)
    sde.zero_grad()
    y0.grad = None
    ys_bp = tsde.sdeint(sde, y0, ts.to(DEV), bm=bm, method='reversible_heun', dt=dt)
    loss_of(ys_bp).backward()
    torch.testing.assert_close(ys_bp.detach(), ys_g, rtol=1e-6, atol=1e-7)
    for (name, p), g_adj in zip(sde.named_parameters(), gp_g):
        denom = p.grad.abs().max().item() + 1e-12
        err = (p.grad - g_adj).abs().max().item() / denom
        print(f"cfg4 param {name}: adjoint vs backprop max err / max|grad| = {err:.3e}")
        assert err <= 2e-3, (name, err)
    torch.testing.assert_close(y0.grad, gy_g, rtol=1e-3, atol=1e-3 * gscale)


# ----------------------------------------------------------------------------------------------------------
# cfg5: BrownianInterval sweeps, 16 channels, 64 sequential dt-spaced queries, batch up to 2^20
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('levy,log2_b', [('none', 20), ('space-time', 20), ('foster', 17), ('davie', 14)])
def test_cfg5_brownian_queries_sampled_rows(levy, log2_b):
    tsde = _tsde()
    B, M, n_q = 1 << log2_b, 16, 64
    h = 2.0 ** -6
    bm = tsde.BrownianInterval(0.0, 1.0, size=(B, M), dtype=torch.float32, device=DEV, entropy=1147481649, dt=h,
                               levy_area_approximation=levy)
    rows = helpers.sample_rows(B, N_SAMPLE, seed=5)
    idx = torch.from_numpy(rows).to(DEV)
    have_h = levy != 'none'
    order = list(range(n_q))
    w_sum = torch.zeros(B, M, device=DEV)
    worst = 0.0
    for k in order:
        ta, tb = k * h, (k + 1) * h
        if levy == 'none':
            W = bm(ta, tb)
        elif levy == 'space-time':
            W, U = bm(ta, tb, return_U=True)
        else:
            W, U, A = bm(ta, tb, return_U=True, return_A=True)
        w_sum += W
        if k % 9 != 0 and k != n_q - 1:
            continue
        grid = bm._root
        Wr, Hr = obm.cell(bm._key, (grid.cell_base + k) & helpers.MASK64, h, len(rows), M, np.float32, have_h,
                          row_ids=rows)
        worst = max(worst, float(np.abs(W[idx].cpu().numpy() - Wr).max()) / math.sqrt(h))
        if have_h:
            Ur = obm.h_to_u(Wr, Hr, h)
            worst = max(worst, float(np.abs(U[idx].cpu().numpy() - Ur).max()) / (h * math.sqrt(h)))
        if levy in ('davie', 'foster'):
            Ar = obm.davie_foster(Wr, Hr, h, levy == 'foster',
                                  obm.levy_noise(bm._key, (grid.cell_base + k) & helpers.MASK64, len(rows), M,
                                                 np.float32, row_ids=rows))
            worst = max(worst, float(np.abs(A[idx].cpu().numpy() - Ar).max()) / h)
            assert float((A + A.transpose(1, 2)).abs().max()) <= 1e-6 * h  # antisymmetric up to fma contraction
    print(f"cfg5 levy={levy} B=2^{log2_b}: max err vs oracle, in units of the increment's std: {worst:.3e}")
    assert worst <= 1e-5
    # additivity over the whole interval (tests/test_brownian_interval.py:261-288), all rows
    total = bm(0.0, 1.0)
    assert float((w_sum - total).abs().max()) <= 2e-5
    # random-order re-queries return the same numbers (determinism, :110-161)
    for k in np.random.RandomState(0).permutation(n_q)[:8]:
        W2 = bm(k * h, (k + 1) * h)
        Wr, _ = obm.cell(bm._key, (bm._root.cell_base + int(k)) & helpers.MASK64, h, len(rows), M, np.float32, False,
                         row_ids=rows)
        assert float(np.abs(W2[idx].cpu().numpy() - Wr).max()) / math.sqrt(h) <= 1e-5
