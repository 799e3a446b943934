"""Fixed-step SDE solvers on the CPU: numpy restatement of the reference's hot path.

Test infrastructure only (see oracle/__init__.py).  Restates
  * the time loop + linear interpolation   torchsde/_core/base_solver.py:92-149, interp.py:15-18
  * every step tableau                      torchsde/_core/methods/*.py
  * the products g.v                        torchsde/_core/base_sde.py:98-102, misc.py:62-63
with the reference's evaluation order.  `sde` is any object with numpy callables
f(t, y), g(t, y) (and for Milstein gdg(t, y, v2) = vjp_y(g; g*v2), base_sde.py:127-155);
`bm(ta, tb, return_U=False)` returns numpy increments.  t is passed as a numpy scalar of ts' dtype.
"""
import numpy as np


def _prod(g, v):
    # diagonal: g*v (base_sde.py:98-99); otherwise batched matrix-vector product (misc.py:62-63)
    if g.ndim == 2:
        return g * v
    return np.einsum('bdm,bm->bd', g, v).astype(g.dtype)


def _sc(x, like):
    return like.dtype.type(x)


class Solver:
    want_u = False

    def __init__(self, sde, bm, dt, options=None):
        self.sde, self.bm, self.dt = sde, bm, dt
        self.options = options or {}

    def init_extra(self, t0, y0):
        return ()

    # base_solver.py:92-149 (fixed-step branch)
    def integrate(self, y0, ts, extra0=None):
        ts = np.asarray(ts)
        step_size = self.dt
        prev_t = curr_t = ts[0]
        prev_y = curr_y = y0
        extra = self.init_extra(ts[0], y0) if extra0 is None else extra0
        ys = [y0]
        for out_t in ts[1:]:
            while curr_t < out_t:
                next_t = min(curr_t + step_size, ts[-1])
                prev_t, prev_y = curr_t, curr_y
                curr_y, extra = self.step(curr_t, next_t, curr_y, extra)
                curr_t = next_t
            ys.append(linear_interp(prev_t, prev_y, curr_t, curr_y, out_t))
        return np.stack(ys, axis=0), extra


def update_step_size(error_estimate, prev_step_size, safety=0.9, facmin=0.2, facmax=1.4, prev_error_ratio=None):
    # adaptive_stepping.py:21-39
    if error_estimate > 1:
        pfactor = 0
        ifactor = 1 / 1.5
    else:
        pfactor = 0.13
        ifactor = 1 / 4.5
    error_ratio = safety / error_estimate
    if prev_error_ratio is None:
        prev_error_ratio = error_ratio
    factor = error_ratio ** ifactor * (error_ratio / prev_error_ratio) ** pfactor
    if error_estimate <= 1:
        prev_error_ratio = error_ratio
        facmin = 1.0
    factor = min(facmax, max(facmin, factor))
    return prev_step_size * factor, prev_error_ratio


def compute_error(y11, y12, rtol, atol, eps=1e-7):
    # adaptive_stepping.py:42-76
    tol = np.maximum(_sc(rtol, y11) * np.maximum(np.abs(y11), np.abs(y12)) + _sc(atol, y11), _sc(eps, y11))
    x = (y11 - y12) / tol
    return float(max(np.sqrt((x ** 2).sum() / x.size), eps))


def integrate_adaptive(solver, y0, ts, rtol, atol, dt_min, extra0=None):
    """base_solver.py:107-149 with the adaptive branch (:117-142)."""
    ts = np.asarray(ts)
    step_size = solver.dt
    prev_t = curr_t = ts[0]
    prev_y = curr_y = y0
    curr_extra = solver.init_extra(ts[0], y0) if extra0 is None else extra0
    ys = [y0]
    prev_error_ratio = None
    n_proposals = 0
    for out_t in ts[1:]:
        while curr_t < out_t:
            next_t = min(curr_t + step_size, ts[-1])
            next_y_full, _ = solver.step(curr_t, next_t, curr_y, curr_extra)
            midpoint_t = 0.5 * (curr_t + next_t)
            midpoint_y, midpoint_extra = solver.step(curr_t, midpoint_t, curr_y, curr_extra)
            next_y, next_extra = solver.step(midpoint_t, next_t, midpoint_y, midpoint_extra)
            error_estimate = compute_error(next_y_full, next_y, rtol, atol)
            step_size, prev_error_ratio = update_step_size(error_estimate, step_size,
                                                           prev_error_ratio=prev_error_ratio)
            n_proposals += 1
            if step_size < dt_min:
                step_size = dt_min
                prev_error_ratio = None
            if error_estimate <= 1 or step_size <= dt_min:
                prev_t, prev_y = curr_t, curr_y
                curr_t, curr_y, curr_extra = next_t, next_y, next_extra
        ys.append(linear_interp(prev_t, prev_y, curr_t, curr_y, out_t))
    return np.stack(ys, axis=0), curr_extra, n_proposals


def linear_interp(t0, y0, t1, y1, t):
    # interp.py:15-18
    w0 = (t1 - t) / (t1 - t0)
    w1 = (t - t0) / (t1 - t0)
    return _sc(w0, y0) * y0 + _sc(w1, y0) * y1


class Euler(Solver):
    # methods/euler.py:29-37
    def step(self, t0, t1, y0, extra0):
        dt = t1 - t0
        I_k = self.bm(t0, t1)
        f, g = self.sde.f(t0, y0), self.sde.g(t0, y0)
        return y0 + f * _sc(dt, y0) + _prod(g, I_k), ()


class Milstein(Solver):
    # methods/milstein.py:52-94
    def __init__(self, sde, bm, dt, options=None, ito=True):
        super().__init__(sde, bm, dt, options)
        self.ito = ito

    def step(self, t0, t1, y0, extra0):
        dt = t1 - t0
        dts = _sc(dt, y0)
        I_k = self.bm(t0, t1)
        v = I_k ** 2 - dts if self.ito else I_k ** 2
        if self.options.get('grad_free', False) and self.sde.noise_type != 'additive':
            f, g = self.sde.f(t0, y0), self.sde.g(t0, y0)
            g_ = g[:, :, 0] if g.ndim == 3 else g
            sqrt_dt = np.sqrt(dt)
            fac = dts * f if self.ito else _sc(0., y0)
            y0_prime = y0 + fac + g_ * _sc(sqrt_dt, y0)
            g_prime = self.sde.g(t0, y0_prime)
            g_prod_I_k = _prod(g, I_k)
            gdg_prod = _prod(g_prime - g, v) / _sc(2 * sqrt_dt, y0)
        else:
            f = self.sde.f(t0, y0)
            g = self.sde.g(t0, y0)
            g_prod_I_k = _prod(g, I_k)
            if self.sde.noise_type == 'additive':
                gdg_prod = _sc(0., y0)
            else:
                gdg_prod = self.sde.gdg(t0, y0, _sc(0.5, y0) * v)
        return y0 + f * dts + g_prod_I_k + gdg_prod, ()


class Heun(Solver):
    # methods/heun.py:35-48
    def step(self, t0, t1, y0, extra0):
        dt = _sc(t1 - t0, y0)
        I_k = self.bm(t0, t1)
        f, g_prod = self.sde.f(t0, y0), _prod(self.sde.g(t0, y0), I_k)
        y0_prime = y0 + dt * f + g_prod
        f_prime, g_prod_prime = self.sde.f(t1, y0_prime), _prod(self.sde.g(t1, y0_prime), I_k)
        return y0 + (dt * (f + f_prime) + g_prod + g_prod_prime) * _sc(0.5, y0), ()


class Midpoint(Solver):
    # methods/midpoint.py:29-45
    def step(self, t0, t1, y0, extra0):
        dt = t1 - t0
        I_k = self.bm(t0, t1)
        f, g_prod = self.sde.f(t0, y0), _prod(self.sde.g(t0, y0), I_k)
        half_dt = 0.5 * dt
        half_dt = np.asarray(dt).dtype.type(half_dt)
        t_prime = t0 + half_dt
        y_prime = y0 + _sc(half_dt, y0) * f + _sc(0.5, y0) * g_prod
        f_prime, g_prod_prime = self.sde.f(t_prime, y_prime), _prod(self.sde.g(t_prime, y_prime), I_k)
        return y0 + _sc(dt, y0) * f_prime + g_prod_prime, ()


class EulerHeun(Solver):
    # methods/euler_heun.py:29-42
    def step(self, t0, t1, y0, extra0):
        dt = _sc(t1 - t0, y0)
        I_k = self.bm(t0, t1)
        f, g_prod = self.sde.f(t0, y0), _prod(self.sde.g(t0, y0), I_k)
        y_prime = y0 + g_prod
        g_prod_prime = _prod(self.sde.g(t1, y_prime), I_k)
        return y0 + dt * f + (g_prod + g_prod_prime) * _sc(0.5, y0), ()


class ReversibleHeun(Solver):
    # methods/reversible_heun.py:58-73
    def init_extra(self, t0, y0):
        return self.sde.f(t0, y0), self.sde.g(t0, y0), y0

    def step(self, t0, t1, y0, extra0):
        f0, g0, z0 = extra0
        dt = t1 - t0
        dW = self.bm(t0, t1)
        z1 = _sc(2, y0) * y0 - z0 + f0 * _sc(dt, y0) + _prod(g0, dW)
        f1, g1 = self.sde.f(t1, z1), self.sde.g(t1, z1)
        half_dt = np.asarray(dt).dtype.type(0.5 * dt)
        y1 = y0 + (f0 + f1) * _sc(half_dt, y0) + _prod(g0 + g1, _sc(0.5, y0) * dW)
        return y1, (f1, g1, z1)


# methods/tableaus/srid2.py:19-54
class srid2:
    STAGES = 4
    C0 = (0, 1, 1 / 2, 0)
    C1 = (0, 1 / 4, 1, 1 / 4)
    A0 = ((), (1,), (1 / 4, 1 / 4), (0, 0, 0))
    A1 = ((), (1 / 4,), (1, 0), (0, 0, 1 / 4))
    B0 = ((), (0,), (1, 1 / 2), (0, 0, 0))
    B1 = ((), (-1 / 2,), (1, 0), (2, -1, 1 / 2))
    alpha = (1 / 6, 1 / 6, 2 / 3, 0)
    beta1 = (-1, 4 / 3, 2 / 3, 0)
    beta2 = (1, -4 / 3, 1 / 3, 0)
    beta3 = (2, -4 / 3, -2 / 3, 0)
    beta4 = (-2, 5 / 3, -2 / 3, 1)


# methods/tableaus/sra1.py:19-36
class sra1:
    STAGES = 2
    C0 = (0, 3 / 4)
    C1 = (1, 0)
    A0 = ((), (3 / 4,))
    B0 = ((), (3 / 2,))
    alpha = (1 / 3, 2 / 3)
    beta1 = (1, 0)
    beta2 = (-1, 1)


class SRK(Solver):
    """methods/srk.py:57-111 — the reference's loops verbatim in structure (including its
    redundant re-evaluations), so that the CUDA path's de-duplicated stages are checked against the
    original formulation."""
    want_u = True

    def step(self, t0, t1, y0, extra0):
        if self.sde.noise_type == 'additive':
            return self.additive_step(t0, t1, y0)
        return self.diagonal_or_scalar_step(t0, t1, y0)

    def diagonal_or_scalar_step(self, t0, t1, y0):
        tt = np.asarray(t0).dtype.type
        dt = t1 - t0
        rdt = tt(1 / dt)
        sqrt_dt = np.sqrt(dt)
        I_k, I_k0 = self.bm(t0, t1, return_U=True)
        s = lambda x: _sc(x, y0)  # noqa
        I_kk = (I_k ** 2 - s(dt)) * s(1 / 2)
        I_kkk = (I_k * I_k * I_k - s(tt(3 * dt)) * I_k) * s(1 / 6)
        y1 = y0
        H0, H1 = [], []
        for st in range(srid2.STAGES):
            H0s, H1s = y0, y0
            for j in range(st):
                f = self.sde.f(t0 + tt(srid2.C0[j] * dt), H0[j])
                g = self.sde.g(t0 + tt(srid2.C1[j] * dt), H1[j])
                g = g[:, :, 0] if g.ndim == 3 else g
                H0s = H0s + s(srid2.A0[st][j]) * f * s(dt) + s(srid2.B0[st][j]) * g * I_k0 * s(rdt)
                H1s = H1s + s(srid2.A1[st][j]) * f * s(dt) + s(srid2.B1[st][j]) * g * s(sqrt_dt)
            H0.append(H0s)
            H1.append(H1s)
            f = self.sde.f(t0 + tt(srid2.C0[st] * dt), H0s)
            g_weight = (s(srid2.beta1[st]) * I_k + s(srid2.beta2[st]) * I_kk / s(sqrt_dt)
                        + s(srid2.beta3[st]) * I_k0 * s(rdt) + s(srid2.beta4[st]) * I_kkk * s(rdt))
            g_prod = _prod(self.sde.g(t0 + tt(srid2.C1[st] * dt), H1s), g_weight)
            y1 = y1 + s(srid2.alpha[st]) * f * s(dt) + g_prod
        return y1, ()

    def additive_step(self, t0, t1, y0):
        tt = np.asarray(t0).dtype.type
        dt = t1 - t0
        rdt = tt(1 / dt)
        I_k, I_k0 = self.bm(t0, t1, return_U=True)
        s = lambda x: _sc(x, y0)  # noqa
        y1 = y0
        H0 = []
        for i in range(sra1.STAGES):
            H0i = y0
            for j in range(i):
                f = self.sde.f(t0 + tt(sra1.C0[j] * dt), H0[j])
                g_weight = s(sra1.B0[i][j]) * I_k0 * s(rdt)
                g_prod = _prod(self.sde.g(t0 + tt(sra1.C1[j] * dt), y0), g_weight)
                H0i = H0i + s(sra1.A0[i][j]) * f * s(dt) + g_prod
            H0.append(H0i)
            f = self.sde.f(t0 + tt(sra1.C0[i] * dt), H0i)
            g_weight = s(sra1.beta1[i]) * I_k + s(sra1.beta2[i]) * I_k0 * s(rdt)
            g_prod = _prod(self.sde.g(t0 + tt(sra1.C1[i] * dt), y0), g_weight)
            y1 = y1 + s(sra1.alpha[i]) * f * s(dt) + g_prod
        return y1, ()


def adjoint_reversible_heun_step(sde, t0, t1, y0, z0, f0, g0, adj_y0, adj_f0, adj_g0, adj_z0, dW, vjp):
    """One backward step, methods/reversible_heun.py:98-144, on separate arrays.
    `vjp(t, z, adj_f, adj_g) -> vjp_z` supplies the autograd vjp of (f, g) at (t, z) (:119-129).
    t0 < t1 are the *reversed* times; the forward SDE is evaluated at -t0, -t1."""
    s = lambda x: _sc(x, y0)  # noqa
    dt = t1 - t0
    half_dt = np.asarray(dt).dtype.type(0.5 * dt)
    half_dW = s(0.5) * dW
    aop = (lambda a, b: a * b) if g0.ndim == 2 else (lambda a, b: a[:, :, None] * b[:, None, :])
    adj_y0_half_dt = adj_y0 * s(half_dt)
    adj_y0_half_dW = aop(adj_y0, half_dW)
    z1 = s(2) * y0 - z0 - f0 * s(dt) - _prod(g0, dW)
    adj_f1 = adj_y0_half_dt
    adj_f0 = adj_f0 + adj_y0_half_dt
    adj_g1 = adj_y0_half_dW
    adj_g0 = adj_g0 + adj_y0_half_dW
    vjp_z = vjp(-t0, z0, adj_f0, adj_g0)
    adj_z0 = adj_z0 + vjp_z
    f1, g1 = sde.f(-t1, z1), sde.g(-t1, z1)
    y1 = y0 - (f0 + f1) * s(half_dt) - _prod(g0 + g1, half_dW)
    adj_y1 = adj_y0 + s(2) * adj_z0
    adj_z1 = -adj_z0
    adj_f1 = adj_f1 + adj_z0 * s(dt)
    adj_g1 = adj_g1 + aop(adj_z0, dW)
    return (y1, z1, f1, g1), (adj_y1, adj_f1, adj_g1, adj_z1)


def make(method, sde, bm, dt, options=None):
    if method == 'euler':
        return Euler(sde, bm, dt, options)
    if method == 'milstein':
        return Milstein(sde, bm, dt, options, ito=sde.sde_type == 'ito')
    if method == 'srk':
        return SRK(sde, bm, dt, options)
    if method == 'heun':
        return Heun(sde, bm, dt, options)
    if method == 'midpoint':
        return Midpoint(sde, bm, dt, options)
    if method == 'euler_heun':
        return EulerHeun(sde, bm, dt, options)
    if method == 'reversible_heun':
        return ReversibleHeun(sde, bm, dt, options)
    raise ValueError(method)
