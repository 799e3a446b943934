"""Autograd support for plain `sdeint` (backpropagation through the solver).

The reference's solvers are sequences of differentiable ATen ops, so gradients flow through `sdeint`
(DOCUMENTATION.md, tests/test_adjoint.py:100-154 compare the adjoint against exactly that).  Here every
tableau launch is one `torch.autograd.Function`: forward = the fused CUDA kernel (C ABI), backward = the
transpose of the tableau.  Every tableau is *linear* in its tensor inputs once the Brownian increment is
fixed, out_o = sum_i c_oi(W, U) . in_i, so the backward of input i is sum_o c_oi . grad_o with
  - 'e' coefficients: a scalar or a (rows, d)-broadcastable tensor (element-wise), or
  - 'g' coefficients: a weight tensor of the Brownian shape — the adjoint of the product g.v
    (element-wise for diagonal / user-supplied products, outer product grad (x) v for (rows, d, m) g).
The table below lists c_oi for every entry point of include/torchsde_b200.h that a `.step` uses.
The backward arithmetic is a handful of torch element-wise ops on the materialised increments (what the
reference's own backward does); the hot no-grad / adjoint paths never come here.
"""
import torch

_S13, _S23, _S16 = 1.0 / 3, 2.0 / 3, 1.0 / 6


def _srk_weights(w, u, dt, rdt, sqrt_dt, three_dt):
    """gw_s of methods/srk.py:80-85 with tableaus/srid2.py:50-54."""
    ikk = (w * w - dt) * 0.5
    ikkk = (w * w * w - three_dt * w) * _S16
    b1 = (-1, 4 / 3, 2 / 3, 0)
    b2 = (1, -4 / 3, 1 / 3, 0)
    b3 = (2, -4 / 3, -2 / 3, 0)
    b4 = (-2, 5 / 3, -2 / 3, 1)
    return [b1[s] * w + b2[s] * ikk / sqrt_dt + b3[s] * u * rdt + b4[s] * ikkk * rdt for s in range(4)]


# name -> function(scalars, w, u) -> list over outputs of list over inputs of (kind, coeff)
def _table(name, sc, w, u):
    E, G = 'e', 'g'
    if name == 'tsde_step_euler':
        dt, = sc
        return [[(E, 1.0), (E, dt), (G, w)]]
    if name == 'tsde_milstein_vjp_seed':
        dt, ito = sc
        v = w * w - dt if ito else w * w
        return [[('gg', 0.5 * v)]]
    if name == 'tsde_step_milstein':
        dt, = sc
        return [[(E, 1.0), (E, dt), (G, w), (E, 1.0)]]
    if name == 'tsde_milstein_gf_predict':
        dt, sqrt_dt, ito = sc
        return [[(E, 1.0), (E, dt if ito else 0.0), (E, sqrt_dt)]]
    if name == 'tsde_step_milstein_gf':
        dt, two_sqrt_dt, ito = sc
        v = (w * w - dt if ito else w * w) / two_sqrt_dt
        return [[(E, 1.0), (E, dt), (G, w - v), (G, v)]]
    if name == 'tsde_step_heun':
        dt, = sc
        return [[(E, 1.0), (E, 0.5 * dt), (E, 0.5 * dt), (G, 0.5 * w), (G, 0.5 * w)]]
    if name == 'tsde_midpoint_predict':
        half_dt, = sc
        return [[(E, 1.0), (E, half_dt), (G, 0.5 * w)]]
    if name == 'tsde_euler_heun_predict':
        return [[(E, 1.0), (G, w)]]
    if name == 'tsde_step_euler_heun':
        dt, = sc
        return [[(E, 1.0), (E, dt), (G, 0.5 * w), (G, 0.5 * w)]]
    if name == 'tsde_reversible_heun_z':
        dt, = sc
        return [[(E, 2.0), (E, -1.0), (E, dt), (G, w)]]
    if name == 'tsde_step_reversible_heun':
        half_dt, = sc
        return [[(E, 1.0), (E, half_dt), (E, half_dt), (G, 0.5 * w), (G, 0.5 * w)]]
    if name == 'tsde_srk_diag_stage1':
        dt, sqrt_dt = sc
        return [[(E, 1.0), (E, dt), (E, 0.0)], [(E, 1.0), (E, 0.25 * dt), (E, -0.5 * sqrt_dt)]]
    if name == 'tsde_srk_diag_stage2':
        dt, rdt, sqrt_dt = sc
        return [[(E, 1.0), (E, 0.25 * dt), (E, u * rdt), (E, 0.25 * dt), (E, 0.5 * u * rdt)],
                [(E, 1.0), (E, dt), (E, sqrt_dt), (E, 0.0), (E, 0.0)]]
    if name == 'tsde_srk_diag_stage3':
        dt, sqrt_dt = sc
        return [[(E, 1.0), (E, 2.0 * sqrt_dt), (E, -sqrt_dt), (E, 0.25 * dt), (E, 0.5 * sqrt_dt)]]
    if name == 'tsde_step_srk_diag':
        dt, rdt, sqrt_dt, three_dt = sc
        gw = _srk_weights(w, u, dt, rdt, sqrt_dt, three_dt)
        return [[(E, 1.0), (E, _S16 * dt), (E, _S16 * dt), (E, _S23 * dt),
                 (E, gw[0]), (E, gw[1]), (E, gw[2]), (E, gw[3])]]
    if name == 'tsde_srk_additive_stage':
        dt, rdt = sc
        return [[(E, 1.0), (E, 0.75 * dt), (G, 1.5 * u * rdt)]]
    if name == 'tsde_step_srk_additive':
        dt, rdt = sc
        return [[(E, 1.0), (E, _S13 * dt), (E, _S23 * dt), (G, w - u * rdt), (G, u * rdt)]]
    if name == 'tsde_linear_interp':
        w0, w1 = sc
        return [[(E, w0), (E, w1)]]
    raise NotImplementedError(f"torchsde_b200: no autograd rule for {name}")


class TableauFn(torch.autograd.Function):
    """One fused tableau launch as an autograd node."""

    @staticmethod
    def forward(ctx, solver, name, use_general, unit, noise, scalars, n_out, *inputs):
        ctx.name, ctx.scalars, ctx.noise, ctx.unit, ctx.use_general = name, scalars, noise, unit, use_general
        ctx.in_shapes = [tuple(t.shape) for t in inputs]
        outs = solver._launch_raw(name, use_general, unit, noise, [t.detach() for t in inputs], scalars, n_out)
        return outs[0] if n_out == 1 else tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        w, u = ctx.noise if ctx.noise is not None else (None, None)
        if ctx.unit:
            w = 1.0  # the product was supplied by the user: the kernel saw W == 1
        table = _table(ctx.name, ctx.scalars, w, u)
        n_in = len(ctx.in_shapes)
        out = [None] * n_in
        for o, g_o in enumerate(grads):
            if g_o is None:
                continue
            for i, (kind, c) in enumerate(table[o]):
                shape = ctx.in_shapes[i]
                if kind == 'e' or ctx.unit:
                    term = g_o * c  # (scalar noise: a (rows, 1) coefficient broadcasts over d)
                elif kind == 'gg':  # g-shaped input AND output (Milstein's vjp seed)
                    term = g_o * (c.unsqueeze(-2) if g_o.dim() == 3 else c)
                else:  # 'g': adjoint of the product g.v
                    if len(shape) == 3:
                        term = g_o.unsqueeze(-1) * c.unsqueeze(-2)
                    else:
                        term = g_o * c
                out[i] = term if out[i] is None else out[i] + term
        out = [None if t is None else t.reshape(s) for t, s in zip(out, ctx.in_shapes)]
        return (None, None, None, None, None, None, None, *out)
