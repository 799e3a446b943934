"""The augmented backward SDE of the continuous adjoint method.

Same mathematics as the reference's ``AdjointSDE`` (torchsde/_core/adjoint_sde.py:23-377): the state is
the flat vector ``(y, adj_y, adj_params...)`` with a dummy batch dimension, the drift / diffusion-vector
products are ``-f``, ``-g.v`` of the forward SDE evaluated at time ``-t`` together with their vector-
Jacobian products w.r.t. ``y`` and the parameters, with the Ito <-> Stratonovich corrections for Ito SDEs
(:111-230) and Milstein's extra term for diagonal noise (:332-377).

This is autograd glue around the *user's* f and g (it has to stay in PyTorch: the vjps are taken through
the user's own op graph); the solver arithmetic on the flat state runs in the fused tableau kernels
(element-wise launches with the products supplied, TSDE_SRC_UNIT).  The noise-type map forward ->
adjoint is :33-38.
"""
import torch

from . import base_sde
from ..settings import NOISE_TYPES, SDE_TYPES


def _flat(tensors):
    return torch.cat([x.reshape(-1) for x in tensors]).unsqueeze(0)


def _vjp(outputs, inputs, grad_outputs=None, create_graph=False, retain_graph=True):
    """Vector-Jacobian product with zeros for unreachable inputs (misc.py:71-81)."""
    single = torch.is_tensor(outputs)
    outputs = [outputs] if single else list(outputs)
    if grad_outputs is not None and torch.is_tensor(grad_outputs):
        grad_outputs = [grad_outputs]
    keep = [i for i, o in enumerate(outputs) if o.requires_grad]
    if not keep:
        return [torch.zeros_like(x) for x in inputs]
    outs = [outputs[i] for i in keep]
    gos = None if grad_outputs is None else [grad_outputs[i] for i in keep]
    grads = torch.autograd.grad(outs, inputs, gos, allow_unused=True, retain_graph=retain_graph,
                                create_graph=create_graph)
    return [torch.zeros_like(x) if g is None else g for g, x in zip(grads, inputs)]


def _jvp(output, inp, tangent, create_graph=False):
    """Jacobian-vector product by the double-vjp trick (misc.py:84-99)."""
    if not output.requires_grad:
        return torch.zeros_like(output)
    dummy = torch.zeros_like(output, requires_grad=True)
    vjp, = torch.autograd.grad(output, inp, dummy, create_graph=True, allow_unused=True)
    if vjp is None or not vjp.requires_grad:
        return torch.zeros_like(output)
    out, = torch.autograd.grad(vjp, dummy, tangent, allow_unused=True, create_graph=create_graph,
                               retain_graph=True)
    return torch.zeros_like(output) if out is None else out


class AdjointSDE(base_sde.BaseSDE):
    is_adjoint_sde = True
    # how the solver must obtain products from this SDE (see methods._ProdMixin)
    f_and_g_prod_mode = 'f_and_g_prod'
    g_prod_mode = 'g_prod'
    user_g_prod = True
    user_f_and_g_prod = True

    def __init__(self, forward_sde, params, shapes):
        noise_type = {NOISE_TYPES.general: NOISE_TYPES.general, NOISE_TYPES.additive: NOISE_TYPES.general,
                      NOISE_TYPES.scalar: NOISE_TYPES.scalar,
                      NOISE_TYPES.diagonal: NOISE_TYPES.diagonal}[forward_sde.noise_type]
        super(AdjointSDE, self).__init__(sde_type=forward_sde.sde_type, noise_type=noise_type)
        self.forward_sde = forward_sde
        self.params = list(params)
        self._shapes = list(shapes)
        self._base_sde = forward_sde._base_sde
        fwd_noise = forward_sde.noise_type
        ito = forward_sde.sde_type == SDE_TYPES.ito
        # which correction applies (adjoint_sde.py:51-72)
        if not ito or fwd_noise == NOISE_TYPES.additive:
            self._correction = None
        elif fwd_noise == NOISE_TYPES.diagonal:
            self._correction = 'diagonal'
        else:
            self._correction = 'default'

    # ---- forward SDE through differentiable torch ops -------------------------------------------
    def _prod(self, g, v):
        # base_sde.py:98-102
        if self.forward_sde.noise_type == NOISE_TYPES.diagonal:
            return g * v
        return torch.bmm(g, v.unsqueeze(-1)).squeeze(-1)

    def _fwd_g_prod(self, t, y, v):
        fwd = self.forward_sde
        if fwd.user_g_prod:
            return fwd.g_prod(t, y, v)
        return self._prod(fwd.g(t, y), v)

    def _fwd_f_and_g_prod(self, t, y, v):
        fwd = self.forward_sde
        mode = fwd.f_and_g_prod_mode
        if mode == 'f_and_g_prod':
            return fwd.f_and_g_prod(t, y, v)
        if mode == 'g_prod':
            return fwd.f(t, y), fwd.g_prod(t, y, v)
        f, g = fwd.f_and_g(t, y)
        return f, self._prod(g, v)

    # ---- state handling (adjoint_sde.py:74-109) ---------------------------------------------------
    def get_state(self, t, y_aug, v=None, extra_states=False):
        shapes = self._shapes if extra_states else self._shapes[:2]
        numels = [int(torch.Size(s).numel()) for s in shapes]
        parts = y_aug.squeeze(0)[:sum(numels)].split(numels)
        y, adj_y, *extras = [p.reshape(s) for p, s in zip(parts, shapes)]
        if not y.requires_grad:
            y = y.detach().requires_grad_()
        return y, adj_y, extras, torch.is_grad_enabled()

    # ---- drift ---------------------------------------------------------------------------------------
    def _drift_out(self, f, g, y, adj_y, create_graph):
        inputs = [y] + self.params
        if self._correction == 'diagonal':          # :169-207
            g_dg, = _vjp(g, [y], g, create_graph=True)
            f = f - g_dg
            vjps = _vjp(f, inputs, adj_y, create_graph=create_graph)
            a_dg, = _vjp(g, [y], adj_y, create_graph=create_graph)
            extra = _vjp(g, inputs, a_dg, create_graph=create_graph)
            vjps = [a + b for a, b in zip(vjps, extra)]
        elif self._correction == 'default':         # :127-167
            cols = [c.squeeze(-1) for c in g.split(1, dim=-1)]
            f = f - sum(_jvp(c, y, c, create_graph=True) for c in cols)
            vjps = _vjp(f, inputs, adj_y, create_graph=create_graph)
            for c in cols:
                a_dg, = _vjp(c, [y], adj_y, create_graph=create_graph)
                extra = _vjp(c, inputs, a_dg, create_graph=create_graph)
                vjps = [a + b for a, b in zip(vjps, extra)]
        else:                                        # :111-125
            vjps = _vjp(f, inputs, adj_y, create_graph=create_graph)
        if not create_graph:
            f = f.detach()
        return _flat([-f] + list(vjps))

    def _diffusion_out(self, g_prod, y, adj_y, create_graph):
        vjps = _vjp(g_prod, [y] + self.params, adj_y, create_graph=create_graph)   # :209-222
        if not create_graph:
            g_prod = g_prod.detach()
        return _flat([-g_prod] + list(vjps))

    def f(self, t, y_aug):
        y, adj_y, _, rg = self.get_state(t, y_aug)
        with torch.enable_grad():
            if self._correction is None:
                f, g = self.forward_sde.f(-t, y), None
            else:
                f, g = self.forward_sde.f_and_g(-t, y)
            return self._drift_out(f, g, y, adj_y, rg)

    def g(self, t, y):
        raise RuntimeError("Adjoint `g` not defined. Please report a bug to torchsde_b200.")

    def f_and_g(self, t, y):
        raise RuntimeError("Adjoint `f_and_g` not defined. Please report a bug to torchsde_b200.")

    def g_prod(self, t, y_aug, v):
        y, adj_y, _, rg = self.get_state(t, y_aug, v)
        with torch.enable_grad():
            return self._diffusion_out(self._fwd_g_prod(-t, y, v), y, adj_y, rg)

    def f_and_g_prod(self, t, y_aug, v):
        y, adj_y, _, rg = self.get_state(t, y_aug)
        with torch.enable_grad():
            if self._correction is None:            # :296-303
                f, g_prod = self._fwd_f_and_g_prod(-t, y, v)
                g = None
            else:                                    # :305-323
                f, g = self.forward_sde.f_and_g(-t, y)
                g_prod = self._prod(g, v)
            return self._drift_out(f, g, y, adj_y, rg), self._diffusion_out(g_prod, y, adj_y, rg)

    def g_prod_and_gdg_prod(self, t, y_aug, v1, v2):
        """Milstein's products for diagonal noise (:332-377)."""
        if self.forward_sde.noise_type != NOISE_TYPES.diagonal:
            raise NotImplementedError
        y, adj_y, _, rg = self.get_state(t, y_aug, v2)
        inputs = [y] + self.params
        with torch.enable_grad():
            g = self.forward_sde.g(-t, y)
            g_prod = self._prod(g, v1)
            vg_dg, = _vjp(g, [y], v2 * g, create_graph=rg)
            dgdy, = _vjp(g.sum(), [y], None, create_graph=rg)
            prod_partials = _vjp(g, inputs, adj_y * v2 * dgdy, create_graph=rg)
            avg_dg, = _vjp(g, [y], (adj_y * v2 * g).detach(), create_graph=True)
            mixed_partials = _vjp(avg_dg.sum(), inputs, None, create_graph=rg)
            vjps = [a - b for a, b in zip(prod_partials, mixed_partials)]
            return self._diffusion_out(g_prod, y, adj_y, rg), _flat([vg_dg] + list(vjps))
