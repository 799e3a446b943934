"""SDE base classes and the one-time function table the solver calls up into.

API mirror of the reference's ``torchsde/_core/base_sde.py``: ``BaseSDE`` (:25-39),
``ForwardSDE`` (:42-73), ``RenameMethodsSDE`` (:212-224), ``SDEIto`` / ``SDEStratonovich``
(:227-236), ``SDELogqp`` (:240-306).

Difference to the reference: ``ForwardSDE`` does not implement ``prod`` / ``g_prod_default`` /
``f_and_g_prod_default*`` with ATen ops.  Those products (``g*v``, ``bmm(g, v)``;
base_sde.py:98-120) are fused into the CUDA tableau kernels; the table only records *which* of
the user's callables exist so that the solver reproduces the reference's call pattern
(base_sde.py:51-61): a user-supplied ``f_and_g_prod`` / ``g_prod`` is honoured (the increment is
then materialised for it), otherwise ``f_and_g`` / ``f``,``g`` are evaluated and the product
happens inside the kernel.
"""
import abc

import torch
from torch import nn

from ..settings import NOISE_TYPES, SDE_TYPES


class BaseSDE(abc.ABC, nn.Module):
    """Base class for all SDEs; validates `noise_type` and `sde_type` (base_sde.py:25-39)."""

    def __init__(self, noise_type, sde_type):
        super(BaseSDE, self).__init__()
        if noise_type not in NOISE_TYPES:
            raise ValueError(f"Expected noise type in {NOISE_TYPES}, but found {noise_type}")
        if sde_type not in SDE_TYPES:
            raise ValueError(f"Expected sde type in {SDE_TYPES}, but found {sde_type}")
        self.noise_type = noise_type
        self.sde_type = sde_type


class ForwardSDE(BaseSDE):

    def __init__(self, sde):
        super(ForwardSDE, self).__init__(sde_type=sde.sde_type, noise_type=sde.noise_type)
        self._base_sde = sde
        # Which specialised callables did the user provide? (base_sde.py:51-61)
        self.user_f_and_g_prod = hasattr(sde, 'f_and_g_prod')
        self.user_g_prod = hasattr(sde, 'g_prod')
        self.user_f_and_g = hasattr(sde, 'f_and_g')
        self.f = getattr(sde, 'f', self.f_default)
        self.g = getattr(sde, 'g', self.g_default)
        self.f_and_g = getattr(sde, 'f_and_g', self.f_and_g_default)
        if self.user_g_prod:
            self.g_prod = sde.g_prod
        if self.user_f_and_g_prod:
            self.f_and_g_prod = sde.f_and_g_prod

    def f_default(self, t, y):
        raise RuntimeError("Method `f` has not been provided, but is required for this method.")

    def g_default(self, t, y):
        raise RuntimeError("Method `g` has not been provided, but is required for this method.")

    def f_and_g_default(self, t, y):
        return self.f(t, y), self.g(t, y)

    # How the solver obtains (f, g.v) for a `f_and_g_prod`-style call site (base_sde.py:51-56):
    #   'f_and_g_prod' : user's f_and_g_prod(t, y, v)
    #   'g_prod'       : user's f(t, y) and g_prod(t, y, v)          (f_and_g_prod_default1 :115-116)
    #   'fused'        : f_and_g(t, y), product inside the kernel    (f_and_g_prod_default2 :118-120)
    @property
    def f_and_g_prod_mode(self):
        if self.user_f_and_g_prod:
            return 'f_and_g_prod'
        if hasattr(self._base_sde, 'f') and self.user_g_prod:
            return 'g_prod'
        return 'fused'

    # How the solver obtains g.v for a `g_prod` call site (base_sde.py:54, :108-109).
    @property
    def g_prod_mode(self):
        return 'g_prod' if self.user_g_prod else 'fused'


class RenameMethodsSDE(BaseSDE):

    def __init__(self, sde, drift='f', diffusion='g', prior_drift='h', diffusion_prod='g_prod',
                 drift_and_diffusion='f_and_g', drift_and_diffusion_prod='f_and_g_prod'):
        super(RenameMethodsSDE, self).__init__(noise_type=sde.noise_type, sde_type=sde.sde_type)
        self._base_sde = sde
        for name, value in zip(('f', 'g', 'h', 'g_prod', 'f_and_g', 'f_and_g_prod'),
                               (drift, diffusion, prior_drift, diffusion_prod, drift_and_diffusion,
                                drift_and_diffusion_prod)):
            try:
                setattr(self, name, getattr(sde, value))
            except AttributeError:
                pass


class SDEIto(BaseSDE):

    def __init__(self, noise_type):
        super(SDEIto, self).__init__(noise_type=noise_type, sde_type=SDE_TYPES.ito)


class SDEStratonovich(BaseSDE):

    def __init__(self, noise_type):
        super(SDEStratonovich, self).__init__(noise_type=noise_type, sde_type=SDE_TYPES.stratonovich)


def _stable_division(a, b, epsilon=1e-7):
    # misc.py:66-68
    b = torch.where(b.abs().detach() > epsilon, b, torch.full_like(b, fill_value=epsilon) * b.sign())
    return a / b


class SDELogqp(BaseSDE):
    """Augments the state with the KL integrand (base_sde.py:240-306).  This is user-level model
    code composed of the user's own f/g/h callables (torch ops), not part of the solver kernels."""

    def __init__(self, sde):
        super(SDELogqp, self).__init__(noise_type=sde.noise_type, sde_type=sde.sde_type)
        self._base_sde = sde
        try:
            self._base_f = sde.f
            self._base_g = sde.g
            self._base_h = sde.h
        except AttributeError as e:
            raise AttributeError("If using logqp then drift, diffusion and prior drift must all be specified.") from e
        if sde.noise_type == NOISE_TYPES.diagonal:
            self.f = self.f_diagonal
            self.g = self.g_diagonal
            self.f_and_g = self.f_and_g_diagonal
        else:
            self.f = self.f_general
            self.g = self.g_general
            self.f_and_g = self.f_and_g_general

    def f_diagonal(self, t, y):
        y = y[:, :-1]
        f, g, h = self._base_f(t, y), self._base_g(t, y), self._base_h(t, y)
        u = _stable_division(f - h, g)
        f_logqp = .5 * (u ** 2).sum(dim=1, keepdim=True)
        return torch.cat([f, f_logqp], dim=1)

    def g_diagonal(self, t, y):
        y = y[:, :-1]
        g = self._base_g(t, y)
        g_logqp = y.new_zeros(size=(y.size(0), 1))
        return torch.cat([g, g_logqp], dim=1)

    def f_and_g_diagonal(self, t, y):
        y = y[:, :-1]
        f, g, h = self._base_f(t, y), self._base_g(t, y), self._base_h(t, y)
        u = _stable_division(f - h, g)
        f_logqp = .5 * (u ** 2).sum(dim=1, keepdim=True)
        g_logqp = y.new_zeros(size=(y.size(0), 1))
        return torch.cat([f, f_logqp], dim=1), torch.cat([g, g_logqp], dim=1)

    def _u_general(self, f, g, h):
        return torch.bmm(g.pinverse(), (f - h).unsqueeze(-1)).squeeze(-1)

    def f_general(self, t, y):
        y = y[:, :-1]
        f, g, h = self._base_f(t, y), self._base_g(t, y), self._base_h(t, y)
        u = self._u_general(f, g, h)
        f_logqp = .5 * (u ** 2).sum(dim=1, keepdim=True)
        return torch.cat([f, f_logqp], dim=1)

    def g_general(self, t, y):
        y = y[:, :-1]
        g = self._base_sde.g(t, y)
        g_logqp = y.new_zeros(size=(g.size(0), 1, g.size(-1)))
        return torch.cat([g, g_logqp], dim=1)

    def f_and_g_general(self, t, y):
        y = y[:, :-1]
        f, g, h = self._base_f(t, y), self._base_g(t, y), self._base_h(t, y)
        u = self._u_general(f, g, h)
        f_logqp = .5 * (u ** 2).sum(dim=1, keepdim=True)
        g_logqp = y.new_zeros(size=(g.size(0), 1, g.size(-1)))
        return torch.cat([f, f_logqp], dim=1), torch.cat([g, g_logqp], dim=1)
