"""Abstract Brownian-motion interface.

The solver only relies on this protocol (it is duck-typed, reference torchsde/_core/base_solver.py:54-57):
a callable ``bm(ta, tb=None, return_U=False, return_A=False)`` plus the read-only attributes ``dtype``,
``device``, ``shape`` and ``levy_area_approximation``; ``size()`` is an alias of ``shape``
(reference: torchsde/_brownian/brownian_base.py:18-50).
"""
import abc


def _required(name):
    def getter(self):
        raise NotImplementedError(f"{type(self).__name__} must define `{name}`")
    getter.__name__ = name
    return property(abc.abstractmethod(getter))


class BaseBrownian(abc.ABC):
    __slots__ = ()

    dtype = _required('dtype')
    device = _required('device')
    shape = _required('shape')
    levy_area_approximation = _required('levy_area_approximation')

    @abc.abstractmethod
    def __call__(self, ta, tb=None, return_U=False, return_A=False):
        """Increment over [ta, tb] (or the value at `ta` when `tb` is None), optionally with the space-time
        Levy area U and the Levy-area approximation A: W | (W, U) | (W, A) | (W, U, A)."""

    @abc.abstractmethod
    def __repr__(self):
        ...

    def size(self):
        return self.shape
