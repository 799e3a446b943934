"""ReverseBrownian / BrownianPath / BrownianTree / brownian_interval_like.

Thin views over the CUDA-backed `BrownianInterval` with the reference's constructor signatures and
semantics (torchsde/_brownian/derived.py: ReverseBrownian :22-50, BrownianPath :52-103, BrownianTree
:106-191, brownian_interval_like :194-205).
"""
from . import brownian_base
from . import interval as brownian_interval


class _View(brownian_base.BaseBrownian):
    """A Brownian motion defined in terms of another one (`self._base`): attributes are forwarded."""

    def __init__(self, base):
        super().__init__()
        self._base = base

    dtype = property(lambda self: self._base.dtype)
    device = property(lambda self: self._base.device)
    shape = property(lambda self: self._base.shape)
    levy_area_approximation = property(lambda self: self._base.levy_area_approximation)


class ReverseBrownian(_View):
    """Time reversal: (ta, tb) -> base(-tb, -ta).  The statistics are not negated because the adjoint
    SDE already negates drift and diffusion (derived.py:27-30)."""

    def __init__(self, base_brownian):
        super().__init__(base_brownian)
        self.base_brownian = base_brownian

    def __call__(self, ta, tb=None, return_U=False, return_A=False):
        return self.base_brownian(-tb, -ta, return_U=return_U, return_A=return_A)

    def __repr__(self):
        return f"{self.__class__.__name__}(base_brownian={self.base_brownian})"


class _Anchored(_View):
    """Interval plus an initial value w0: a point query `bm(t)` returns w0 + W(t0, t) (derived.py:79-84)."""

    def __init__(self, w0, interval):
        super().__init__(interval)
        self._w0 = w0
        self._interval = interval

    def __call__(self, t, tb=None, return_U=False, return_A=False):
        # first argument deliberately called t (derived.py:80)
        out = self._interval(t, tb, return_U=return_U, return_A=return_A)
        if tb is None and not return_U and not return_A:
            out = out + self._w0
        return out

    def __repr__(self):
        return f"{self.__class__.__name__}(interval={self._interval})"


class BrownianPath(_Anchored):
    """Brownian path that remembers every computed value: an interval over [t0, t0 + 1] with an unbounded
    node cache (derived.py:66-77).  `window_size` is accepted and unused, as in the reference."""

    def __init__(self, t0, w0, window_size=8):
        del window_size
        super().__init__(w0, brownian_interval.BrownianInterval(
            t0=t0, t1=t0 + 1, size=w0.shape, dtype=w0.dtype, device=w0.device, cache_size=None))


class BrownianTree(_Anchored):
    """Brownian motion whose sample path depends on `entropy` only, not on where or in which order it is
    queried: the dyadic (`halfway_tree`) interval resolved to `tol` (derived.py:122-168).  `cache_depth`
    and `safety` are accepted and unused, as in the reference."""

    def __init__(self, t0, w0, t1=None, w1=None, entropy=None, tol=1e-6, pool_size=24, cache_depth=9,
                 safety=None):
        del cache_depth, safety
        super().__init__(w0, brownian_interval.BrownianInterval(
            t0=t0, t1=t0 + 1 if t1 is None else t1, size=w0.shape, dtype=w0.dtype, device=w0.device,
            entropy=entropy, tol=tol, pool_size=pool_size, halfway_tree=True,
            W=None if w1 is None else w1 - w0))


def brownian_interval_like(y, t0=0., t1=1., size=None, dtype=None, device=None, **kwargs):
    """BrownianInterval with the size, dtype and device of `y` unless overridden (derived.py:194-205)."""
    return brownian_interval.BrownianInterval(
        t0=t0, t1=t1, size=y.shape if size is None else size, dtype=y.dtype if dtype is None else dtype,
        device=y.device if device is None else device, **kwargs)
