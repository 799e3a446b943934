"""ReverseBrownian / BrownianPath / BrownianTree / brownian_interval_like.

Thin wrappers with the reference's signatures and semantics
(torchsde/_brownian/derived.py:22-50, 52-103, 106-191, 194-205) over the CUDA-backed
`BrownianInterval`.
"""
from . import brownian_base
from . import interval as brownian_interval


class ReverseBrownian(brownian_base.BaseBrownian):
    """(ta, tb) -> base(-tb, -ta); no sign flip (derived.py:27-30)."""

    def __init__(self, base_brownian):
        super(ReverseBrownian, self).__init__()
        self.base_brownian = base_brownian

    def __call__(self, ta, tb=None, return_U=False, return_A=False):
        return self.base_brownian(-tb, -ta, return_U=return_U, return_A=return_A)

    def __repr__(self):
        return f"{self.__class__.__name__}(base_brownian={self.base_brownian})"

    @property
    def dtype(self):
        return self.base_brownian.dtype

    @property
    def device(self):
        return self.base_brownian.device

    @property
    def shape(self):
        return self.base_brownian.shape

    @property
    def levy_area_approximation(self):
        return self.base_brownian.levy_area_approximation


class BrownianPath(brownian_base.BaseBrownian):
    """Brownian path, storing every computed value (derived.py:52-103): a BrownianInterval over
    [t0, t0 + 1] with an unbounded cache; point queries add w0."""

    def __init__(self, t0, w0, window_size=8):
        t1 = t0 + 1
        self._w0 = w0
        self._interval = brownian_interval.BrownianInterval(t0=t0, t1=t1, size=w0.shape, dtype=w0.dtype,
                                                            device=w0.device, cache_size=None)
        super(BrownianPath, self).__init__()

    def __call__(self, t, tb=None, return_U=False, return_A=False):
        out = self._interval(t, tb, return_U=return_U, return_A=return_A)
        if tb is None and not return_U and not return_A:
            out = out + self._w0
        return out

    def __repr__(self):
        return f"{self.__class__.__name__}(interval={self._interval})"

    @property
    def dtype(self):
        return self._interval.dtype

    @property
    def device(self):
        return self._interval.device

    @property
    def shape(self):
        return self._interval.shape

    @property
    def levy_area_approximation(self):
        return self._interval.levy_area_approximation


class BrownianTree(brownian_base.BaseBrownian):
    """Brownian tree with fixed entropy (derived.py:106-191): the dyadic `halfway_tree`, so the map
    entropy -> path does not depend on the locations or order of the queries."""

    def __init__(self, t0, w0, t1=None, w1=None, entropy=None, tol=1e-6, pool_size=24, cache_depth=9,
                 safety=None):
        if t1 is None:
            t1 = t0 + 1
        if w1 is None:
            W = None
        else:
            W = w1 - w0
        self._w0 = w0
        self._interval = brownian_interval.BrownianInterval(t0=t0, t1=t1, size=w0.shape, dtype=w0.dtype,
                                                            device=w0.device, entropy=entropy, tol=tol,
                                                            pool_size=pool_size, halfway_tree=True, W=W)
        super(BrownianTree, self).__init__()

    def __call__(self, t, tb=None, return_U=False, return_A=False):
        out = self._interval(t, tb, return_U=return_U, return_A=return_A)
        if tb is None and not return_U and not return_A:
            out = out + self._w0
        return out

    def __repr__(self):
        return f"{self.__class__.__name__}(interval={self._interval})"

    @property
    def dtype(self):
        return self._interval.dtype

    @property
    def device(self):
        return self._interval.device

    @property
    def shape(self):
        return self._interval.shape

    @property
    def levy_area_approximation(self):
        return self._interval.levy_area_approximation


def brownian_interval_like(y, t0=0., t1=1., size=None, dtype=None, device=None, **kwargs):
    """BrownianInterval with the size, dtype and device of `y` (derived.py:194-205)."""
    size = y.shape if size is None else size
    dtype = y.dtype if dtype is None else dtype
    device = y.device if device is None else device
    return brownian_interval.BrownianInterval(t0=t0, t1=t1, size=size, dtype=dtype, device=device, **kwargs)
