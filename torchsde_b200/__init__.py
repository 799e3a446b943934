"""torchsde_b200 — B200 (sm_100a) native SDE-integration core behind the torchsde API.

The eleven public names are the ones the reference exports (``torchsde/__init__.py:15-19``); their signatures are
checked against the reference's by the test suite.
"""
from ._core.sdeint import sdeint
from ._core.adjoint import sdeint_adjoint
from ._core.base_sde import BaseSDE
from ._core.base_sde import SDEIto
from ._core.base_sde import SDEStratonovich
from ._brownian import BaseBrownian
from ._brownian import BrownianInterval
from ._brownian import BrownianPath
from ._brownian import BrownianTree
from ._brownian import ReverseBrownian
from ._brownian import brownian_interval_like

__all__ = ['sdeint', 'sdeint_adjoint', 'BaseSDE', 'SDEIto', 'SDEStratonovich', 'BaseBrownian', 'BrownianInterval',
           'BrownianPath', 'BrownianTree', 'ReverseBrownian', 'brownian_interval_like']
__version__ = '0.1.0'
