"""torchsde_b200 — B200 (sm_100a) native SDE-integration core behind the torchsde API.

Public names mirror ``torchsde/__init__.py:15-19`` of the reference.
"""
from ._brownian import (BaseBrownian, BrownianInterval, BrownianPath, BrownianTree, ReverseBrownian,
                        brownian_interval_like)
from ._core.base_sde import BaseSDE, SDEIto, SDEStratonovich
from ._core.sdeint import sdeint
from ._core.adjoint import sdeint_adjoint

BrownianInterval.__init__.__annotations__ = {}

__version__ = '0.1.0'
