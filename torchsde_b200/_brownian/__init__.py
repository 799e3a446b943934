from .brownian_base import BaseBrownian
from .interval import BrownianInterval, GridBinding
from .derived import ReverseBrownian, BrownianPath, BrownianTree, brownian_interval_like
