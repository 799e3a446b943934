from torchsde_b200.settings import *  # noqa: F401,F403
from torchsde_b200.settings import LEVY_AREA_APPROXIMATIONS, METHODS, METHOD_OPTIONS, NOISE_TYPES, SDE_TYPES  # noqa: F401
